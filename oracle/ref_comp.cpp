// ORACLE -- test infrastructure only (see ref_internal.h).
// Colour spaces, gradients, dither, the 28 operators (integer + float) and the
// stride / surface compositors: src/color.zig, src/internal/color_vector.zig,
// src/gradient.zig, src/Dither.zig, src/compositor.zig.
#include <algorithm>

#include "blue_noise_table.h"
#include "ref_internal.h"

namespace zref {

// ------------------------------------------------------------------ colour
static inline float clampf(float v, float lo, float hi) { return std::max(lo, std::min(v, hi)); }
static inline float zmodf(float a, float b) {  // Zig @mod (floored), b > 0
  float r = std::fmod(a, b);
  if (r < 0) r += b;
  return r;
}
static inline float lerpf(float a, float b, float t) { return a + (b - a) * t; }  // color.zig:563-565

static const float kGamma = 2.2f;  // color.zig:146-149

// HSL.toRGB (color.zig:459-478)
static float hsl_channel(float n, float hue, float sat, float light) {
  float k = std::fmod(n + hue / 30.0f, 12.0f);  // @rem
  float a = sat * std::min(light, 1.0f - light);
  return light - a * std::max(-1.0f, std::min(std::min(k - 3.0f, 9.0f - k), 1.0f));
}
static RGBAF hsl_to_rgb(RGBAF h) {  // h,s,l,a in r,g,b,a slots
  float hue = std::fmod(h.r, 360.0f);
  if (hue < 0) hue += 360.0f;
  return {hsl_channel(0, hue, h.g, h.b), hsl_channel(8, hue, h.g, h.b), hsl_channel(4, hue, h.g, h.b), h.a};
}
// HSL.fromRGB (color.zig:419-457)
static RGBAF hsl_from_rgb(RGBAF s) {
  float mx = std::max(s.r, std::max(s.g, s.b));
  float mn = std::min(s.r, std::min(s.g, s.b));
  float range = mx - mn;
  float light = (mn + mx) / 2;
  float sat = (light == 0 || light == 1) ? 0 : (mx - light) / std::min(light, 1 - light);
  float hue = 0;
  if (range != 0) {
    if (mx == s.r)
      hue = 60 * zmodf((s.g - s.b) / range, 6);
    else if (mx == s.g)
      hue = 60 * ((s.b - s.r) / range + 2);
    else if (mx == s.b)
      hue = 60 * ((s.r - s.g) / range + 4);
  }
  if (sat < 0) {
    hue += 180;
    if (hue >= 360) hue -= 360;
    sat = std::fabs(sat);
  }
  return {hue, sat, light, s.a};
}

// LinearRGB.fromColor / SRGB.fromColor / HSL.fromColor (color.zig:166-195, 408-413)
static RGBAF color_to_linear(const z2d_color& c) {
  RGBAF v{c.c[0], c.c[1], c.c[2], c.c[3]};
  switch (c.space) {
    case Z2D_COLOR_LINEAR_RGB: return v;
    case Z2D_COLOR_SRGB:  // removeGamma, then applyGamma with gamma 1 (identity)
      return {std::pow(v.r, kGamma), std::pow(v.g, kGamma), std::pow(v.b, kGamma), v.a};
    default: return hsl_to_rgb(v);
  }
}
static RGBAF color_to_srgb(const z2d_color& c) {
  RGBAF v{c.c[0], c.c[1], c.c[2], c.c[3]};
  if (c.space == Z2D_COLOR_SRGB) return v;
  RGBAF lin = (c.space == Z2D_COLOR_LINEAR_RGB) ? v : hsl_to_rgb(v);  // removeGamma of linear == identity
  const float inv = 1 / kGamma;
  return {std::pow(lin.r, inv), std::pow(lin.g, inv), std::pow(lin.b, inv), lin.a};
}
static RGBAF color_to_hsl(const z2d_color& c) {
  if (c.space == Z2D_COLOR_HSL) return {c.c[0], c.c[1], c.c[2], c.c[3]};
  return hsl_from_rgb(color_to_linear(c));
}

// ---------------------------------------------------------------- gradients
bool Xf::inverse(Xf& o) const {  // Transformation.zig:103-162
  if (by == 0 && cx == 0) {
    if (ax == 0 || dy == 0) return false;
    if (ax != 1 || dy != 1) {
      o = {1 / ax, 0, 0, 1 / dy, -tx / ax, -ty / dy};
    } else {
      o = {1, 0, 0, 1, -tx, -ty};
    }
    return true;
  }
  double d = det();
  if (d == 0) return false;
  double k = 1 / d;
  Xf adj{dy, -by, -cx, ax, by * ty - dy * tx, cx * tx - ax * ty};
  o = {adj.ax * k, adj.by * k, adj.cx * k, adj.dy * k, adj.tx * k, adj.ty * k};
  return true;
}

void grad_prepare(const z2d_gradient& g, Grad& o) {
  o.type = g.type;
  o.method = g.method;
  o.polar = g.polar;
  for (int i = 0; i < 6; i++) o.geom[i] = g.geom[i];
  o.inv = {g.inv_ctm[0], g.inv_ctm[1], g.inv_ctm[2], g.inv_ctm[3], g.inv_ctm[4], g.inv_ctm[5]};
  o.inv_identity = o.inv.is_identity();
  if (g.type == Z2D_GRADIENT_RADIAL) {  // Radial.initBuffer (gradient.zig:262-300)
    o.inner_r = std::max(0.0, g.geom[2]);
    o.outer_r = std::max(0.0, g.geom[5]);
    o.cdx = g.geom[3] - g.geom[0];
    o.cdy = g.geom[4] - g.geom[1];
    o.dr = o.outer_r - o.inner_r;
    o.min_dr = -o.inner_r;
    double a = 0;
    a += o.cdx * o.cdx;
    a += o.cdy * o.cdy;
    a += o.dr * -o.dr;
    o.a = a;
    o.inv_a = (a != 0) ? 1 / a : 0;
  } else if (g.type == Z2D_GRADIENT_CONIC) {  // Conic.initBuffer (gradient.zig:689)
    double two_pi = M_PI * 2;
    double r = std::fmod(g.geom[2], two_pi);
    if (r < 0) r += two_pi;
    o.geom[2] = r;
  }
  o.offsets.clear();
  o.colors.clear();
  for (uint32_t i = 0; i < g.n_stops; i++) {
    o.offsets.push_back(g.stops[i].offset);
    const z2d_color& c = g.stops[i].color;
    o.colors.push_back(g.method == Z2D_INTERP_LINEAR_RGB ? color_to_linear(c)
                       : g.method == Z2D_INTERP_SRGB     ? color_to_srgb(c)
                                                         : color_to_hsl(c));
  }
}

float grad_offset(const Grad& g, int x, int y) {
  if (g.type == Z2D_GRADIENT_RADIAL && g.inner_r == 0 && g.outer_r == 0) return -1;  // gradient.zig:611
  double px = (double)x + 0.5, py = (double)y + 0.5;
  if (!g.inv_identity) g.inv.point(px, py);
  switch (g.type) {
    case Z2D_GRADIENT_LINEAR: {  // gradient.zig:349-372
      double ex = g.geom[2] - g.geom[0], ey = g.geom[3] - g.geom[1];
      double dist = 0;
      dist += ex * ex;
      dist += ey * ey;
      if (dist == 0) return -1;
      double inv_dist = 1 / dist;
      double sx = px - g.geom[0], sy = py - g.geom[1];
      double d = 0;
      d += ex * sx;
      d += ey * sy;
      double v = d * inv_dist;
      return (float)std::max(0.0, std::min(v, 1.0));
    }
    case Z2D_GRADIENT_RADIAL: {  // gradient.zig:605-648
      double pdx = px - g.geom[0], pdy = py - g.geom[1];
      double b = 0;
      b += pdx * g.cdx;
      b += pdy * g.cdy;
      b += g.inner_r * g.dr;
      double c = 0;
      c += pdx * pdx;
      c += pdy * pdy;
      c += -g.inner_r * g.inner_r;
      if (g.a == 0) {
        if (b == 0) return -1;
        double t = 0.5 * c / b;
        if (t * g.dr >= g.min_dr) return (float)std::max(0.0, std::min(t, 1.0));
        return -1;
      }
      double discr = 0;
      discr += b * b;
      discr += g.a * -c;
      if (discr >= 0) {
        double sq = std::sqrt(discr);
        double t0 = (b + sq) * g.inv_a, t1 = (b - sq) * g.inv_a;
        if (t0 * g.dr >= g.min_dr) return (float)std::max(0.0, std::min(t0, 1.0));
        if (t1 * g.dr >= g.min_dr) return (float)std::max(0.0, std::min(t1, 1.0));
      }
      return -1;
    }
    default: {  // conic, gradient.zig:731-741
      double dx = px - g.geom[0], dy = py - g.geom[1];
      double two_pi = M_PI * 2;
      double ang = std::fmod(std::atan2(dy, dx) - g.geom[2], two_pi);
      if (ang < 0) ang += two_pi;
      return (float)(ang / two_pi);
    }
  }
}

StopHit grad_search(const Grad& g, float offset) {  // gradient.zig:832-897
  const size_t n = g.offsets.size();
  if (offset < 0 || n == 0) return {{0, 0, 0, 0}, {0, 0, 0, 0}, 0};
  float off = std::min(offset, 1.0f);
  size_t left = 0, right = n, mid = 0;
  while (left < right) {
    mid = left + (right - left) / 2;
    if (off >= g.offsets[mid] && (mid == n - 1 || off <= g.offsets[mid + 1])) break;
    if (off < g.offsets[mid]) {
      right = mid;
      continue;
    }
    if (off > g.offsets[mid]) {
      left = mid + 1;
      continue;
    }
  }
  if (mid == n - 1) return {g.colors[mid], g.colors[mid], off - g.offsets[mid]};
  if (mid == 0 && off < g.offsets[mid]) return {g.colors[mid], g.colors[mid], off / g.offsets[mid]};
  float start = g.offsets[mid], end = g.offsets[mid + 1];
  float rel = end - start;
  float t = (rel != 0) ? (off - start) / rel : 0;
  return {g.colors[mid], g.colors[mid + 1], t};
}

static inline int round255(float v) { return (int)std::round(255.0f * v); }
static RGBA16 encode_raw(RGBAF c) { return {round255(c.r), round255(c.g), round255(c.b), round255(c.a)}; }
static RGBA16 premul16(RGBA16 v) { return {v.r * v.a / 255, v.g * v.a / 255, v.b * v.a / 255, v.a}; }  // pixel_vector.zig:18-25
static RGBAF demul_f(RGBAF c) {  // color_vector.zig:286-293
  if (c.a == 0) return {0, 0, 0, c.a};
  return {c.r / c.a, c.g / c.a, c.b / c.a, c.a};
}

// RGBVec.interpolate*: premultiply, lerp (color_vector.zig:295-342)
static RGBAF rgb_lerp_premul(RGBAF a, RGBAF b, float t) {
  RGBAF am{a.r * a.a, a.g * a.a, a.b * a.a, a.a}, bm{b.r * b.a, b.g * b.a, b.b * b.a, b.a};
  return {lerpf(am.r, bm.r, t), lerpf(am.g, bm.g, t), lerpf(am.b, bm.b, t), lerpf(am.a, bm.a, t)};
}
// HSL.interpolateVec (color_vector.zig:404-448) -> demultiplied HSL
static RGBAF hsl_interp(RGBAF a, RGBAF b, float t, uint32_t polar) {
  RGBAF am{a.r, a.g * a.a, a.b * a.a, a.a}, bm{b.r, b.g * b.a, b.b * b.a, b.a};
  float ah = am.r, bh = bm.r;
  switch (polar) {
    case Z2D_POLAR_SHORTER: {
      bool gt = bh - ah > 180.0f, lt = bh - ah < -180.0f;
      if (gt) ah = ah + 360.0f;
      if (lt) bh = bh + 360.0f;
      break;
    }
    case Z2D_POLAR_LONGER: {
      float d = bh - ah;
      bool c0 = (0 < d) && (d < 180.0f), c1 = (-180.0f < d) && (d <= 0);
      if (c0) ah = ah + 360.0f;
      if (c1) bh = bh + 360.0f;
      break;
    }
    case Z2D_POLAR_INCREASING:
      if (bh < ah) bh = bh + 360.0f;
      break;
    default:
      if (ah < bh) ah = ah + 360.0f;
  }
  float h = am.r + (bh - ah) * t;  // lerpPolar(a_mul.h, a_mul_h, b_mul_h, t)
  RGBAF r{zmodf(h, 360.0f), lerpf(am.g, bm.g, t), lerpf(am.b, bm.b, t), lerpf(am.a, bm.a, t)};
  if (r.a == 0) return {r.r, 0, 0, r.a};
  return {r.r, r.g / r.a, r.b / r.a, r.a};
}

// color_vector.interpolateEncodeVec (54-78): premultiplied RGBA8 for the integer pipeline.
RGBA16 grad_encode(const Grad& g, const StopHit& h) {
  switch (g.method) {
    case Z2D_INTERP_LINEAR_RGB: return encode_raw(rgb_lerp_premul(h.c0, h.c1, h.t));
    case Z2D_INTERP_SRGB: {
      RGBAF d = demul_f(rgb_lerp_premul(h.c0, h.c1, h.t));
      RGBAF lin{std::pow(d.r, kGamma), std::pow(d.g, kGamma), std::pow(d.b, kGamma), d.a};
      return premul16(encode_raw(lin));
    }
    default: return premul16(encode_raw(hsl_to_rgb(hsl_interp(h.c0, h.c1, h.t, g.polar))));
  }
}
// color_vector.interpolateVec (28-52): demultiplied *linear* colour (float pipeline, dither).
RGBAF grad_linear(const Grad& g, const StopHit& h) {
  switch (g.method) {
    case Z2D_INTERP_LINEAR_RGB: return demul_f(rgb_lerp_premul(h.c0, h.c1, h.t));
    case Z2D_INTERP_SRGB: {
      RGBAF d = demul_f(rgb_lerp_premul(h.c0, h.c1, h.t));
      return {std::pow(d.r, kGamma), std::pow(d.g, kGamma), std::pow(d.b, kGamma), d.a};
    }
    default: return hsl_to_rgb(hsl_interp(h.c0, h.c1, h.t, g.polar));
  }
}

// ------------------------------------------------------------------ dither
static float m_bayer(int x, int y) {  // Dither.zig:141-147
  int _y = y ^ x;
  unsigned m = (unsigned)((_y & 1) << 5 | (x & 1) << 4 | (_y & 2) << 2 | (x & 2) << 1 | (_y & 4) >> 1 | (x & 4) >> 2);
  return (float)m * (2.0f / 128.0f) - (63.0f / 128.0f);
}
static float m_blue(int x, int y) {  // Dither.zig:149-158 (x selects the row)
  int xm = ((x % 64) + 64) % 64, ym = ((y % 64) + 64) % 64;
  unsigned m = z2d_blue_noise_64x64[(xm << 6) | ym];
  return (float)m * (2.0f / 8192.0f) - (4095.0f / 8192.0f);
}
// color_vector.fromDitherVec (90-186): linear demultiplied, dithered.
static RGBAF dither_linear(const Src& s, int x, int y) {
  RGBAF c;
  if (s.dither_source == Z2D_DITHER_SRC_GRADIENT)
    c = grad_linear(s.grad, grad_search(s.grad, grad_offset(s.grad, x, y)));
  else
    c = s.dither_color;
  float m;
  switch (s.dither_type) {
    case Z2D_DITHER_BAYER: m = m_bayer(x, y); break;
    case Z2D_DITHER_BLUE_NOISE: m = m_blue(x, y); break;
    default: return c;
  }
  float scale = 1.0f / (float)((1 << s.dither_scale) - 1);
  auto ap = [&](float v) { return clampf(v + m * scale, 0.0f, 1.0f); };
  return {ap(c.r), ap(c.g), ap(c.b), ap(c.a)};
}

void src_from_pattern(const z2d_pattern& p, Src& o) {
  switch (p.kind) {
    case Z2D_PATTERN_OPAQUE:
      o.kind = Z2D_PARAM_PIXEL;
      o.px = p.pixel;
      break;
    case Z2D_PATTERN_GRADIENT:
      o.kind = Z2D_PARAM_GRADIENT;
      grad_prepare(*p.gradient, o.grad);
      break;
    default:
      o.kind = Z2D_PARAM_DITHER;
      o.dither_type = p.dither_type;
      o.dither_source = p.dither_source;
      o.dither_scale = p.dither_scale;
      if (p.dither_source == Z2D_DITHER_SRC_GRADIENT) {
        grad_prepare(*p.gradient, o.grad);
      } else if (p.dither_source == Z2D_DITHER_SRC_COLOR) {
        o.dither_color = color_to_linear(p.dither_color);
      } else {
        // LinearRGB.decodeRGBA(RGBA.fromPixel(px)) (color.zig:204-212): integer
        // demultiply, /255, gamma 1.
        RGBA16 v = px_to_rgba16(p.pixel);
        if (v.a == 0)
          v = {0, 0, 0, 0};
        else
          v = {v.r * 255 / v.a, v.g * 255 / v.a, v.b * 255 / v.a, v.a};
        o.dither_color = {(float)v.r / 255.0f, (float)v.g / 255.0f, (float)v.b / 255.0f, (float)v.a / 255.0f};
      }
  }
}

void src_from_param(const z2d_comp_param& p, Src& o) {
  switch (p.kind) {
    case Z2D_PARAM_NONE: o.kind = Z2D_PARAM_NONE; break;
    case Z2D_PARAM_SURFACE:
      o.kind = Z2D_PARAM_SURFACE;
      o.sfc = (const Sfc*)p.surface;
      break;
    default: src_from_pattern(p.pattern, o);
  }
}

// RGBA16Vec.from{Pixel,Gradient,Dither} / RGBAFloat.Vector.from* (compositor.zig:705-1153)
static RGBA16 src_int(const Src& s, int x, int y, size_t sfc_idx) {
  switch (s.kind) {
    case Z2D_PARAM_PIXEL: return px_to_rgba16(s.px);
    case Z2D_PARAM_GRADIENT: return grad_encode(s.grad, grad_search(s.grad, grad_offset(s.grad, x, y)));
    case Z2D_PARAM_DITHER: return premul16(encode_raw(dither_linear(s, x, y)));  // LinearRGB.encodeRGBAVec
    default: return sfc_load(*s.sfc, sfc_idx);
  }
}
static RGBAF decode_raw(RGBA16 v) { return {(float)v.r / 255.0f, (float)v.g / 255.0f, (float)v.b / 255.0f, (float)v.a / 255.0f}; }
static RGBAF src_float(const Src& s, int x, int y, size_t sfc_idx) {
  switch (s.kind) {
    case Z2D_PARAM_PIXEL: return decode_raw(px_to_rgba16(s.px));
    case Z2D_PARAM_GRADIENT: return grad_linear(s.grad, grad_search(s.grad, grad_offset(s.grad, x, y)));  // de-multiplied (sic)
    case Z2D_PARAM_DITHER: return dither_linear(s, x, y);
    default: return decode_raw(sfc_load(*s.sfc, sfc_idx));
  }
}

// --------------------------------------------------------- integer operators
bool op_requires_float(uint32_t op) {
  switch (op) {
    case Z2D_OP_COLOR_DODGE: case Z2D_OP_COLOR_BURN: case Z2D_OP_SOFT_LIGHT: case Z2D_OP_HUE:
    case Z2D_OP_SATURATION: case Z2D_OP_COLOR: case Z2D_OP_LUMINOSITY: return true;
    default: return false;
  }
}
bool op_is_bounded(uint32_t op) {
  switch (op) {
    case Z2D_OP_SRC_IN: case Z2D_OP_DST_IN: case Z2D_OP_SRC_OUT: case Z2D_OP_DST_ATOP: return false;
    default: return true;
  }
}

static inline int M(int a, int b) { return a * b / 255; }           // mul (1518-1523), truncating
static inline int IM(int a, int b) { return M(a, 255 - b); }       // invMul
static inline int RM(int a, int b) { return M(a, 255 + b); }       // rInvMul

RGBA16 int_op(uint32_t op, RGBA16 d, RGBA16 s) {
  const int sa = s.a, da = d.a;
  const int so = sa + da - M(sa, da);  // "source-over" alpha used by all blend modes
  auto ch3 = [&](auto f, int a) { return RGBA16{f(s.r, d.r), f(s.g, d.g), f(s.b, d.b), a}; };
  switch (op) {
    case Z2D_OP_SRC: return s;
    case Z2D_OP_DST: return d;
    case Z2D_OP_SRC_OVER: return ch3([&](int sc, int dc) { return sc + IM(dc, sa); }, so);
    case Z2D_OP_DST_OVER: return ch3([&](int sc, int dc) { return dc + IM(sc, da); }, so);
    case Z2D_OP_SRC_IN: return ch3([&](int sc, int) { return M(sc, da); }, M(sa, da));
    case Z2D_OP_DST_IN: return ch3([&](int, int dc) { return M(dc, sa); }, M(da, sa));
    case Z2D_OP_SRC_OUT: return ch3([&](int sc, int) { return IM(sc, da); }, IM(sa, da));
    case Z2D_OP_DST_OUT: return ch3([&](int, int dc) { return IM(dc, sa); }, IM(da, sa));
    case Z2D_OP_SRC_ATOP: return ch3([&](int sc, int dc) { return M(sc, da) + IM(dc, sa); }, da);
    case Z2D_OP_DST_ATOP: return ch3([&](int sc, int dc) { return M(dc, sa) + IM(sc, da); }, sa);
    case Z2D_OP_XOR: return ch3([&](int sc, int dc) { return IM(sc, da) + IM(dc, sa); }, IM(sa, da) + IM(da, sa));
    case Z2D_OP_PLUS: return ch3([&](int sc, int dc) { return std::min(255, sc + dc); }, std::min(255, sa + da));
    case Z2D_OP_MULTIPLY: return ch3([&](int sc, int dc) { return M(sc, dc) + IM(sc, da) + IM(dc, sa); }, so);
    case Z2D_OP_SCREEN: return ch3([&](int sc, int dc) { return sc + dc - M(sc, dc); }, so);
    case Z2D_OP_OVERLAY:  // 1356-1377
      return ch3(
          [&](int sc, int dc) {
            if (2 * dc <= da) return M(2 * sc, dc) + IM(sc, da) + IM(dc, sa);
            return RM(sc, da) + RM(dc, sa) - M(2 * dc, sc) - M(da, sa);
          },
          so);
    case Z2D_OP_DARKEN: return ch3([&](int sc, int dc) { return std::min(M(sc, da), M(dc, sa)) + IM(sc, da) + IM(dc, sa); }, so);
    case Z2D_OP_LIGHTEN: return ch3([&](int sc, int dc) { return std::max(M(sc, da), M(dc, sa)) + IM(sc, da) + IM(dc, sa); }, so);
    case Z2D_OP_HARD_LIGHT:  // 1453-1474
      return ch3(
          [&](int sc, int dc) {
            if (2 * sc <= sa) return M(2 * sc, dc) + IM(sc, da) + IM(dc, sa);
            return RM(sc, da) + RM(dc, sa) - M(sa, da) - M(2 * sc, dc);
          },
          so);
    case Z2D_OP_DIFFERENCE: return ch3([&](int sc, int dc) { return sc + dc - 2 * std::min(M(sc, da), M(dc, sa)); }, so);
    case Z2D_OP_EXCLUSION:
      return ch3([&](int sc, int dc) { return (M(sc, da) + M(dc, sa) - 2 * M(sc, dc)) + IM(sc, da) + IM(dc, sa); }, so);
    default: return {0, 0, 0, 0};  // clear + the 7 float-only operators (1179-1188)
  }
}

// ----------------------------------------------------------- float operators
namespace {
struct C3 {
  float r, g, b;
};
inline float lum(C3 c) { return c.r * 0.3f + c.g * 0.59f + c.b * 0.11f; }
inline float max3(C3 c) { return std::max(std::max(c.r, c.g), c.b); }
inline float min3(C3 c) { return std::min(std::min(c.r, c.g), c.b); }
inline float sat(C3 c) { return max3(c) - min3(c); }
inline C3 set_lum(C3 c, float l) {
  float d = l - lum(c);
  return {c.r + d, c.g + d, c.b + d};
}
inline C3 set_sat(C3 c, float s) {
  float n = min3(c), x = max3(c), d = x - n;
  auto f = [&](float v) { return d == 0.0f ? 0.0f : (v - n) * s / d; };
  return {f(c.r), f(c.g), f(c.b)};
}
// Vector form of clipColor (compositor.zig:2282-2316): the x>a select is applied
// after (and overrides) the n<0 select; both read the unmodified channel.
inline C3 clip_color(C3 c, float a) {
  float l = lum(c), n = min3(c), x = max3(c);
  float t_l_n = l - n, t_x_l = x - l;
  auto f = [&](float v) {
    float r = v;
    if (n < 0.0f) r = (t_l_n == 0.0f) ? 0.0f : l + ((v - l) * l) / t_l_n;
    if (x > a) r = (t_x_l == 0.0f) ? 0.0f : l + ((v - l) * (a - l)) / t_x_l;
    return r;
  };
  return {f(c.r), f(c.g), f(c.b)};
}
inline RGBAF nonsep_out(C3 c, RGBAF d, RGBAF s) {  // toRGBA (2259-2267)
  return {s.r * (1.0f - d.a) + d.r * (1.0f - s.a) + c.r, s.g * (1.0f - d.a) + d.g * (1.0f - s.a) + c.g,
          s.b * (1.0f - d.a) + d.b * (1.0f - s.a) + c.b, s.a + d.a - s.a * d.a};
}
}  // namespace

RGBAF float_op(uint32_t op, RGBAF d, RGBAF s) {
  const float sa = s.a, da = d.a;
  const float so = sa + da - sa * da;
  auto ch3 = [&](auto f, float a) { return RGBAF{f(s.r, d.r), f(s.g, d.g), f(s.b, d.b), a}; };
  switch (op) {
    case Z2D_OP_CLEAR: return {0, 0, 0, 0};
    case Z2D_OP_SRC: return s;
    case Z2D_OP_DST: return d;
    case Z2D_OP_SRC_OVER: return ch3([&](float sc, float dc) { return sc + dc * (1.0f - sa); }, so);
    case Z2D_OP_DST_OVER: return ch3([&](float sc, float dc) { return dc + sc * (1.0f - da); }, so);
    case Z2D_OP_SRC_IN: return ch3([&](float sc, float) { return sc * da; }, sa * da);
    case Z2D_OP_DST_IN: return ch3([&](float, float dc) { return dc * sa; }, sa * da);
    case Z2D_OP_SRC_OUT: return ch3([&](float sc, float) { return sc * (1.0f - da); }, sa * (1.0f - da));
    case Z2D_OP_DST_OUT: return ch3([&](float, float dc) { return dc * (1.0f - sa); }, da * (1.0f - sa));
    case Z2D_OP_SRC_ATOP: return ch3([&](float sc, float dc) { return sc * da + dc * (1.0f - sa); }, da);
    case Z2D_OP_DST_ATOP: return ch3([&](float sc, float dc) { return dc * sa + sc * (1.0f - da); }, sa);
    case Z2D_OP_XOR:
      return ch3([&](float sc, float dc) { return sc * (1.0f - da) + dc * (1.0f - sa); }, sa + da - 2.0f * sa * da);
    case Z2D_OP_PLUS: return ch3([&](float sc, float dc) { return std::min(1.0f, sc + dc); }, std::min(1.0f, sa + da));
    case Z2D_OP_MULTIPLY: return ch3([&](float sc, float dc) { return sc * dc + sc * (1.0f - da) + dc * (1.0f - sa); }, so);
    case Z2D_OP_SCREEN: return ch3([&](float sc, float dc) { return sc + dc - sc * dc; }, so);
    case Z2D_OP_OVERLAY:
      return ch3(
          [&](float sc, float dc) {
            if (2.0f * dc <= da) return 2.0f * sc * dc + sc * (1.0f - da) + dc * (1.0f - sa);
            return sc * (1.0f + da) + dc * (1.0f + sa) - 2.0f * dc * sc - da * sa;
          },
          so);
    case Z2D_OP_DARKEN:
      return ch3([&](float sc, float dc) { return std::min(sc * da, dc * sa) + sc * (1.0f - da) + dc * (1.0f - sa); }, so);
    case Z2D_OP_LIGHTEN:
      return ch3([&](float sc, float dc) { return std::max(sc * da, dc * sa) + sc * (1.0f - da) + dc * (1.0f - sa); }, so);
    case Z2D_OP_COLOR_DODGE:  // 1826-1902
      return ch3(
          [&](float sc, float dc) {
            if (sc == sa && dc == 0.0f) return sc * (1.0f - da);
            if (sc == sa) return sa * da + sc * (1.0f - da) + dc * (1.0f - sa);
            return sa * da * std::min(1.0f, dc / da * sa / (sa - sc)) + sc * (1.0f - da) + dc * (1.0f - sa);
          },
          so);
    case Z2D_OP_COLOR_BURN:  // 1904-1981; p0 is called as p0(sca, dca, sa) -> tests dca == sa (sic)
      return ch3(
          [&](float sc, float dc) {
            if (sc == 0.0f && dc == sa) return sa * da + dc * (1.0f - sa);
            if (sc == 0.0f) return dc * (1.0f - sa);
            return sa * da * (1.0f - std::min(1.0f, (1.0f - dc / da) * sa / sc)) + sc * (1.0f - da) + dc * (1.0f - sa);
          },
          so);
    case Z2D_OP_HARD_LIGHT:
      return ch3(
          [&](float sc, float dc) {
            if (2.0f * sc <= sa) return 2.0f * sc * dc + sc * (1.0f - da) + dc * (1.0f - sa);
            return sc * (1.0f + da) + dc * (1.0f + sa) - sa * da - 2.0f * sc * dc;
          },
          so);
    case Z2D_OP_SOFT_LIGHT:  // 2046-2150
      return ch3(
          [&](float sc, float dc) {
            if (da == 0.0f) return sc;
            float m = dc / da;
            if (2.0f * sc <= sa) return dc * (sa + (2.0f * sc - sa) * (1.0f - m)) + sc * (1.0f - da) + dc * (1.0f - sa);
            if (2.0f * sc > sa && 4.0f * dc <= da)
              return dc * sa + da * (2.0f * sc - sa) * (4.0f * m * (4.0f * m + 1.0f) * (m - 1.0f) + 7.0f * m) +
                     sc * (1.0f - da) + dc * (1.0f - sa);
            return da * (2.0f * sc - sa) * (std::sqrt(m) - m) + sc - sc * da + dc;
          },
          so);
    case Z2D_OP_DIFFERENCE: return ch3([&](float sc, float dc) { return sc + dc - 2.0f * std::min(sc * da, dc * sa); }, so);
    case Z2D_OP_EXCLUSION:
      return ch3(
          [&](float sc, float dc) { return (sc * da + dc * sa - 2.0f * sc * dc) + sc * (1.0f - da) + dc * (1.0f - sa); }, so);
    case Z2D_OP_HUE: {  // 2176-2183
      C3 c{s.r * sa, s.g * sa, s.b * sa};  // fromRGBA(src, src.a)
      C3 cd{d.r, d.g, d.b};
      c = set_sat(c, sat(cd) * sa);
      c = set_lum(c, lum(cd) * sa);
      c = clip_color(c, sa * da);
      return nonsep_out(c, d, s);
    }
    case Z2D_OP_SATURATION: {
      C3 c{d.r * sa, d.g * sa, d.b * sa};
      C3 cs{s.r, s.g, s.b}, cd{d.r, d.g, d.b};
      c = set_sat(c, sat(cs) * da);
      c = set_lum(c, lum(cd) * sa);
      c = clip_color(c, sa * da);
      return nonsep_out(c, d, s);
    }
    case Z2D_OP_COLOR: {
      C3 c{s.r * da, s.g * da, s.b * da};
      C3 cd{d.r, d.g, d.b};
      c = set_lum(c, lum(cd) * sa);
      c = clip_color(c, sa * da);
      return nonsep_out(c, d, s);
    }
    default: {  // luminosity
      C3 c{d.r * sa, d.g * sa, d.b * sa};
      C3 cs{s.r, s.g, s.b};
      c = set_lum(c, lum(cs) * da);
      c = clip_color(c, sa * da);
      return nonsep_out(c, d, s);
    }
  }
}

// ------------------------------------------------------ stride compositor
void stride_run(Sfc& dst, size_t dst_idx, size_t len, int x, int y, const StrideOp* ops, size_t n_ops, uint32_t precision) {
  for (size_t i = 0; i < len; i++) {
    const int px = x + (int)i;
    if (precision == Z2D_PRECISION_INTEGER) {
      RGBA16 d{0, 0, 0, 0}, s{0, 0, 0, 0};
      for (size_t k = 0; k < n_ops; k++) {
        const StrideOp& o = ops[k];
        s = o.src ? src_int(*o.src, px, y, o.src_idx + i) : d;
        d = o.dst ? src_int(*o.dst, px, y, o.dst_idx + i) : sfc_load(dst, dst_idx + i);
        d = int_op(o.op, d, s);
      }
      sfc_store(dst, dst_idx + i, d);
    } else {
      RGBAF d{0, 0, 0, 0}, s{0, 0, 0, 0};
      for (size_t k = 0; k < n_ops; k++) {
        const StrideOp& o = ops[k];
        s = o.src ? src_float(*o.src, px, y, o.src_idx + i) : d;
        d = o.dst ? src_float(*o.dst, px, y, o.dst_idx + i) : decode_raw(sfc_load(dst, dst_idx + i));
        d = float_op(o.op, d, s);
      }
      sfc_store(dst, dst_idx + i, encode_raw(d));  // encodeRGBAVecRaw
    }
  }
}

// SurfaceCompositor.run (compositor.zig:302-440)
void surface_run(Sfc& dst, int dst_x, int dst_y, const SurfOp* ops, size_t n_ops, uint32_t precision) {
  if (n_ops == 0) return;
  if (dst_x >= dst.w || dst_y >= dst.h) return;
  for (size_t k = 0; k < n_ops; k++)
    if (op_requires_float(ops[k].op)) precision = Z2D_PRECISION_FLOAT;
  int src_w, src_h;
  switch (ops[0].src.kind) {
    case Z2D_PARAM_NONE: return;
    case Z2D_PARAM_SURFACE:
      src_w = ops[0].src.sfc->w;
      src_h = ops[0].src.sfc->h;
      break;
    default:
      if (dst_x != 0 || dst_y != 0) return;
      src_w = dst.w;
      src_h = dst.h;
  }
  const int src_start_x = dst_x < 0 ? -dst_x : 0, src_start_y = dst_y < 0 ? -dst_y : 0;
  const int width = (src_w + dst_x > dst.w) ? dst.w - dst_x : src_w;
  const int height = (src_h + dst_y > dst.h) ? dst.h - dst_y : src_h;
  if (src_start_x >= width || src_start_y >= height) return;
  const int scan_lo = src_start_y, scan_hi = std::max(scan_lo, height);
  const int dst_start_x = src_start_x + dst_x;
  const size_t scan_w = (size_t)std::max(0, width - src_start_x);
  std::vector<StrideOp> sops(n_ops);
  for (int src_y = scan_lo; src_y < scan_hi; src_y++) {
    const int dst_start_y = src_y + dst_y;
    if (dst_start_x < 0 || dst_start_y < 0 || dst_start_x >= dst.w || dst_start_y >= dst.h) continue;  // empty stride
    for (size_t k = 0; k < n_ops; k++) {
      sops[k].op = ops[k].op;
      sops[k].dst = ops[k].dst.kind == Z2D_PARAM_NONE ? nullptr : &ops[k].dst;
      sops[k].src = ops[k].src.kind == Z2D_PARAM_NONE ? nullptr : &ops[k].src;
      if (ops[k].dst.kind == Z2D_PARAM_SURFACE)  // dst override surfaces are addressed in dst space (399-401)
        sops[k].dst_idx = (size_t)ops[k].dst.sfc->w * (size_t)dst_start_y + (size_t)dst_start_x;
      if (ops[k].src.kind == Z2D_PARAM_SURFACE)  // src surfaces in src space (417-419)
        sops[k].src_idx = (size_t)ops[k].src.sfc->w * (size_t)src_y + (size_t)src_start_x;
    }
    stride_run(dst, (size_t)dst.w * (size_t)dst_start_y + (size_t)dst_start_x, scan_w, dst_start_x, dst_start_y,
               sops.data(), n_ops, precision);
  }
}

// ---------------------------------------------------------- shared.zig
static bool fill_reduces_to_source(uint32_t op, const Src& pat) {  // shared.zig:111-117
  if (pat.kind != Z2D_PARAM_PIXEL) return false;
  return op == Z2D_OP_SRC || (op == Z2D_OP_SRC_OVER && px_is_opaque(pat.px));
}

void composite_opaque(uint32_t op, Sfc& s, const Src& pat, int x, int y, size_t len, uint32_t prec) {  // shared.zig:9-46
  if (op == Z2D_OP_CLEAR) {
    sfc_clear_stride(s, x, y, len);
  } else if (fill_reduces_to_source(op, pat)) {
    sfc_paint_stride(s, x, y, len, pat.px);
  } else {
    if (x < 0 || y < 0 || x >= s.w || y >= s.h || len == 0) return;
    StrideOp o{op, nullptr, &pat};
    stride_run(s, (size_t)s.w * (size_t)y + (size_t)x, len, x, y, &o, 1, prec);
  }
}

void composite_opacity(uint32_t op, Sfc& s, const Src& pat, int x, int y, size_t len, uint32_t prec, uint8_t opacity) {  // shared.zig:50-107
  if (op == Z2D_OP_CLEAR) {
    sfc_clear_stride(s, x, y, len);
  } else if (fill_reduces_to_source(op, pat)) {
    sfc_composite_stride(s, x, y, len, pat.px, op, opacity);
  } else {
    if (x < 0 || y < 0 || x >= s.w || y >= s.h || len == 0) return;
    Src mask;
    mask.kind = Z2D_PARAM_PIXEL;
    mask.px = z2d_pixel{Z2D_FMT_ALPHA8, 0, 0, 0, opacity};
    StrideOp o[2] = {{Z2D_OP_DST_IN, &pat, &mask}, {op, nullptr, nullptr}};
    stride_run(s, (size_t)s.w * (size_t)y + (size_t)x, len, x, y, o, 2, prec);
  }
}

}  // namespace zref
