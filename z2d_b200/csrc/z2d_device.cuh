// Device-side pixel formats, sources (pixel / gradient / dither) and the 28
// compositing operators in integer and float precision.
//
// Shared by the fused tile rasteriser (raster.cu) and the surface compositor
// (composite.cu).  Semantics follow the reference's CPU compositor exactly:
//   pixel formats  src/pixel.zig:47-56,146-362,569-626; compositor.zig:675-698
//   integer ops    src/compositor.zig:1158-1568   (u16 lanes, mul = a*b/255 trunc)
//   float ops      src/compositor.zig:1571-2440   (f32, no FMA contraction)
//   gradients      src/gradient.zig:349-372,605-648,731-741,832-897
//   interpolation  src/internal/color_vector.zig:28-86,191-459
//   dither         src/internal/color_vector.zig:90-186,482-522
// The whole library is compiled with -fmad=false so that no a*b+c is fused
// (the reference never fuses; several results depend on it).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/z2d_cuda.h"

#define Z2D_HD __host__ __device__ __forceinline__
#define Z2D_D __device__ __forceinline__
#define Z2D_DN __device__ __noinline__  // only for functions whose arguments and results are values (see Z2D_LAMBDA)
// Local lambdas must be inlined too: ptxas otherwise emits them as called sub-functions of the kernel that receive GENERIC
// pointers into the caller's stack frame, and with the large flatten kernels such a call was observed to be passed a frame
// address built from a clobbered uniform register (an illegal read in k_flatten_count depending on the build).
#define Z2D_LAMBDA __attribute__((always_inline))

namespace z2d {

// ------------------------------------------------------------------ types
struct RGBA16 {  // channels 0..255 kept in ints (reference: u16 lanes)
  int r, g, b, a;
};
struct RGBAF {
  float r, g, b, a;
};

struct DevGrad {  // gradient prepared on the host (stops converted to the interpolation space)
  uint32_t type, method, polar, n_stops;
  uint32_t stop_base, inv_identity, _p0, _p1;
  double geom[6];
  double inv[6];  // ax,by,cx,dy,tx,ty
  double cdx, cdy, dr, min_dr, a, inv_a, inner_r, outer_r;
};

struct DevSrc {  // a compositor parameter / pattern, evaluated per pixel
  uint32_t kind;  // Z2D_PARAM_*
  uint32_t grad;  // index into the gradient table (GRADIENT, DITHER over gradient)
  uint32_t dither_type, dither_source, dither_scale;
  uint32_t px_rgba;    // PIXEL: RGBA16.fromPixel packed r | g<<8 | b<<16 | a<<24
  uint32_t px_format;  // PIXEL: original pixel format and raw channel values
  uint32_t px_raw;     //        r | g<<8 | b<<16 | a<<24 as given
  float dcol[4];       // DITHER over pixel/color: linear de-multiplied colour
  const uint8_t* sdata;  // SURFACE
  uint32_t sfmt;
  int32_t sw, sh;
  uint32_t _pad;
};

struct GradTables {
  const DevGrad* grads;
  const float* stop_offsets;
  const float4* stop_colors;
  const uint16_t* blue_noise;
};

// ---------------------------------------------------------------- formats
Z2D_HD int fmt_bits(uint32_t fmt) {
  return fmt == Z2D_FMT_ALPHA8 ? 8 : fmt == Z2D_FMT_ALPHA4 ? 4 : fmt == Z2D_FMT_ALPHA2 ? 2 : fmt == Z2D_FMT_ALPHA1 ? 1 : 32;
}

// Alpha(T).shlr (pixel.zig:587-626)
Z2D_HD int scale_alpha(int val, int from_bits, int to_bits) {
  if (from_bits == 1) return val * ((1 << to_bits) - 1);
  if (val == 0) return 0;
  if (from_bits == to_bits) return val;
  if (from_bits > to_bits) return val >> (from_bits - to_bits);
  int diff = to_bits - from_bits;
  if (diff == 2 || diff == 4) return (val << diff) + val;
  return (val << (to_bits - from_bits)) | (val << (to_bits - 2 * from_bits)) | (val << (to_bits - 3 * from_bits)) | val;
}

Z2D_HD int px_alpha_bits(uint32_t fmt) { return fmt == Z2D_FMT_ALPHA4 ? 4 : fmt == Z2D_FMT_ALPHA2 ? 2 : fmt == Z2D_FMT_ALPHA1 ? 1 : 8; }

// pixel.RGBA.fromPixel widened (pixel.zig:399-433, compositor.zig:675-687)
Z2D_HD RGBA16 pixel_to_rgba16(uint32_t format, int r, int g, int b, int a) {
  switch (format) {
    case Z2D_FMT_XRGB:
    case Z2D_FMT_RGB: return {r, g, b, 255};
    case Z2D_FMT_ARGB:
    case Z2D_FMT_RGBA: return {r, g, b, a};
    default: return {0, 0, 0, scale_alpha(a, px_alpha_bits(format), 8)};
  }
}

// "raw" pixel = the value stored for one pixel: 4 bytes for the 32-bit
// formats (memory order, little endian), the byte for alpha8, the 4/2/1-bit
// sample for packed formats.
Z2D_HD RGBA16 raw_to_rgba16(uint32_t fmt, uint32_t raw) {  // fromStride / fromPixelT
  switch (fmt) {
    case Z2D_FMT_ARGB: return {(int)((raw >> 16) & 255), (int)((raw >> 8) & 255), (int)(raw & 255), (int)(raw >> 24)};
    case Z2D_FMT_XRGB: return {(int)((raw >> 16) & 255), (int)((raw >> 8) & 255), (int)(raw & 255), 255};
    case Z2D_FMT_RGB: return {(int)(raw & 255), (int)((raw >> 8) & 255), (int)((raw >> 16) & 255), 255};
    case Z2D_FMT_RGBA: return {(int)(raw & 255), (int)((raw >> 8) & 255), (int)((raw >> 16) & 255), (int)(raw >> 24)};
    case Z2D_FMT_ALPHA8: return {0, 0, 0, (int)(raw & 255)};
    default: {
      int bits = fmt_bits(fmt);
      return {0, 0, 0, scale_alpha((int)raw, bits, 8)};
    }
  }
}

Z2D_HD uint32_t rgba16_to_raw(uint32_t fmt, RGBA16 v) {  // toStride / toPixelT (padding byte written as 0)
  uint32_t r = (uint32_t)v.r & 255u, g = (uint32_t)v.g & 255u, b = (uint32_t)v.b & 255u, a = (uint32_t)v.a & 255u;
  switch (fmt) {
    case Z2D_FMT_ARGB: return b | (g << 8) | (r << 16) | (a << 24);
    case Z2D_FMT_XRGB: return b | (g << 8) | (r << 16);
    case Z2D_FMT_RGB: return r | (g << 8) | (b << 16);
    case Z2D_FMT_RGBA: return r | (g << 8) | (b << 16) | (a << 24);
    case Z2D_FMT_ALPHA8: return a;
    default: return a >> (8 - fmt_bits(fmt));
  }
}

// T.fromPixel(px) for a surface of format `fmt` (pixel.zig:399-433, 569-577)
Z2D_HD uint32_t pixel_to_raw(uint32_t fmt, uint32_t pformat, int r, int g, int b, int a) {
  if (fmt <= Z2D_FMT_RGBA) return rgba16_to_raw(fmt, pixel_to_rgba16(pformat, r, g, b, a));
  int to_bits = fmt_bits(fmt);
  if (pformat == Z2D_FMT_XRGB || pformat == Z2D_FMT_RGB) return (uint32_t)((1 << to_bits) - 1);
  return (uint32_t)scale_alpha(a, px_alpha_bits(pformat), to_bits);
}

// --- raw pixel load / store on a tightly packed surface (surface.zig:373,612,632)
Z2D_D uint32_t load_raw(const uint8_t* data, uint32_t fmt, size_t idx) {
  switch (fmt) {
    case Z2D_FMT_ALPHA8: return data[idx];
    case Z2D_FMT_ALPHA4: return (data[idx >> 1] >> ((idx & 1) * 4)) & 15u;
    case Z2D_FMT_ALPHA2: return (data[idx >> 2] >> ((idx & 3) * 2)) & 3u;
    case Z2D_FMT_ALPHA1: return (data[idx >> 3] >> (idx & 7)) & 1u;
    default: return ((const uint32_t*)data)[idx];
  }
}

// Store one pixel.  Sub-byte formats share bytes between neighbouring pixels
// that other threads may be writing: update the containing aligned 32-bit word
// with atomicAnd + atomicOr (no other thread touches *these* bits).
Z2D_D void store_raw(uint8_t* data, uint32_t fmt, size_t idx, uint32_t raw) {
  switch (fmt) {
    case Z2D_FMT_ALPHA8: data[idx] = (uint8_t)raw; break;
    case Z2D_FMT_ALPHA4:
    case Z2D_FMT_ALPHA2:
    case Z2D_FMT_ALPHA1: {
      int bits = fmt_bits(fmt);
      size_t bit = idx * (size_t)bits;
      uint32_t* word = (uint32_t*)(data + ((bit >> 5) << 2));
      uint32_t sh = (uint32_t)(bit & 31);
      uint32_t mask = ((1u << bits) - 1u) << sh;
      uint32_t val = (raw << sh) & mask;
      atomicAnd(word, ~mask);
      atomicOr(word, val);
      break;
    }
    default: ((uint32_t*)data)[idx] = raw;
  }
}

// ---------------------------------------------------------- integer operators
Z2D_HD bool op_requires_float(uint32_t op) {
  return op == Z2D_OP_COLOR_DODGE || op == Z2D_OP_COLOR_BURN || op == Z2D_OP_SOFT_LIGHT || op == Z2D_OP_HUE ||
         op == Z2D_OP_SATURATION || op == Z2D_OP_COLOR || op == Z2D_OP_LUMINOSITY;
}
Z2D_HD bool op_is_bounded(uint32_t op) {
  return !(op == Z2D_OP_SRC_IN || op == Z2D_OP_DST_IN || op == Z2D_OP_SRC_OUT || op == Z2D_OP_DST_ATOP);
}

// mul, truncating (compositor.zig:1518-1523).  Operands are channel values / 255 +- alpha, never negative,
// so the unsigned form (mul.hi + shift instead of a signed division sequence) is exact.
Z2D_HD int iM(int a, int b) { return (int)((unsigned)(a * b) / 255u); }
Z2D_HD int iIM(int a, int b) { return iM(a, 255 - b); }   // invMul
Z2D_HD int iRM(int a, int b) { return iM(a, 255 + b); }   // rInvMul
Z2D_HD int imin(int a, int b) { return a < b ? a : b; }
Z2D_HD int imax(int a, int b) { return a > b ? a : b; }

Z2D_HD int int_op_ch(uint32_t op, int sc, int dc, int sa, int da) {
  switch (op) {
    case Z2D_OP_SRC: return sc;
    case Z2D_OP_DST: return dc;
    case Z2D_OP_SRC_OVER: return sc + iIM(dc, sa);
    case Z2D_OP_DST_OVER: return dc + iIM(sc, da);
    case Z2D_OP_SRC_IN: return iM(sc, da);
    case Z2D_OP_DST_IN: return iM(dc, sa);
    case Z2D_OP_SRC_OUT: return iIM(sc, da);
    case Z2D_OP_DST_OUT: return iIM(dc, sa);
    case Z2D_OP_SRC_ATOP: return iM(sc, da) + iIM(dc, sa);
    case Z2D_OP_DST_ATOP: return iM(dc, sa) + iIM(sc, da);
    case Z2D_OP_XOR: return iIM(sc, da) + iIM(dc, sa);
    case Z2D_OP_PLUS: return imin(255, sc + dc);
    case Z2D_OP_MULTIPLY: return iM(sc, dc) + iIM(sc, da) + iIM(dc, sa);
    case Z2D_OP_SCREEN: return sc + dc - iM(sc, dc);
    case Z2D_OP_OVERLAY:
      if (2 * dc <= da) return iM(2 * sc, dc) + iIM(sc, da) + iIM(dc, sa);
      return iRM(sc, da) + iRM(dc, sa) - iM(2 * dc, sc) - iM(da, sa);
    case Z2D_OP_DARKEN: return imin(iM(sc, da), iM(dc, sa)) + iIM(sc, da) + iIM(dc, sa);
    case Z2D_OP_LIGHTEN: return imax(iM(sc, da), iM(dc, sa)) + iIM(sc, da) + iIM(dc, sa);
    case Z2D_OP_HARD_LIGHT:
      if (2 * sc <= sa) return iM(2 * sc, dc) + iIM(sc, da) + iIM(dc, sa);
      return iRM(sc, da) + iRM(dc, sa) - iM(sa, da) - iM(2 * sc, dc);
    case Z2D_OP_DIFFERENCE: return sc + dc - 2 * imin(iM(sc, da), iM(dc, sa));
    case Z2D_OP_EXCLUSION: return (iM(sc, da) + iM(dc, sa) - 2 * iM(sc, dc)) + iIM(sc, da) + iIM(dc, sa);
    default: return 0;  // clear and the 7 float-only operators (compositor.zig:1179-1188)
  }
}

Z2D_HD int int_op_alpha(uint32_t op, int sa, int da) {
  switch (op) {
    case Z2D_OP_SRC: return sa;
    case Z2D_OP_DST: return da;
    case Z2D_OP_SRC_IN:
    case Z2D_OP_DST_IN: return iM(sa, da);
    case Z2D_OP_SRC_OUT: return iIM(sa, da);
    case Z2D_OP_DST_OUT: return iIM(da, sa);
    case Z2D_OP_SRC_ATOP: return da;
    case Z2D_OP_DST_ATOP: return sa;
    case Z2D_OP_XOR: return iIM(sa, da) + iIM(da, sa);
    case Z2D_OP_PLUS: return imin(255, sa + da);
    case Z2D_OP_CLEAR: case Z2D_OP_COLOR_DODGE: case Z2D_OP_COLOR_BURN: case Z2D_OP_SOFT_LIGHT: case Z2D_OP_HUE:
    case Z2D_OP_SATURATION: case Z2D_OP_COLOR: case Z2D_OP_LUMINOSITY: return 0;
    default: return sa + da - iM(sa, da);
  }
}

Z2D_HD RGBA16 int_op(uint32_t op, RGBA16 d, RGBA16 s) {
  return {int_op_ch(op, s.r, d.r, s.a, d.a), int_op_ch(op, s.g, d.g, s.a, d.a), int_op_ch(op, s.b, d.b, s.a, d.a),
          int_op_alpha(op, s.a, d.a)};
}

// Same operator table with ONE dispatch per pixel instead of one per channel (for call sites
// where `op` is a run-time, warp-uniform value).
#define Z2D_CH3(EXPR, AEXPR)                                                       \
  {                                                                                \
    auto f = [&](int sc, int dc) -> int { return EXPR; };                          \
    return RGBA16{f(s.r, d.r), f(s.g, d.g), f(s.b, d.b), AEXPR};                   \
  }
Z2D_HD RGBA16 int_op_sw(uint32_t op, RGBA16 d, RGBA16 s) {
  const int sa = s.a, da = d.a;
  const int so = sa + da - iM(sa, da);
  switch (op) {
    case Z2D_OP_SRC: return s;
    case Z2D_OP_DST: return d;
    case Z2D_OP_SRC_OVER: Z2D_CH3(sc + iIM(dc, sa), so)
    case Z2D_OP_DST_OVER: Z2D_CH3(dc + iIM(sc, da), so)
    case Z2D_OP_SRC_IN: Z2D_CH3(iM(sc, da), iM(sa, da))
    case Z2D_OP_DST_IN: Z2D_CH3(iM(dc, sa), iM(sa, da))
    case Z2D_OP_SRC_OUT: Z2D_CH3(iIM(sc, da), iIM(sa, da))
    case Z2D_OP_DST_OUT: Z2D_CH3(iIM(dc, sa), iIM(da, sa))
    case Z2D_OP_SRC_ATOP: Z2D_CH3(iM(sc, da) + iIM(dc, sa), da)
    case Z2D_OP_DST_ATOP: Z2D_CH3(iM(dc, sa) + iIM(sc, da), sa)
    case Z2D_OP_XOR: Z2D_CH3(iIM(sc, da) + iIM(dc, sa), iIM(sa, da) + iIM(da, sa))
    case Z2D_OP_PLUS: Z2D_CH3(imin(255, sc + dc), imin(255, sa + da))
    case Z2D_OP_MULTIPLY: Z2D_CH3(iM(sc, dc) + iIM(sc, da) + iIM(dc, sa), so)
    case Z2D_OP_SCREEN: Z2D_CH3(sc + dc - iM(sc, dc), so)
    case Z2D_OP_OVERLAY: Z2D_CH3((2 * dc <= da) ? iM(2 * sc, dc) + iIM(sc, da) + iIM(dc, sa) : iRM(sc, da) + iRM(dc, sa) - iM(2 * dc, sc) - iM(da, sa), so)
    case Z2D_OP_DARKEN: Z2D_CH3(imin(iM(sc, da), iM(dc, sa)) + iIM(sc, da) + iIM(dc, sa), so)
    case Z2D_OP_LIGHTEN: Z2D_CH3(imax(iM(sc, da), iM(dc, sa)) + iIM(sc, da) + iIM(dc, sa), so)
    case Z2D_OP_HARD_LIGHT: Z2D_CH3((2 * sc <= sa) ? iM(2 * sc, dc) + iIM(sc, da) + iIM(dc, sa) : iRM(sc, da) + iRM(dc, sa) - iM(sa, da) - iM(2 * sc, dc), so)
    case Z2D_OP_DIFFERENCE: Z2D_CH3(sc + dc - 2 * imin(iM(sc, da), iM(dc, sa)), so)
    case Z2D_OP_EXCLUSION: Z2D_CH3((iM(sc, da) + iM(dc, sa) - 2 * iM(sc, dc)) + iIM(sc, da) + iIM(dc, sa), so)
    default: return RGBA16{0, 0, 0, 0};
  }
}
#undef Z2D_CH3

// ------------------------------------------------------------ float operators
Z2D_HD float fminz(float a, float b) { return b < a ? b : a; }  // std::min / Zig @min on non-NaN; NaN in b -> a
Z2D_HD float fmaxz(float a, float b) { return a < b ? b : a; }

Z2D_HD float float_op_ch(uint32_t op, float sc, float dc, float sa, float da) {
  switch (op) {
    case Z2D_OP_CLEAR: return 0.0f;
    case Z2D_OP_SRC: return sc;
    case Z2D_OP_DST: return dc;
    case Z2D_OP_SRC_OVER: return sc + dc * (1.0f - sa);
    case Z2D_OP_DST_OVER: return dc + sc * (1.0f - da);
    case Z2D_OP_SRC_IN: return sc * da;
    case Z2D_OP_DST_IN: return dc * sa;
    case Z2D_OP_SRC_OUT: return sc * (1.0f - da);
    case Z2D_OP_DST_OUT: return dc * (1.0f - sa);
    case Z2D_OP_SRC_ATOP: return sc * da + dc * (1.0f - sa);
    case Z2D_OP_DST_ATOP: return dc * sa + sc * (1.0f - da);
    case Z2D_OP_XOR: return sc * (1.0f - da) + dc * (1.0f - sa);
    case Z2D_OP_PLUS: return fminz(1.0f, sc + dc);
    case Z2D_OP_MULTIPLY: return sc * dc + sc * (1.0f - da) + dc * (1.0f - sa);
    case Z2D_OP_SCREEN: return sc + dc - sc * dc;
    case Z2D_OP_OVERLAY:
      if (2.0f * dc <= da) return 2.0f * sc * dc + sc * (1.0f - da) + dc * (1.0f - sa);
      return sc * (1.0f + da) + dc * (1.0f + sa) - 2.0f * dc * sc - da * sa;
    case Z2D_OP_DARKEN: return fminz(sc * da, dc * sa) + sc * (1.0f - da) + dc * (1.0f - sa);
    case Z2D_OP_LIGHTEN: return fmaxz(sc * da, dc * sa) + sc * (1.0f - da) + dc * (1.0f - sa);
    case Z2D_OP_COLOR_DODGE:  // compositor.zig:1826-1902
      if (sc == sa && dc == 0.0f) return sc * (1.0f - da);
      if (sc == sa) return sa * da + sc * (1.0f - da) + dc * (1.0f - sa);
      return sa * da * fminz(1.0f, dc / da * sa / (sa - sc)) + sc * (1.0f - da) + dc * (1.0f - sa);
    case Z2D_OP_COLOR_BURN:  // compositor.zig:1904-1981; first predicate tests dca == sa (sic, 1915/1932)
      if (sc == 0.0f && dc == sa) return sa * da + dc * (1.0f - sa);
      if (sc == 0.0f) return dc * (1.0f - sa);
      return sa * da * (1.0f - fminz(1.0f, (1.0f - dc / da) * sa / sc)) + sc * (1.0f - da) + dc * (1.0f - sa);
    case Z2D_OP_HARD_LIGHT:
      if (2.0f * sc <= sa) return 2.0f * sc * dc + sc * (1.0f - da) + dc * (1.0f - sa);
      return sc * (1.0f + da) + dc * (1.0f + sa) - sa * da - 2.0f * sc * dc;
    case Z2D_OP_SOFT_LIGHT: {  // compositor.zig:2046-2150
      if (da == 0.0f) return sc;
      float m = dc / da;
      if (2.0f * sc <= sa) return dc * (sa + (2.0f * sc - sa) * (1.0f - m)) + sc * (1.0f - da) + dc * (1.0f - sa);
      if (2.0f * sc > sa && 4.0f * dc <= da)
        return dc * sa + da * (2.0f * sc - sa) * (4.0f * m * (4.0f * m + 1.0f) * (m - 1.0f) + 7.0f * m) + sc * (1.0f - da) +
               dc * (1.0f - sa);
      return da * (2.0f * sc - sa) * (sqrtf(m) - m) + sc - sc * da + dc;
    }
    case Z2D_OP_DIFFERENCE: return sc + dc - 2.0f * fminz(sc * da, dc * sa);
    case Z2D_OP_EXCLUSION: return (sc * da + dc * sa - 2.0f * sc * dc) + sc * (1.0f - da) + dc * (1.0f - sa);
    default: return 0.0f;
  }
}

struct C3 {
  float r, g, b;
};
Z2D_HD float c3_lum(C3 c) { return c.r * 0.3f + c.g * 0.59f + c.b * 0.11f; }
Z2D_HD float c3_max(C3 c) { return fmaxz(fmaxz(c.r, c.g), c.b); }
Z2D_HD float c3_min(C3 c) { return fminz(fminz(c.r, c.g), c.b); }
Z2D_HD C3 c3_set_lum(C3 c, float l) {
  float d = l - c3_lum(c);
  return {c.r + d, c.g + d, c.b + d};
}
Z2D_HD C3 c3_set_sat(C3 c, float s) {
  float n = c3_min(c), x = c3_max(c), d = x - n;
  if (d == 0.0f) return {0.0f, 0.0f, 0.0f};
  return {(c.r - n) * s / d, (c.g - n) * s / d, (c.b - n) * s / d};
}
// vector form of clipColor (compositor.zig:2282-2316): the x>a select overrides the n<0 one
Z2D_HD float clip_ch(float v, float l, float n, float x, float a) {
  float r = v;
  float t_l_n = l - n, t_x_l = x - l;
  if (n < 0.0f) r = (t_l_n == 0.0f) ? 0.0f : l + ((v - l) * l) / t_l_n;
  if (x > a) r = (t_x_l == 0.0f) ? 0.0f : l + ((v - l) * (a - l)) / t_x_l;
  return r;
}
Z2D_HD C3 c3_clip(C3 c, float a) {
  float l = c3_lum(c), n = c3_min(c), x = c3_max(c);
  return {clip_ch(c.r, l, n, x, a), clip_ch(c.g, l, n, x, a), clip_ch(c.b, l, n, x, a)};
}

Z2D_HD RGBAF float_op(uint32_t op, RGBAF d, RGBAF s) {
  const float sa = s.a, da = d.a;
  if (op >= Z2D_OP_HUE) {  // non-separable (compositor.zig:2176-2421)
    C3 c, cs{s.r, s.g, s.b}, cd{d.r, d.g, d.b};
    switch (op) {
      case Z2D_OP_HUE:
        c = {s.r * sa, s.g * sa, s.b * sa};
        c = c3_set_sat(c, (c3_max(cd) - c3_min(cd)) * sa);
        c = c3_set_lum(c, c3_lum(cd) * sa);
        break;
      case Z2D_OP_SATURATION:
        c = {d.r * sa, d.g * sa, d.b * sa};
        c = c3_set_sat(c, (c3_max(cs) - c3_min(cs)) * da);
        c = c3_set_lum(c, c3_lum(cd) * sa);
        break;
      case Z2D_OP_COLOR:
        c = {s.r * da, s.g * da, s.b * da};
        c = c3_set_lum(c, c3_lum(cd) * sa);
        break;
      default:
        c = {d.r * sa, d.g * sa, d.b * sa};
        c = c3_set_lum(c, c3_lum(cs) * da);
    }
    c = c3_clip(c, sa * da);
    return {s.r * (1.0f - da) + d.r * (1.0f - sa) + c.r, s.g * (1.0f - da) + d.g * (1.0f - sa) + c.g,
            s.b * (1.0f - da) + d.b * (1.0f - sa) + c.b, sa + da - sa * da};
  }
  float a;
  switch (op) {
    case Z2D_OP_CLEAR: a = 0.0f; break;
    case Z2D_OP_SRC: a = sa; break;
    case Z2D_OP_DST: a = da; break;
    case Z2D_OP_SRC_IN:
    case Z2D_OP_DST_IN: a = sa * da; break;
    case Z2D_OP_SRC_OUT: a = sa * (1.0f - da); break;
    case Z2D_OP_DST_OUT: a = da * (1.0f - sa); break;
    case Z2D_OP_SRC_ATOP: a = da; break;
    case Z2D_OP_DST_ATOP: a = sa; break;
    case Z2D_OP_XOR: a = sa + da - 2.0f * sa * da; break;
    case Z2D_OP_PLUS: a = fminz(1.0f, sa + da); break;
    default: a = sa + da - sa * da;
  }
  return {float_op_ch(op, s.r, d.r, sa, da), float_op_ch(op, s.g, d.g, sa, da), float_op_ch(op, s.b, d.b, sa, da), a};
}

// -------------------------------------------------------------- encode / decode
Z2D_HD float round_half_away_f(float v) {
  float t = truncf(v);
  float f = v - t;
  if (fabsf(f) >= 0.5f) t += (v < 0.0f ? -1.0f : 1.0f);
  return t;
}
Z2D_HD int round255(float c) { return (int)round_half_away_f(255.0f * c); }
Z2D_HD RGBA16 encode_raw(RGBAF c) { return {round255(c.r), round255(c.g), round255(c.b), round255(c.a)}; }  // encodeRGBAVecRaw
Z2D_HD RGBAF decode_raw(RGBA16 v) { return {(float)v.r / 255.0f, (float)v.g / 255.0f, (float)v.b / 255.0f, (float)v.a / 255.0f}; }
Z2D_HD RGBA16 premul16(RGBA16 v) { return {v.r * v.a / 255, v.g * v.a / 255, v.b * v.a / 255, v.a}; }  // pixel_vector.zig:18-25
Z2D_HD RGBAF demul_f(RGBAF c) {
  if (c.a == 0.0f) return {0.0f, 0.0f, 0.0f, c.a};
  return {c.r / c.a, c.g / c.a, c.b / c.a, c.a};
}

// ------------------------------------------------------------------ gradients
Z2D_D float zmodf(float a, float b) {
  float r = fmodf(a, b);
  if (r < 0.0f) r += b;
  return r;
}
Z2D_D float lerpf(float a, float b, float t) { return a + (b - a) * t; }

Z2D_D float grad_offset(const DevGrad& g, int x, int y) {
  if (g.type == Z2D_GRADIENT_RADIAL && g.inner_r == 0.0 && g.outer_r == 0.0) return -1.0f;  // gradient.zig:611
  double px = (double)x + 0.5, py = (double)y + 0.5;
  if (!g.inv_identity) {
    double ix = px, iy = py;
    px = g.inv[0] * ix + g.inv[1] * iy;
    py = g.inv[2] * ix + g.inv[3] * iy;
    px += g.inv[4];
    py += g.inv[5];
  }
  if (g.type == Z2D_GRADIENT_LINEAR) {  // gradient.zig:349-372
    double ex = g.geom[2] - g.geom[0], ey = g.geom[3] - g.geom[1];
    double dist = 0.0;
    dist += ex * ex;
    dist += ey * ey;
    if (dist == 0.0) return -1.0f;
    double inv_dist = 1.0 / dist;
    double sx = px - g.geom[0], sy = py - g.geom[1];
    double d = 0.0;
    d += ex * sx;
    d += ey * sy;
    double v = d * inv_dist;
    v = v < 1.0 ? v : 1.0;
    v = v > 0.0 ? v : 0.0;
    return (float)v;
  }
  if (g.type == Z2D_GRADIENT_RADIAL) {  // gradient.zig:605-648
    double pdx = px - g.geom[0], pdy = py - g.geom[1];
    double b = 0.0;
    b += pdx * g.cdx;
    b += pdy * g.cdy;
    b += g.inner_r * g.dr;
    double c = 0.0;
    c += pdx * pdx;
    c += pdy * pdy;
    c += -g.inner_r * g.inner_r;
    double t;
    if (g.a == 0.0) {
      if (b == 0.0) return -1.0f;
      t = 0.5 * c / b;
      if (!(t * g.dr >= g.min_dr)) return -1.0f;
    } else {
      double discr = 0.0;
      discr += b * b;
      discr += g.a * -c;
      if (!(discr >= 0.0)) return -1.0f;
      double sq = sqrt(discr);
      double t0 = (b + sq) * g.inv_a, t1 = (b - sq) * g.inv_a;
      if (t0 * g.dr >= g.min_dr)
        t = t0;
      else if (t1 * g.dr >= g.min_dr)
        t = t1;
      else
        return -1.0f;
    }
    t = t < 1.0 ? t : 1.0;
    t = t > 0.0 ? t : 0.0;
    return (float)t;
  }
  // conic (gradient.zig:731-741)
  double dx = px - g.geom[0], dy = py - g.geom[1];
  const double two_pi = 6.283185307179586476925286766559;
  double ang = fmod(atan2(dy, dx) - g.geom[2], two_pi);
  if (ang < 0.0) ang += two_pi;
  return (float)(ang / two_pi);
}

struct StopHit {
  float4 c0, c1;
  float t;
};

Z2D_D StopHit grad_search(const DevGrad& g, const GradTables& T, float offset) {  // gradient.zig:832-897
  const uint32_t n = g.n_stops;
  StopHit h;
  if (offset < 0.0f || n == 0) {
    h.c0 = h.c1 = make_float4(0.f, 0.f, 0.f, 0.f);
    h.t = 0.f;
    return h;
  }
  const float* offs = T.stop_offsets + g.stop_base;
  const float4* cols = T.stop_colors + g.stop_base;
  float off = fminz(offset, 1.0f);
  uint32_t left = 0, right = n, mid = 0;
  while (left < right) {
    mid = left + (right - left) / 2;
    float om = offs[mid];
    if (off >= om && (mid == n - 1 || off <= offs[mid + 1])) break;
    if (off < om) {
      right = mid;
      continue;
    }
    if (off > om) {
      left = mid + 1;
      continue;
    }
  }
  if (mid == n - 1) {
    h.c0 = h.c1 = cols[mid];
    h.t = off - offs[mid];
    return h;
  }
  if (mid == 0 && off < offs[mid]) {
    h.c0 = h.c1 = cols[mid];
    h.t = off / offs[mid];
    return h;
  }
  float start = offs[mid], end = offs[mid + 1];
  float rel = end - start;
  h.c0 = cols[mid];
  h.c1 = cols[mid + 1];
  h.t = (rel != 0.0f) ? (off - start) / rel : 0.0f;
  return h;
}

Z2D_D RGBAF rgb_lerp_premul(float4 a, float4 b, float t) {  // color_vector.zig:295-342
  float ar = a.x * a.w, ag = a.y * a.w, ab = a.z * a.w;
  float br = b.x * b.w, bg = b.y * b.w, bb = b.z * b.w;
  return {lerpf(ar, br, t), lerpf(ag, bg, t), lerpf(ab, bb, t), lerpf(a.w, b.w, t)};
}

Z2D_D float hsl_channel(float n, float hue, float sat, float light) {  // color_vector.zig:376-386
  float k = fmodf(n + hue / 30.0f, 12.0f);
  float a = sat * fminz(light, 1.0f - light);
  return light - a * fmaxz(-1.0f, fminz(fminz(k - 3.0f, 9.0f - k), 1.0f));
}
Z2D_D RGBAF hsl_to_rgb(RGBAF h) {
  float hue = fmodf(h.r, 360.0f);
  if (hue < 0.0f) hue += 360.0f;
  return {hsl_channel(0.0f, hue, h.g, h.b), hsl_channel(8.0f, hue, h.g, h.b), hsl_channel(4.0f, hue, h.g, h.b), h.a};
}
Z2D_D RGBAF hsl_interp(float4 a, float4 b, float t, uint32_t polar) {  // color_vector.zig:404-448
  float as = a.y * a.w, al = a.z * a.w, bs = b.y * b.w, bl = b.z * b.w;
  float ah = a.x, bh = b.x;
  switch (polar) {
    case Z2D_POLAR_SHORTER: {
      bool gt = bh - ah > 180.0f, lt = bh - ah < -180.0f;
      if (gt) ah = ah + 360.0f;
      if (lt) bh = bh + 360.0f;
      break;
    }
    case Z2D_POLAR_LONGER: {
      float d = bh - ah;
      bool c0 = (0.0f < d) && (d < 180.0f), c1 = (-180.0f < d) && (d <= 0.0f);
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
  float hr = a.x + (bh - ah) * t;
  RGBAF r{zmodf(hr, 360.0f), lerpf(as, bs, t), lerpf(al, bl, t), lerpf(a.w, b.w, t)};
  if (r.a == 0.0f) return {r.r, 0.0f, 0.0f, r.a};
  return {r.r, r.g / r.a, r.b / r.a, r.a};
}

static constexpr float kGamma = 2.2f;

// interpolateEncodeVec (color_vector.zig:54-78): premultiplied RGBA8 for the integer pipeline
Z2D_D RGBA16 grad_encode(const DevGrad& g, const StopHit& h) {
  if (g.method == Z2D_INTERP_LINEAR_RGB) return encode_raw(rgb_lerp_premul(h.c0, h.c1, h.t));
  if (g.method == Z2D_INTERP_SRGB) {
    RGBAF d = demul_f(rgb_lerp_premul(h.c0, h.c1, h.t));
    RGBAF lin{powf(d.r, kGamma), powf(d.g, kGamma), powf(d.b, kGamma), d.a};
    return premul16(encode_raw(lin));
  }
  return premul16(encode_raw(hsl_to_rgb(hsl_interp(h.c0, h.c1, h.t, g.polar))));
}
// interpolateVec (color_vector.zig:28-52): de-multiplied linear colour (float pipeline, dither)
Z2D_D RGBAF grad_linear(const DevGrad& g, const StopHit& h) {
  if (g.method == Z2D_INTERP_LINEAR_RGB) return demul_f(rgb_lerp_premul(h.c0, h.c1, h.t));
  if (g.method == Z2D_INTERP_SRGB) {
    RGBAF d = demul_f(rgb_lerp_premul(h.c0, h.c1, h.t));
    return {powf(d.r, kGamma), powf(d.g, kGamma), powf(d.b, kGamma), d.a};
  }
  return hsl_to_rgb(hsl_interp(h.c0, h.c1, h.t, g.polar));
}

// --------------------------------------------------------------------- dither
Z2D_D float m_bayer(int x, int y) {  // Dither.zig:141-147
  int _y = y ^ x;
  unsigned m = (unsigned)((_y & 1) << 5 | (x & 1) << 4 | (_y & 2) << 2 | (x & 2) << 1 | (_y & 4) >> 1 | (x & 4) >> 2);
  return (float)m * (2.0f / 128.0f) - (63.0f / 128.0f);
}
Z2D_D float m_blue(const GradTables& T, int x, int y) {  // Dither.zig:149-158 (x is the major index)
  int xm = x & 63, ym = y & 63;  // @mod(x, 64) on two's complement ints
  unsigned m = T.blue_noise[(xm << 6) | ym];
  return (float)m * (2.0f / 8192.0f) - (4095.0f / 8192.0f);
}
Z2D_D float clamp01(float v) { return fmaxz(0.0f, fminz(v, 1.0f)); }

Z2D_D RGBAF dither_linear(const DevSrc& s, const GradTables& T, int x, int y) {  // color_vector.zig:90-186
  RGBAF c;
  if (s.dither_source == Z2D_DITHER_SRC_GRADIENT) {
    const DevGrad& g = T.grads[s.grad];
    c = grad_linear(g, grad_search(g, T, grad_offset(g, x, y)));
  } else {
    c = {s.dcol[0], s.dcol[1], s.dcol[2], s.dcol[3]};
  }
  float m;
  if (s.dither_type == Z2D_DITHER_BAYER)
    m = m_bayer(x, y);
  else if (s.dither_type == Z2D_DITHER_BLUE_NOISE)
    m = m_blue(T, x, y);
  else
    return c;
  float scale = 1.0f / (float)((1 << s.dither_scale) - 1);
  float ms = m * scale;
  return {clamp01(c.r + ms), clamp01(c.g + ms), clamp01(c.b + ms), clamp01(c.a + ms)};
}

// ---------------------------------------------------------- source evaluation
Z2D_D RGBA16 unpack_rgba(uint32_t v) { return {(int)(v & 255), (int)((v >> 8) & 255), (int)((v >> 16) & 255), (int)(v >> 24)}; }

// RGBA16Vec.from{Pixel,Gradient,Dither,Stride} (compositor.zig:705-800)
Z2D_D RGBA16 src_int(const DevSrc& s, const GradTables& T, int x, int y, size_t sidx) {
  switch (s.kind) {
    case Z2D_PARAM_PIXEL: return unpack_rgba(s.px_rgba);
    case Z2D_PARAM_GRADIENT: {
      const DevGrad& g = T.grads[s.grad];
      return grad_encode(g, grad_search(g, T, grad_offset(g, x, y)));
    }
    case Z2D_PARAM_DITHER: return premul16(encode_raw(dither_linear(s, T, x, y)));
    default: return raw_to_rgba16(s.sfmt, load_raw(s.sdata, s.sfmt, sidx));
  }
}
// RGBAFloat.Vector.from* (compositor.zig:1057-1131): gradient / dither sources are DE-multiplied (sic)
Z2D_D RGBAF src_float(const DevSrc& s, const GradTables& T, int x, int y, size_t sidx) {
  switch (s.kind) {
    case Z2D_PARAM_PIXEL: return decode_raw(unpack_rgba(s.px_rgba));
    case Z2D_PARAM_GRADIENT: {
      const DevGrad& g = T.grads[s.grad];
      return grad_linear(g, grad_search(g, T, grad_offset(g, x, y)));
    }
    case Z2D_PARAM_DITHER: return dither_linear(s, T, x, y);
    default: return decode_raw(raw_to_rgba16(s.sfmt, load_raw(s.sdata, s.sfmt, sidx)));
  }
}

}  // namespace z2d
