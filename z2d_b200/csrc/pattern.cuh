// Pattern sources (gradient, dither) evaluated per pixel with everything that does not depend on the pixel hoisted out:
// shared by the surface compositor (k_composite_gen, composite.cuh) and the tile kernel (k_raster_tiles, raster.cuh).
#pragma once

namespace z2d {

constexpr int kGenMaxStops = 16;

// gradient.getOffset with the row-invariant terms computed once per row (identity inverse CTM; otherwise grad_offset)
struct GradRowEval {
  const DevGrad* g;
  bool hoist;
  double ex, ey, inv_dist, eysy;          // linear
  double r0dr, nr0sq, pdy, pdycdy, pdy2;  // radial
  double dy;                              // conic
  int y;
  Z2D_D void init(const DevGrad* gg) {
    g = gg;
    hoist = g->inv_identity != 0u;
    if (g->type == Z2D_GRADIENT_LINEAR) {
      ex = g->geom[2] - g->geom[0];
      ey = g->geom[3] - g->geom[1];
      double dist = 0.0;
      dist += ex * ex;
      dist += ey * ey;
      if (dist == 0.0) hoist = false;  // (-1 for every pixel: leave it to grad_offset)
      inv_dist = 1.0 / dist;
    } else if (g->type == Z2D_GRADIENT_RADIAL) {
      r0dr = g->inner_r * g->dr;
      nr0sq = -g->inner_r * g->inner_r;
      if (g->inner_r == 0.0 && g->outer_r == 0.0) hoist = false;
    }
  }
  Z2D_D void set_row(int yy) {
    y = yy;
    const double py = (double)yy + 0.5;
    if (g->type == Z2D_GRADIENT_LINEAR) {
      const double sy = py - g->geom[1];
      eysy = ey * sy;
    } else if (g->type == Z2D_GRADIENT_RADIAL) {
      pdy = py - g->geom[1];
      pdycdy = pdy * g->cdy;
      pdy2 = pdy * pdy;
    } else {
      dy = py - g->geom[1];
    }
  }
  Z2D_D float offset(int x) const {
    if (!hoist) return grad_offset(*g, x, y);
    const double px = (double)x + 0.5;
    if (g->type == Z2D_GRADIENT_LINEAR) {  // gradient.zig:349-372
      const double sx = px - g->geom[0];
      double d = 0.0;
      d += ex * sx;
      d += eysy;
      double v = d * inv_dist;
      v = v < 1.0 ? v : 1.0;
      v = v > 0.0 ? v : 0.0;
      return (float)v;
    }
    if (g->type == Z2D_GRADIENT_RADIAL) {  // gradient.zig:605-648
      const double pdx = px - g->geom[0];
      double b = 0.0;
      b += pdx * g->cdx;
      b += pdycdy;
      b += r0dr;
      double c = 0.0;
      c += pdx * pdx;
      c += pdy2;
      c += nr0sq;
      double t;
      if (g->a == 0.0) {
        if (b == 0.0) return -1.0f;
        t = 0.5 * c / b;
        if (!(t * g->dr >= g->min_dr)) return -1.0f;
      } else {
        double discr = 0.0;
        discr += b * b;
        discr += g->a * -c;
        if (!(discr >= 0.0)) return -1.0f;
        const double sq = sqrt(discr);
        const double t0 = (b + sq) * g->inv_a, t1 = (b - sq) * g->inv_a;
        if (t0 * g->dr >= g->min_dr)
          t = t0;
        else if (t1 * g->dr >= g->min_dr)
          t = t1;
        else
          return -1.0f;
      }
      t = t < 1.0 ? t : 1.0;
      t = t > 0.0 ? t : 0.0;
      return (float)t;
    }
    const double dx = px - g->geom[0];  // gradient.zig:731-741
    const double two_pi = 6.283185307179586476925286766559;
    double ang = fmod(atan2(dy, dx) - g->geom[2], two_pi);
    if (ang < 0.0) ang += two_pi;
    return (float)(ang / two_pi);
  }
};

// alpha of the interpolated colour only: every interpolation method ends in the same f32 lerp of the stop alphas
// (color_vector.zig:295-342, 404-448), and both encodings keep it as round(255 * a) (premultiplication leaves alpha alone)
Z2D_D float hit_alpha(const StopHit& h) { return lerpf(h.c0.w, h.c1.w, h.t); }


// FC: 0 = 32-bit formats, 1 = alpha8, 2 = alpha4 / alpha2 / alpha1

// One gradient / dither source, its gradient and stops staged in shared memory by the caller (pattern_stage).
struct PatternSampler {
  const DevSrc* src;
  const DevGrad* sg;
  GradTables T;
  GradRowEval ev;
  bool has_grad, dither;
  float dscale;
  Z2D_D void init(const DevSrc* s, const DevGrad* sg_, const GradTables& T_) {
    src = s;
    sg = sg_;
    T = T_;
    dither = s->kind == Z2D_PARAM_DITHER;
    has_grad = !dither || s->dither_source == Z2D_DITHER_SRC_GRADIENT;
    dscale = dither ? 1.0f / (float)((1 << s->dither_scale) - 1) : 0.0f;
    if (has_grad) ev.init(sg);
  }
  Z2D_D void set_row(int y) {
    if (has_grad) ev.set_row(y);
  }
  // de-multiplied linear colour after dithering (color_vector.zig:90-186); ALPHA_ONLY: only .a is meaningful
  template <bool ALPHA_ONLY>
  Z2D_D RGBAF dithered(int x, int y) const {
    RGBAF c;
    if (has_grad) {
      const StopHit hit = grad_search(*sg, T, ev.offset(x));
      if (ALPHA_ONLY) c = RGBAF{0.f, 0.f, 0.f, hit_alpha(hit)};
      else c = grad_linear(*sg, hit);
    } else {
      c = RGBAF{src->dcol[0], src->dcol[1], src->dcol[2], src->dcol[3]};
    }
    if (src->dither_type == Z2D_DITHER_BAYER || src->dither_type == Z2D_DITHER_BLUE_NOISE) {
      const float m = src->dither_type == Z2D_DITHER_BAYER ? m_bayer(x, y) : m_blue(T, x, y);
      const float ms = m * dscale;
      if (ALPHA_ONLY) c.a = clamp01(c.a + ms);
      else c = RGBAF{clamp01(c.r + ms), clamp01(c.g + ms), clamp01(c.b + ms), clamp01(c.a + ms)};
    }
    return c;
  }
  // integer pipeline: premultiplied RGBA8 (RGBA16Vec.fromGradient / fromDither, compositor.zig:749-800)
  template <bool ALPHA_ONLY>
  Z2D_D RGBA16 sample_int(int x, int y) const {
    if (!dither) {
      const StopHit hit = grad_search(*sg, T, ev.offset(x));
      if (ALPHA_ONLY) return RGBA16{0, 0, 0, round255(hit_alpha(hit))};
      return grad_encode(*sg, hit);
    }
    const RGBAF c = dithered<ALPHA_ONLY>(x, y);
    if (ALPHA_ONLY) return RGBA16{0, 0, 0, round255(c.a)};
    return premul16(encode_raw(c));
  }
  // float pipeline: de-multiplied linear colour (RGBAFloat.Vector.from*, compositor.zig:1086-1131)
  template <bool ALPHA_ONLY>
  Z2D_D RGBAF sample_float(int x, int y) const {
    if (!dither) {
      const StopHit hit = grad_search(*sg, T, ev.offset(x));
      if (ALPHA_ONLY) return RGBAF{0.f, 0.f, 0.f, hit_alpha(hit)};
      return grad_linear(*sg, hit);
    }
    return dithered<ALPHA_ONLY>(x, y);
  }
};

// Stage the gradient of `src` and its stops (<= kGenMaxStops) into shared memory; threads tid, tid + nthreads, ... cooperate.
// Returns the tables to sample with (stop arrays redirected to the staged copies).  The caller synchronises afterwards.
Z2D_D GradTables pattern_stage(const DevSrc& src, const GradTables& T, DevGrad* sg, float* s_off, float4* s_col, int tid, int nthreads) {
  const bool has_grad = src.kind == Z2D_PARAM_GRADIENT || src.dither_source == Z2D_DITHER_SRC_GRADIENT;
  if (has_grad) {
    const DevGrad& g0 = T.grads[src.grad];
    const uint32_t* gw = reinterpret_cast<const uint32_t*>(&g0);
    uint32_t* sw = reinterpret_cast<uint32_t*>(sg);
    for (int k = tid; k < (int)(sizeof(DevGrad) / 4); k += nthreads) sw[k] = (k == 4) ? 0u : gw[k];  // word 4: stop_base -> 0
    for (int k = tid; k < (int)g0.n_stops && k < kGenMaxStops; k += nthreads) {
      s_off[k] = T.stop_offsets[g0.stop_base + k];
      s_col[k] = T.stop_colors[g0.stop_base + k];
    }
  }
  GradTables R = T;
  R.stop_offsets = s_off;
  R.stop_colors = s_col;
  return R;
}

}  // namespace z2d
