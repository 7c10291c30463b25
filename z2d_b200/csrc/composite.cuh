// K5, second half (included by kernels.cu): full-row, single-operator composites that the 32-bit / single-pixel fast kernel
// (k_composite_fast) does not cover.
//
//   k_composite_lut   alpha8 / alpha4 / alpha2 / alpha1 destination, single-pixel source, any operator, either precision.
//                     The new value of a destination BYTE depends only on its old value (1, 2, 4 or 8 packed pixels and a
//                     constant source), so the launch builds a 256-entry byte table with the generic per-pixel code and then
//                     streams the surface through it, 16 bytes per thread and trip: no per-pixel operator dispatch, no
//                     read-modify-write atomics on shared words (surface.zig:787-876 packs LSB-first, rows not byte aligned;
//                     a composite over full rows is one contiguous bit range).
//   k_composite_gen   gradient / dither source, any destination format: every thread owns whole 16-byte chunks of the
//                     destination (4 ... 128 pixels), evaluates the pattern with the row-invariant part of the offset
//                     arithmetic hoisted (same operations in the same order as gradient.zig:349-372, 605-648, 731-741), keeps
//                     the gradient and its stops in shared memory, and for alpha-only destinations computes the source alpha
//                     only (RGBA16 -> alphaN keeps nothing else, compositor.zig:693-696).
#pragma once

namespace z2d {

// =============================================================================== k_composite_lut
__global__ void __launch_bounds__(256) k_composite_lut(const __grid_constant__ CompArgs A) {
  __shared__ uint8_t lut[256];
  const int bits = fmt_bits(A.fmt), per = 8 / bits;
  {
    const uint32_t t = threadIdx.x;
    uint32_t out = 0;
    for (int k = 0; k < per; k++) {
      const uint32_t raw = (t >> (k * bits)) & ((1u << bits) - 1u);
      out |= (comp_pixel(A, raw, 0, 0, 0, 0) & ((1u << bits) - 1u)) << (k * bits);
    }
    lut[t] = (uint8_t)out;
  }
  __syncthreads();
  const uint32_t op = A.ops[0].op;
  const bool noread = op == Z2D_OP_CLEAR || op == Z2D_OP_SRC;  // the table is constant: write only
  const size_t first_px = (size_t)A.dst_start_y * (size_t)A.w, n_px = (size_t)A.scan_w * (size_t)A.rows;
  const size_t bit_lo = first_px * (size_t)bits, bit_hi = bit_lo + n_px * (size_t)bits;
  const size_t B0 = (bit_lo + 7) >> 3, B1 = bit_hi >> 3;  // whole bytes [B0, B1)
  uint8_t* p = A.data;
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, gsz = (size_t)gridDim.x * blockDim.x;
  size_t V0 = B1, V1 = B1;
  if (B0 < B1) {
    const size_t a = (B0 + 15) & ~(size_t)15, b = B1 & ~(size_t)15;
    if (a < b) {
      V0 = a;
      V1 = b;
    }
  }
  // 16-byte chunks [V0, V1)
  {
    uint4* q = reinterpret_cast<uint4*>(p);
    const uint32_t c0 = lut[0];
    const uint32_t fillw = c0 * 0x01010101u;
    for (size_t c = (V0 >> 4) + gid; c < (V1 >> 4); c += gsz) {
      uint4 v;
      if (noread) {
        v = make_uint4(fillw, fillw, fillw, fillw);
      } else {
        v = q[c];
#define Z2D_LUTW(w) \
  w = (uint32_t)lut[w & 255u] | ((uint32_t)lut[(w >> 8) & 255u] << 8) | ((uint32_t)lut[(w >> 16) & 255u] << 16) | ((uint32_t)lut[w >> 24] << 24);
        Z2D_LUTW(v.x) Z2D_LUTW(v.y) Z2D_LUTW(v.z) Z2D_LUTW(v.w)
#undef Z2D_LUTW
      }
      q[c] = v;
    }
  }
  // the whole bytes around the chunks: [B0, V0) and [V1, B1), one thread each
  if (B0 < B1) {
    const size_t head = V0 - B0, tail = B1 - V1;
    if (gid < head + tail) {
      const size_t i = gid < head ? B0 + gid : V1 + (gid - head);
      p[i] = lut[p[i]];
    }
  }
  // partial bytes at either end (the bit range need not start or end on a byte)
  if (gid == 0) {
    const uint32_t lo_bits = (uint32_t)(bit_lo & 7), hi_bits = (uint32_t)(bit_hi & 7);
    const size_t hb = bit_lo >> 3, tb = bit_hi >> 3;
    if (lo_bits && hb == tb) {  // both ends in one byte
      const uint32_t m = ((1u << hi_bits) - 1u) & ~((1u << lo_bits) - 1u);
      const uint32_t old = p[hb];
      p[hb] = (uint8_t)((old & ~m) | (lut[old] & m));
    } else {
      if (lo_bits) {
        const uint32_t m = 0xffu & ~((1u << lo_bits) - 1u);
        const uint32_t old = p[hb];
        p[hb] = (uint8_t)((old & ~m) | (lut[old] & m));
      }
      if (hi_bits) {
        const uint32_t m = (1u << hi_bits) - 1u;
        const uint32_t old = p[tb];
        p[tb] = (uint8_t)((old & ~m) | (lut[old] & m));
      }
    }
  }
}

// =============================================================================== k_composite_gen
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

template <int PREC, bool DITHER, bool ALPHA_ONLY>
struct GenSrc {
  RGBA16 si;
  RGBAF sf;
};

// FC: 0 = 32-bit formats, 1 = alpha8, 2 = alpha4 / alpha2 / alpha1
#ifndef Z2D_GEN_MIN_CTAS
#define Z2D_GEN_MIN_CTAS 4  // 64 registers: 0.67 ms against 0.90 ms at 121 registers / 2 CTAs (rgba, linear gradient, 8192^2)
#endif
template <int FC, int PREC, bool DITHER>
__global__ void __launch_bounds__(256, Z2D_GEN_MIN_CTAS) k_composite_gen(const __grid_constant__ CompArgs A) {
  __shared__ DevGrad sg;
  __shared__ float s_off[kGenMaxStops];
  __shared__ float4 s_col[kGenMaxStops];
  const DevSrc& src = A.ops[0].src;
  const bool has_grad = !DITHER || src.dither_source == Z2D_DITHER_SRC_GRADIENT;
  if (has_grad) {
    const DevGrad& g0 = A.T.grads[src.grad];
    if (threadIdx.x == 0) {
      sg = g0;
      sg.stop_base = 0;
    }
    if (threadIdx.x < g0.n_stops) {
      s_off[threadIdx.x] = A.T.stop_offsets[g0.stop_base + threadIdx.x];
      s_col[threadIdx.x] = A.T.stop_colors[g0.stop_base + threadIdx.x];
    }
  }
  __syncthreads();
  GradTables T = A.T;
  T.stop_offsets = s_off;
  T.stop_colors = s_col;
  GradRowEval ev;
  if (has_grad) ev.init(&sg);
  const uint32_t op = A.ops[0].op, fmt = A.fmt;
  const int bits = FC == 0 ? 32 : FC == 1 ? 8 : fmt_bits(fmt);
  const Fmt32 fd = fmt32_of(fmt);
  const int W = A.w;
  const size_t first_px = (size_t)A.dst_start_y * (size_t)W, end_px = first_px + (size_t)A.scan_w * (size_t)A.rows;
  const int ppc = 128 / bits;  // pixels per 16-byte chunk
  const size_t c_lo = first_px / (size_t)ppc, c_hi = (end_px + (size_t)ppc - 1) / (size_t)ppc;
  uint4* q = reinterpret_cast<uint4*>(A.data);
  const float dscale = DITHER ? 1.0f / (float)((1 << src.dither_scale) - 1) : 0.0f;

  // one source sample: integer pipeline -> premultiplied RGBA8 (alpha only when the destination keeps nothing else),
  // float pipeline -> de-multiplied linear colour (compositor.zig:1086-1131)
  auto sample = [&](int x, int y, RGBA16& si, RGBAF& sf) Z2D_LAMBDA {
    if (!DITHER) {
      const StopHit hit = grad_search(sg, T, ev.offset(x));
      if (PREC == Z2D_PRECISION_INTEGER) {
        if (FC != 0) si = RGBA16{0, 0, 0, round255(hit_alpha(hit))};
        else si = grad_encode(sg, hit);
      } else {
        if (FC != 0) sf = RGBAF{0.f, 0.f, 0.f, hit_alpha(hit)};
        else sf = grad_linear(sg, hit);
      }
      return;
    }
    RGBAF c;
    if (has_grad) {
      const StopHit hit = grad_search(sg, T, ev.offset(x));
      if (FC != 0) c = RGBAF{0.f, 0.f, 0.f, hit_alpha(hit)};
      else c = grad_linear(sg, hit);
    } else {
      c = RGBAF{src.dcol[0], src.dcol[1], src.dcol[2], src.dcol[3]};
    }
    if (src.dither_type == Z2D_DITHER_BAYER || src.dither_type == Z2D_DITHER_BLUE_NOISE) {
      const float m = src.dither_type == Z2D_DITHER_BAYER ? m_bayer(x, y) : m_blue(T, x, y);
      const float ms = m * dscale;
      if (FC != 0) c.a = clamp01(c.a + ms);
      else c = RGBAF{clamp01(c.r + ms), clamp01(c.g + ms), clamp01(c.b + ms), clamp01(c.a + ms)};
    }
    if (PREC == Z2D_PRECISION_INTEGER) {
      if (FC != 0) si = RGBA16{0, 0, 0, round255(c.a)};
      else si = premul16(encode_raw(c));
    } else {
      sf = c;
    }
  };
  // one destination pixel (raw sample in the destination format) -> new raw sample
  auto blend = [&](uint32_t raw, int x, int y) Z2D_LAMBDA -> uint32_t {
    RGBA16 si{0, 0, 0, 0};
    RGBAF sf{0.f, 0.f, 0.f, 0.f};
    sample(x, y, si, sf);
    if (FC == 0) {
      const RGBA16 d = unpack32(fd, raw);
      if (PREC == Z2D_PRECISION_INTEGER) return pack32(fd, int_op_sw(op, d, si));
      return pack32(fd, encode_raw(float_op(op, decode_raw(d), sf)));
    }
    const int da = FC == 1 ? (int)raw : scale_alpha((int)raw, bits, 8);
    int a;
    if (PREC == Z2D_PRECISION_INTEGER) {
      a = int_op_alpha(op, si.a, da);
    } else {
      a = round255(float_op(op, RGBAF{0.f, 0.f, 0.f, (float)da / 255.0f}, sf).a);
    }
    return FC == 1 ? ((uint32_t)a & 255u) : (((uint32_t)a & 255u) >> ((8 - bits) & 7));
  };

  // ONE copy of the per-pixel code (pattern evaluation + 28 operators are tens of KB of instructions: unrolling the pixels of
  // a chunk made the kernel instruction-fetch bound); the chunk's words sit in a small local array.
  const uint32_t smask = bits == 32 ? 0xffffffffu : ((1u << bits) - 1u);
  for (size_t c = c_lo + (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < c_hi; c += (size_t)gridDim.x * blockDim.x) {
    const uint4 v = q[c];
    uint32_t w[4] = {v.x, v.y, v.z, v.w};
    const size_t p0 = c * (size_t)ppc;
    int y = (int)(p0 / (size_t)W), x = (int)(p0 - (size_t)y * (size_t)W);
    y += A.y_origin;  // patterns are evaluated at the canvas row (band destinations)
    if (has_grad) ev.set_row(y);
    const bool interior = p0 >= first_px && p0 + (size_t)ppc <= end_px;
#pragma unroll 1
    for (int k = 0; k < ppc; k++) {
      if (interior || (p0 + (size_t)k >= first_px && p0 + (size_t)k < end_px)) {
        const int wi = (k * bits) >> 5, sh = (k * bits) & 31;
        const uint32_t r = blend((w[wi] >> sh) & smask, x, y);
        w[wi] = (w[wi] & ~(smask << sh)) | ((r & smask) << sh);
      }
      if (++x == W) {
        x = 0;
        ++y;
        if (has_grad) ev.set_row(y);
      }
    }
    q[c] = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

}  // namespace z2d
