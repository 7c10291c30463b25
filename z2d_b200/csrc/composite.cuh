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
#ifndef Z2D_GEN_MIN_CTAS
#define Z2D_GEN_MIN_CTAS 4  // 64 registers: 0.67 ms against 0.90 ms at 121 registers / 2 CTAs (rgba, linear gradient, 8192^2)
#endif
template <int FC, int PREC, bool DITHER>
__global__ void __launch_bounds__(256, Z2D_GEN_MIN_CTAS) k_composite_gen(const __grid_constant__ CompArgs A) {
  __shared__ DevGrad sg;
  __shared__ float s_off[kGenMaxStops];
  __shared__ float4 s_col[kGenMaxStops];
  const DevSrc& src = A.ops[0].src;
  const GradTables T = pattern_stage(src, A.T, &sg, s_off, s_col, (int)threadIdx.x, (int)blockDim.x);
  __syncthreads();
  PatternSampler ps;
  ps.init(&src, &sg, T);
  const bool has_grad = ps.has_grad;
  const uint32_t op = A.ops[0].op, fmt = A.fmt;
  const int bits = FC == 0 ? 32 : FC == 1 ? 8 : fmt_bits(fmt);
  const Fmt32 fd = fmt32_of(fmt);
  const int W = A.w;
  const size_t first_px = (size_t)A.dst_start_y * (size_t)W, end_px = first_px + (size_t)A.scan_w * (size_t)A.rows;
  const int ppc = 128 / bits;  // pixels per 16-byte chunk
  const size_t c_lo = first_px / (size_t)ppc, c_hi = (end_px + (size_t)ppc - 1) / (size_t)ppc;
  uint4* q = reinterpret_cast<uint4*>(A.data);
  (void)DITHER;

  // one destination pixel (raw sample in the destination format) -> new raw sample
  auto blend = [&](uint32_t raw, int x, int y) Z2D_LAMBDA -> uint32_t {
    RGBA16 si{0, 0, 0, 0};
    RGBAF sf{0.f, 0.f, 0.f, 0.f};
    if (PREC == Z2D_PRECISION_INTEGER) si = ps.template sample_int<FC != 0>(x, y);
    else sf = ps.template sample_float<FC != 0>(x, y);
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
    if (has_grad) ps.set_row(y);
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
        if (has_grad) ps.set_row(y);
      }
    }
    q[c] = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

}  // namespace z2d
