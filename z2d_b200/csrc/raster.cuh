// K4: fused coverage + compositing, one warp per 16x16-pixel tile (included by kernels.cu).
//
// For every draw that touches the tile, IN SUBMISSION ORDER:
//   1. classify the draw's edges of this tile-row against the tile from their integer
//      headers (warp-uniform): entirely right -> ignored; entirely left -> a per-sub-scanline
//      scalar winding ("backdrop"); crossing -> evaluated exactly,
//        x_i = round(x_start + x_inc * ((ys + 0.5) - top))          (Polygon.zig:302-307)
//      and accumulated into bit-sliced winding counters (one 64-bit plane per bit, one bit
//      per sample column), so that sample s is inside iff sum(dir | x_i <= s) != 0 / odd
//      (Polygon.zig:326-353);
//   2. popcount the 4x4 samples of each pixel -> coverage 0..16;
//   3. composite the source into the tile (held in shared memory for the whole batch) with
//      the draw's operator (shared.zig, multisample.zig:195-227, supersample.zig:159-184).
// The tile is read from HBM once and written back once per batch.
#pragma once

namespace z2d {

Z2D_D RGBA16 mask_mul16(RGBA16 s, int m) { return {iM(s.r, m), iM(s.g, m), iM(s.b, m), iM(s.a, m)}; }  // dst_in(dst:=s, src:=alpha8 m)

// generic StrideCompositor batch: [dst_in(pattern, mask)]? ; op   (shared.zig:24-45, 78-102), any source / precision
static __device__ __noinline__ uint32_t composite_generic(const DevDraw& d, const GradTables T, uint32_t precision, uint32_t fmt, uint32_t raw,
                                                   int mask8, bool use_mask, int x, int y) {
  if (precision == Z2D_PRECISION_INTEGER) {
    RGBA16 s = src_int(d.src, T, x, y, 0);
    if (use_mask) s = mask_mul16(s, mask8);
    return rgba16_to_raw(fmt, int_op_sw(d.op, raw_to_rgba16(fmt, raw), s));
  }
  RGBAF s = src_float(d.src, T, x, y, 0);
  if (use_mask) {
    const float ma = (float)mask8 / 255.0f;
    s = {s.r * ma, s.g * ma, s.b * ma, ma * s.a};
  }
  RGBAF r = float_op(d.op, decode_raw(raw_to_rgba16(fmt, raw)), s);
  return rgba16_to_raw(fmt, encode_raw(r));
}

// 32-bit formats: channel positions as uniform registers (no per-pixel format switch)
struct Fmt32 {
  int rs, gs, bs;
  uint32_t has_a;
};
Z2D_D Fmt32 fmt32_of(uint32_t fmt) {
  switch (fmt) {
    case Z2D_FMT_ARGB: return {16, 8, 0, 1u};
    case Z2D_FMT_XRGB: return {16, 8, 0, 0u};
    case Z2D_FMT_RGB: return {0, 8, 16, 0u};
    default: return {0, 8, 16, 1u};
  }
}
Z2D_D RGBA16 unpack32(const Fmt32& f, uint32_t raw) {
  return {(int)((raw >> f.rs) & 255u), (int)((raw >> f.gs) & 255u), (int)((raw >> f.bs) & 255u), f.has_a ? (int)(raw >> 24) : 255};
}
Z2D_D uint32_t pack32(const Fmt32& f, RGBA16 v) {
  return (((uint32_t)v.r & 255u) << f.rs) | (((uint32_t)v.g & 255u) << f.gs) | (((uint32_t)v.b & 255u) << f.bs) |
         (f.has_a ? (((uint32_t)v.a & 255u) << 24) : 0u);
}

struct TileFmt {
  uint32_t fmt;
  bool is32;
  Fmt32 f;
};
Z2D_D RGBA16 tf_unpack(const TileFmt& t, uint32_t raw) { return t.is32 ? unpack32(t.f, raw) : raw_to_rgba16(t.fmt, raw); }
Z2D_D uint32_t tf_pack(const TileFmt& t, RGBA16 v) { return t.is32 ? pack32(t.f, v) : rgba16_to_raw(t.fmt, v); }

// cov: number of covered samples (MSAA/SSAA: 0..16, none: 0..1).  Returns the new raw pixel.
Z2D_D uint32_t composite_cov(const DrawHot& h, const DevDraw& d, const GradTables& T, const TileFmt& tf, RGBA16 spx, uint32_t raw, int cov,
                             int x, int y) {
  const uint32_t fmt = tf.fmt;
  int mask;
  bool use_mask;
  if (h.aa == Z2D_AA_SUPERSAMPLE_4X) {
    // mask = box average of 16 samples in the mask surface's own format (supersample.zig:82-91, pixel.zig:435-464,633-646)
    if (fmt == Z2D_FMT_ALPHA4 || fmt == Z2D_FMT_ALPHA2 || fmt == Z2D_FMT_ALPHA1) {
      const int bits = fmt_bits(fmt);
      mask = scale_alpha((((1 << bits) - 1) * cov) / 16, bits, 8);
    } else {
      mask = (255 * cov) / 16;
    }
    if (h.src_kind == Z2D_PARAM_PIXEL && h.precision == Z2D_PRECISION_INTEGER)
      return tf_pack(tf, int_op_sw(h.op, tf_unpack(tf, raw), mask_mul16(spx, mask)));
    return composite_generic(d, T, h.precision, fmt, raw, mask, true, x, y);
  }
  if (cov == 0) return raw;
  if (h.op == Z2D_OP_CLEAR) return 0u;  // shared.zig:18,60 (also at partial coverage)
  if (h.aa == Z2D_AA_NONE || cov == 16) {
    if (h.reduces) return h.paint_raw;  // Surface.paintStride
    mask = 255;
    use_mask = false;
  } else {
    mask = 16 * cov - 1;  // multisample.zig:223
    use_mask = true;
  }
  // the opaque-pixel fast path of the reference is integer-only (surface.zig:557-581)
  const uint32_t prec = h.reduces ? (uint32_t)Z2D_PRECISION_INTEGER : h.precision;
  if (h.src_kind == Z2D_PARAM_PIXEL && prec == Z2D_PRECISION_INTEGER) {
    RGBA16 s = spx;
    if (use_mask) s = mask_mul16(s, mask);
    return tf_pack(tf, int_op_sw(h.op, tf_unpack(tf, raw), s));
  }
  return composite_generic(d, T, prec, fmt, raw, mask, use_mask, x, y);
}

// composite_cov for a gradient / dither source sampled through a PatternSampler (gradient and stops staged in shared memory,
// row-invariant offset arithmetic hoisted): same coverage -> mask rules, same operator arithmetic as composite_generic.
Z2D_D uint32_t composite_cov_pattern(const DrawHot& h, const PatternSampler& ps, const TileFmt& tf, uint32_t raw, int cov, int x, int y) {
  const uint32_t fmt = tf.fmt;
  int mask;
  bool use_mask;
  if (h.aa == Z2D_AA_SUPERSAMPLE_4X) {
    if (fmt == Z2D_FMT_ALPHA4 || fmt == Z2D_FMT_ALPHA2 || fmt == Z2D_FMT_ALPHA1) {
      const int bits = fmt_bits(fmt);
      mask = scale_alpha((((1 << bits) - 1) * cov) / 16, bits, 8);
    } else {
      mask = (255 * cov) / 16;
    }
    use_mask = true;
  } else {
    if (cov == 0) return raw;
    if (h.op == Z2D_OP_CLEAR) return 0u;
    if (h.aa == Z2D_AA_NONE || cov == 16) {
      mask = 255;
      use_mask = false;
    } else {
      mask = 16 * cov - 1;
      use_mask = true;
    }
  }
  if (h.precision == Z2D_PRECISION_INTEGER) {
    RGBA16 s = tf.is32 ? ps.sample_int<false>(x, y) : ps.sample_int<true>(x, y);  // (alpha formats keep nothing but alpha)
    if (use_mask) s = mask_mul16(s, mask);
    return tf_pack(tf, int_op_sw(h.op, tf_unpack(tf, raw), s));
  }
  RGBAF s = tf.is32 ? ps.sample_float<false>(x, y) : ps.sample_float<true>(x, y);
  if (use_mask) {
    const float ma = (float)mask / 255.0f;
    s = {s.r * ma, s.g * ma, s.b * ma, ma * s.a};
  }
  return tf_pack(tf, encode_raw(float_op(h.op, decode_raw(tf_unpack(tf, raw)), s)));
}

// ------------------------------------------------------------------------- coverage
template <int W>
Z2D_D void wind_add(uint64_t (&p)[W], uint64_t mask, bool up) {
  uint64_t c = mask;
  if (up) {  // +1 on every column >= col
#pragma unroll
    for (int k = 0; k < W; k++) {
      const uint64_t t = p[k] & c;
      p[k] ^= c;
      c = t;
    }
  } else {  // -1
#pragma unroll
    for (int k = 0; k < W; k++) {
      const uint64_t t = ~p[k] & c;
      p[k] ^= c;
      c = t;
    }
  }
}

Z2D_D double4 ld_edge(const DevEdge* e) {  // 2 x 128-bit read-only loads (warp-uniform address)
  const double2* q = reinterpret_cast<const double2*>(e);
  const double2 a = __ldg(q), b = __ldg(q + 1);
  return make_double4(a.x, a.y, b.x, b.y);
}

// Column (relative to the tile's first sample) at which an ACTIVE edge crosses the sub-scanline
// centre ys + 0.5, or -1 when that is right of the tile.
Z2D_D int edge_col(const double4& ev, double top, int ys, int sx0, int ncols) {
  const double xf = round_half_away(ev.z + (ev.w * (((double)ys + 0.5) - top)));  // Polygon.zig:305
  const double cf = xf - (double)sx0;
  if (!(cf < (double)ncols)) return -1;
  return cf < 0.0 ? 0 : (int)cf;
}

// header helpers: {xlo, xhi, first active row | dir << 31, last active row} (k_bin_scatter)
Z2D_D bool hdr_active(const int4& h, int ys) { return ys >= (h.z & 0x7fffffff) && ys <= h.w; }  // == top < ys + 0.5 <= bottom

constexpr uint32_t kCrossListCap = 128;  // crossing-edge positions kept per (draw, tile) by tile_cover

template <int W, bool TWO>
Z2D_D void cross_pass(const DevEdge* __restrict__ be, const int4* __restrict__ hd, uint32_t n_be, const uint16_t* __restrict__ clist,
                      uint32_t ncross, int ys0, int sx0, int ncols, bool even_odd, int wl0, int wl1, uint64_t& m0, uint64_t& m1,
                      uint32_t& n_eval) {
  uint64_t p0[W], p1[TWO ? W : 1];
#pragma unroll
  for (int k = 0; k < W; k++) {  // start from the backdrop winding (two's complement, bit-sliced)
    p0[k] = ((wl0 >> k) & 1) ? ~0ull : 0ull;
    if (TWO) p1[k] = ((wl1 >> k) & 1) ? ~0ull : 0ull;
  }
  const int sx_hi = sx0 + ncols;
  auto apply = [&](const double4& ev, bool up, bool a0, bool a1) Z2D_LAMBDA {
    const double top = up ? ev.y : ev.x;
    n_eval += (uint32_t)a0 + (uint32_t)(TWO && a1);
    if (a0) {
      const int c0 = edge_col(ev, top, ys0, sx0, ncols);
      if (c0 >= 0) {
        const uint64_t mask = ~0ull << c0;
        if (even_odd) p0[0] ^= mask; else wind_add<W>(p0, mask, up);
      }
    }
    if (TWO && a1) {
      const int c1 = edge_col(ev, top, ys0 + 1, sx0, ncols);
      if (c1 >= 0) {
        const uint64_t mask = ~0ull << c1;
        if (even_odd) p1[0] ^= mask; else wind_add<TWO ? W : 1>(p1, mask, up);
      }
    }
  };
  // `clist`: positions (in the band) of the edges that cross the tile, compacted by tile_cover's classification pass -- a text
  // run bins hundreds of edges per tile row of which a tile sees a few dozen
  if (ncross <= 64) {
    // Which of the crossing edges are live on THIS lane's rows (warp-uniform loop, integer tests only) ...
    uint64_t my0 = 0, my1 = 0, up_bits = 0;
    for (uint32_t k = 0; k < ncross; k++) {
      const int4 h = __ldg(hd + clist[k]);
      const uint64_t bit = 1ull << k;
      if (hdr_active(h, ys0)) my0 |= bit;
      if (TWO && hdr_active(h, ys0 + 1)) my1 |= bit;
      if (h.z < 0) up_bits |= bit;
    }
    // ... then every lane walks only its own edges: the warp iterates max-over-rows(edges per row) times, not once per
    // crossing edge of the tile, and no lane idles on an edge that does not reach its rows.
    for (uint64_t mine = my0 | my1; mine; mine &= mine - 1) {
      const int k = __ffsll((long long)mine) - 1;
      const uint64_t bit = 1ull << k;
      apply(ld_edge(be + clist[k]), (up_bits & bit) != 0, (my0 & bit) != 0, (my1 & bit) != 0);
    }
  } else if (ncross <= kCrossListCap) {
    for (uint32_t k = 0; k < ncross; k++) {
      const uint32_t i = clist[k];
      const int4 h = __ldg(hd + i);
      const bool a0 = hdr_active(h, ys0), a1 = TWO && hdr_active(h, ys0 + 1);
      if (a0 || a1) apply(ld_edge(be + i), h.z < 0, a0, a1);
    }
  } else {  // (the list overflowed: every edge of the band)
    for (uint32_t i = 0; i < n_be; i++) {
      const int4 h = __ldg(hd + i);
      const bool a0 = hdr_active(h, ys0), a1 = TWO && hdr_active(h, ys0 + 1);
      if ((a0 || a1) && h.x <= sx_hi && h.y >= sx0) apply(ld_edge(be + i), h.z < 0, a0, a1);
    }
  }
  if (even_odd) {
    m0 = p0[0];
    if (TWO) m1 = p1[0];
  } else {
    uint64_t a = 0, b = 0;
#pragma unroll
    for (int k = 0; k < W; k++) {
      a |= p0[k];
      if (TWO) b |= p1[k];
    }
    m0 = a;
    if (TWO) m1 = b;
  }
}

// rare: more than 60 edges of one draw cross one tile; 32 planes in local memory, one row at a time
static __device__ __noinline__ uint64_t cross_row_wide(const DevEdge* __restrict__ be, const int4* __restrict__ hd, uint32_t n_be,
                                                const uint16_t* __restrict__ clist, uint32_t ncross, int ys, int sx0, int ncols, bool even_odd,
                                                int wl) {
  uint64_t p[32];
  for (int k = 0; k < 32; k++) p[k] = ((wl >> k) & 1) ? ~0ull : 0ull;
  const int sx_hi = sx0 + ncols;
  const bool listed = ncross <= kCrossListCap;
  const uint32_t n_it = listed ? ncross : n_be;
  for (uint32_t it = 0; it < n_it; it++) {
    const uint32_t i = listed ? (uint32_t)clist[it] : it;
    const int4 h = __ldg(hd + i);
    if (h.x > sx_hi || h.y < sx0 || !hdr_active(h, ys)) continue;
    const double4 ev = ld_edge(be + i);
    const bool up = h.z < 0;
    const int col = edge_col(ev, up ? ev.y : ev.x, ys, sx0, ncols);
    if (col < 0) continue;
    uint64_t cy = ~0ull << col;
    if (even_odd) {
      p[0] ^= cy;
    } else if (up) {
      for (int k = 0; k < 32 && cy; k++) { const uint64_t t = p[k] & cy; p[k] ^= cy; cy = t; }
    } else {
      for (int k = 0; k < 32 && cy; k++) { const uint64_t t = ~p[k] & cy; p[k] ^= cy; cy = t; }
    }
  }
  if (even_odd) return p[0];
  uint64_t a = 0;
  for (int k = 0; k < 32; k++) a |= p[k];
  return a;
}

// Inside-masks of this lane's sub-scanlines ys0 (and ys0+1 when `two`) for one draw in one tile.  Called by the whole warp
// with warp-uniform be / hd / n_be; `wdiff` is the warp's 65-entry scratch in shared memory and ys_tile0 the tile's
// first (sub-)scanline.
Z2D_D void tile_cover(const DevEdge* __restrict__ be, const int4* __restrict__ hd, uint32_t n_be, int ys_tile0, int ys0, bool two, int sx0,
                      int ncols, uint32_t rule, int* __restrict__ wdiff, uint64_t& m0, uint64_t& m1, uint32_t& n_eval) {
  const bool even_odd = rule == Z2D_FILL_EVEN_ODD;
  const int lane = (int)(threadIdx.x & 31u);
  const int sx_hi = sx0 + ncols;
  const int nrows = two ? 64 : 16;
  // pass 1, lanes stride over the edges: an edge entirely left of the tile adds its direction to the backdrop winding of
  // the rows it is active on (a +dir / -dir pair in a per-row difference array); crossing edges are collected in a bit mask
  wdiff[lane] = 0;
  wdiff[lane + 32] = 0;
  if (lane == 0) wdiff[64] = 0;
  uint16_t* clist = reinterpret_cast<uint16_t*>(wdiff + 68);  // kCrossListCap entries behind the 65 difference slots (528-byte scratch)
  __syncwarp();
  uint32_t ncross = 0;
  for (uint32_t base = 0; base < n_be; base += 32) {
    const uint32_t i = base + (uint32_t)lane;
    bool cross = false;
    if (i < n_be) {
      const int4 h = __ldg(hd + i);
      if (h.x <= sx_hi) {      // else entirely right of the tile
        if (h.y < sx0) {       // entirely left: only its winding matters
          const int r0 = max((h.z & 0x7fffffff) - ys_tile0, 0), r1 = min(h.w - ys_tile0, nrows - 1);
          if (r0 <= r1) {
            const int dir = h.z < 0 ? 1 : -1;
            atomicAdd(&wdiff[r0], dir);
            atomicAdd(&wdiff[r1 + 1], -dir);
          }
        } else {
          cross = true;
        }
      }
    }
    const uint32_t b = __ballot_sync(0xffffffffu, cross);
    if (cross) {
      const uint32_t k = ncross + (uint32_t)__popc(b & ((1u << lane) - 1u));
      if (k < kCrossListCap && i < 65536u) clist[k] = (uint16_t)i;
    }
    ncross += (uint32_t)__popc(b);
  }
  if (n_be > 65536u && ncross <= kCrossListCap) ncross = kCrossListCap + 1;  // (positions would not fit 16 bits: full loops)
  __syncwarp();
  int wl0, wl1 = 0;
  {  // prefix sum of the difference array -> backdrop winding of this lane's rows
    const int a = two ? wdiff[2 * lane] : (lane < 16 ? wdiff[lane] : 0);
    const int b = two ? wdiff[2 * lane + 1] : 0;
    int incl = a + b;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (two) {
      wl0 = incl - b;
      wl1 = incl;
    } else {
      wl0 = __shfl_sync(0xffffffffu, incl, lane >> 1);
    }
  }
  __syncwarp();  // wdiff is rewritten by the next call
  if (ncross == 0) {
    m0 = (even_odd ? (wl0 & 1) : (wl0 != 0)) ? ~0ull : 0ull;
    m1 = (even_odd ? (wl1 & 1) : (wl1 != 0)) ? ~0ull : 0ull;
    return;
  }
  // |sum over crossing edges| <= ncross: a backdrop beyond that keeps the whole row inside (non-zero rule)
  const bool full0 = !even_odd && (wl0 > (int)ncross || -wl0 > (int)ncross);
  const bool full1 = !even_odd && (wl1 > (int)ncross || -wl1 > (int)ncross);
  const int b0 = full0 ? 0 : wl0, b1 = full1 ? 0 : wl1;
  if (ncross <= 7) {
    if (two) cross_pass<5, true>(be, hd, n_be, clist, ncross, ys0, sx0, ncols, even_odd, b0, b1, m0, m1, n_eval);
    else cross_pass<5, false>(be, hd, n_be, clist, ncross, ys0, sx0, ncols, even_odd, b0, b1, m0, m1, n_eval);
  } else if (ncross <= 60) {
    uint64_t dummy = 0;
    cross_pass<8, false>(be, hd, n_be, clist, ncross, ys0, sx0, ncols, even_odd, b0, 0, m0, dummy, n_eval);
    if (two) cross_pass<8, false>(be, hd, n_be, clist, ncross, ys0 + 1, sx0, ncols, even_odd, b1, 0, m1, dummy, n_eval);
  } else {
    m0 = cross_row_wide(be, hd, n_be, clist, ncross, ys0, sx0, ncols, even_odd, b0);
    m1 = two ? cross_row_wide(be, hd, n_be, clist, ncross, ys0 + 1, sx0, ncols, even_odd, b1) : 0ull;
  }
  if (full0) m0 = ~0ull;
  if (full1) m1 = ~0ull;
}

// ---- common case: at most 32 edges of the draw in this tile-row, so lane j holds the header of edge j.
// Both per-row quantities the crossing pass needs come from ONE shared-memory difference array and one warp scan:
//   .x  winding difference: +dir at the first row an edge LEFT of the tile is active on, -dir after its last row;
//       the prefix sum is the backdrop winding of every row;
//   .y  xor difference: bit j toggled at the first row and after the last row of CROSSING edge j; the prefix xor is,
//       per row, the set of crossing edges active on it -- every lane then walks only the edges of its own rows.
Z2D_D double4 lds_edge(const uint4* __restrict__ es, int i) {  // edge i of the pair from shared memory ({y0, y1} | {x_start, x_inc})
  const uint4 a = es[i], b = es[32 + i];
  return make_double4(__hiloint2double((int)a.y, (int)a.x), __hiloint2double((int)a.w, (int)a.z), __hiloint2double((int)b.y, (int)b.x),
                      __hiloint2double((int)b.w, (int)b.z));
}

template <int W, bool TWO>
Z2D_D void cross_pass32(const uint4* __restrict__ es, uint32_t my0, uint32_t my1, uint32_t up_b, int ys0, int sx0, int ncols, bool even_odd,
                        int wl0, int wl1, uint64_t& m0, uint64_t& m1, uint32_t& n_eval) {
  uint64_t p0[W], p1[TWO ? W : 1];
#pragma unroll
  for (int k = 0; k < W; k++) {  // start from the backdrop winding (two's complement, bit-sliced)
    p0[k] = (uint64_t)(int64_t)((int32_t)((uint32_t)wl0 << (31 - k)) >> 31);
    if (TWO) p1[k] = (uint64_t)(int64_t)((int32_t)((uint32_t)wl1 << (31 - k)) >> 31);
  }
  for (uint32_t mine = my0 | (TWO ? my1 : 0u); mine; mine &= mine - 1) {
    const int i = __ffs((int)mine) - 1;
    const uint32_t bit = 1u << i;
    const double4 ev = lds_edge(es, i);
    const bool up = (up_b & bit) != 0u, a0 = (my0 & bit) != 0u, a1 = TWO && (my1 & bit) != 0u;
    const double top = up ? ev.y : ev.x;
    n_eval += (uint32_t)a0 + (uint32_t)a1;
    if (a0) {
      const int c0 = edge_col(ev, top, ys0, sx0, ncols);
      if (c0 >= 0) {
        const uint64_t mask = ~0ull << c0;
        if (even_odd) p0[0] ^= mask; else wind_add<W>(p0, mask, up);
      }
    }
    if (TWO && a1) {
      const int c1 = edge_col(ev, top, ys0 + 1, sx0, ncols);
      if (c1 >= 0) {
        const uint64_t mask = ~0ull << c1;
        if (even_odd) p1[0] ^= mask; else wind_add<TWO ? W : 1>(p1, mask, up);
      }
    }
  }
  uint64_t a = 0, b = 0;
#pragma unroll
  for (int k = 0; k < W; k++) {  // (even-odd is instantiated with W = 1: plane 0 is the parity)
    a |= p0[k];
    if (TWO) b |= p1[k];
  }
  m0 = a;
  if (TWO) m1 = b;
}

// Second half of tile_cover32 / tile_cover_big: the difference array is complete (every lane's atomics issued), `cross` / `up`
// say whether this lane's slot holds a crossing edge and its direction.
template <bool TWO>
Z2D_D void tile_cover_tail(const uint4* __restrict__ es, uint2* __restrict__ diff, bool cross, bool up, int ys0, int sx0, bool even_odd,
                           uint64_t& m0, uint64_t& m1, uint32_t& n_eval) {
  constexpr int ncols = TWO ? 64 : 16;
  const int lane = (int)(threadIdx.x & 31u);
  const uint32_t cross_b = __ballot_sync(0xffffffffu, cross), up_b = __ballot_sync(0xffffffffu, up);
  __syncwarp();
  int wl0, wl1 = 0;
  uint32_t my0, my1 = 0u;
  {
    uint4 d4 = make_uint4(0u, 0u, 0u, 0u);  // {winding diff, xor diff} of rows 2 * lane, 2 * lane + 1 (16 rows: of row `lane`)
    if (TWO) {
      d4 = reinterpret_cast<const uint4*>(diff)[lane];
    } else if (lane < 16) {
      const uint2 d2 = diff[lane];
      d4.x = d2.x;
      d4.y = d2.y;
    }
    int incl = (int)d4.x + (int)d4.z;
    uint32_t inclx = d4.y ^ d4.w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      const uint32_t tx = __shfl_up_sync(0xffffffffu, inclx, o);
      if (lane >= o) {
        incl += t;
        inclx ^= tx;
      }
    }
    if (TWO) {
      wl0 = incl - (int)d4.z;
      wl1 = incl;
      my0 = inclx ^ d4.w;
      my1 = inclx;
    } else {
      wl0 = __shfl_sync(0xffffffffu, incl, lane >> 1);
      my0 = __shfl_sync(0xffffffffu, inclx, lane >> 1);
    }
  }
  __syncwarp();  // diff is rewritten by the next call
  if (cross_b == 0u) {
    m0 = (even_odd ? (wl0 & 1) : (wl0 != 0)) ? ~0ull : 0ull;
    m1 = (even_odd ? (wl1 & 1) : (wl1 != 0)) ? ~0ull : 0ull;
    return;
  }
  if (even_odd) {
    cross_pass32<1, TWO>(es, my0, my1, up_b, ys0, sx0, ncols, true, wl0, wl1, m0, m1, n_eval);
    return;
  }
  // the crossing edges of a row move its winding by at most their number: a backdrop beyond that keeps the whole row inside
  const int k0 = __popc(my0), k1 = __popc(my1);
  const bool full0 = abs(wl0) > k0, full1 = TWO && abs(wl1) > k1;
  const int b0 = full0 ? 0 : wl0, b1 = full1 ? 0 : wl1;
  const int bound = max(full0 ? 0 : abs(wl0) + k0, full1 ? 0 : abs(wl1) + k1);  // largest |winding| any of this lane's rows can reach
  const int bmax = __reduce_max_sync(0xffffffffu, bound);
  if (TWO && bmax <= 3) {
    cross_pass32<3, TWO>(es, my0, my1, up_b, ys0, sx0, ncols, false, b0, b1, m0, m1, n_eval);
  } else if (bmax <= 15) {  // one row at a time: 5 planes x 2 rows would not fit the register budget of the kernel
    uint64_t dummy = 0;
    cross_pass32<5, false>(es, my0, 0u, up_b, ys0, sx0, ncols, false, b0, 0, m0, dummy, n_eval);
    if (TWO) cross_pass32<5, false>(es, my1, 0u, up_b, ys0 + 1, sx0, ncols, false, b1, 0, m1, dummy, n_eval);
  } else {  // (at most 32 crossing edges here: 8 planes hold +-64)
    uint64_t dummy = 0;
    cross_pass32<8, false>(es, my0, 0u, up_b, ys0, sx0, ncols, false, b0, 0, m0, dummy, n_eval);
    if (TWO) cross_pass32<8, false>(es, my1, 0u, up_b, ys0 + 1, sx0, ncols, false, b1, 0, m1, dummy, n_eval);
  }
  if (full0) m0 = ~0ull;
  if (full1) m1 = ~0ull;
}

// BIG = false: n_be <= 32, lane j holds edge j.
// BIG = true (RICH kernel): also n_be > 32 -- strokes with round joins and caps bin dozens of short edges per tile row, most of
// them elsewhere along x: lanes stride over the headers, edges left of the tile go straight into the winding differences and
// the crossing ones are compacted into the 32 shared-memory slots, so that the rest runs exactly as in the small case (ONE copy
// of the tail: a second one cost the stroke workload 30 %).  Returns false only when more than 32 edges cross the tile.
template <bool TWO, bool BIG>
Z2D_D bool tile_cover32(const DevEdge* __restrict__ be, const int4* __restrict__ hd, uint32_t n_be, int ys_tile0, int ys0, int sx0,
                        uint32_t rule, uint2* __restrict__ diff, uint4* __restrict__ es, uint32_t* __restrict__ meta, uint64_t& m0, uint64_t& m1,
                        uint32_t& n_eval) {
  constexpr int ncols = TWO ? 64 : 16, nrows = TWO ? 64 : 16;
  const int lane = (int)(threadIdx.x & 31u);
  const int sx_hi = sx0 + ncols;
  reinterpret_cast<uint4*>(diff)[lane] = make_uint4(0u, 0u, 0u, 0u);  // entries 0 .. 65
  if (lane == 0) reinterpret_cast<uint4*>(diff)[32] = make_uint4(0u, 0u, 0u, 0u);
  bool cross = false, up = false;
  __syncwarp();
  if (!BIG || n_be <= 32u) {
    if ((uint32_t)lane < n_be) {
      const int4 h = __ldg(hd + lane);
      {  // the edge itself goes to shared memory now (coalesced, independent of the header): the crossing pass gathers from there
        const uint4* g = reinterpret_cast<const uint4*>(be + lane);
        const uint4 e0 = __ldg(g), e1 = __ldg(g + 1);
        es[lane] = e0;
        es[32 + lane] = e1;
      }
      const int r0 = max((h.z & 0x7fffffff) - ys_tile0, 0), r1 = min(h.w - ys_tile0, nrows - 1);
      up = h.z < 0;
      if (h.x <= sx_hi && r0 <= r1) {  // else entirely right of the tile, or not active on its rows
        if (h.y < sx0) {               // entirely left: only its winding matters
          const int dir = up ? 1 : -1;
          atomicAdd(reinterpret_cast<int*>(&diff[r0].x), dir);
          atomicAdd(reinterpret_cast<int*>(&diff[r1 + 1].x), -dir);
        } else {
          cross = true;
          atomicXor(&diff[r0].y, 1u << lane);
          atomicXor(&diff[r1 + 1].y, 1u << lane);
        }
      }
    }
  } else {
    uint32_t n_slots = 0;
    for (uint32_t base = 0; base < n_be; base += 32) {
      const uint32_t i = base + (uint32_t)lane;
      bool c = false, u = false;
      int r0 = 0, r1 = 0;
      if (i < n_be) {
        const int4 h = __ldg(hd + i);
        r0 = max((h.z & 0x7fffffff) - ys_tile0, 0);
        r1 = min(h.w - ys_tile0, nrows - 1);
        u = h.z < 0;
        if (h.x <= sx_hi && r0 <= r1) {
          if (h.y < sx0) {
            const int dir = u ? 1 : -1;
            atomicAdd(reinterpret_cast<int*>(&diff[r0].x), dir);
            atomicAdd(reinterpret_cast<int*>(&diff[r1 + 1].x), -dir);
          } else {
            c = true;
          }
        }
      }
      const uint32_t cb = __ballot_sync(0xffffffffu, c);
      if (c) {
        const uint32_t slot = n_slots + (uint32_t)__popc(cb & ((1u << lane) - 1u));
        if (slot < 32u) {
          const uint4* g = reinterpret_cast<const uint4*>(be + i);
          es[slot] = __ldg(g);
          es[32 + slot] = __ldg(g + 1);
          meta[slot] = (uint32_t)r0 | ((uint32_t)(r1 + 1) << 8) | (u ? 0x10000u : 0u);
        }
      }
      n_slots += (uint32_t)__popc(cb);
    }
    if (n_slots > 32u) return false;
    __syncwarp();
    if ((uint32_t)lane < n_slots) {
      const uint32_t mt = meta[lane];
      cross = true;
      up = (mt & 0x10000u) != 0u;
      atomicXor(&diff[mt & 0xffu].y, 1u << lane);
      atomicXor(&diff[(mt >> 8) & 0xffu].y, 1u << lane);
    }
  }
  tile_cover_tail<TWO>(es, diff, cross, up, ys0, sx0, rule == Z2D_FILL_EVEN_ODD, m0, m1, n_eval);
  return true;
}

// Draws flagged kDrawUnpaired only.  The reference pairs the sorted (filtered) crossings of a scanline and drops a
// trailing unmatched one ("for (0..filtered_edge_set.len / 2)", multisample.zig:156 / supersample.zig / direct.zig), so
// the inside run that would extend to +infinity is not drawn.  Which crossing that is depends on the order the
// reference's sort leaves equal crossings in; k_edge_sim (kernels.cu) replays the reference's scanline loop for these
// draws and records the x of the dropped crossing per (sub-)scanline: clear every sample at or right of it.
Z2D_D uint64_t cut_open_tail(const RasterArgs& A, const DrawHot& h, int ys, int sx0, int ncols, uint64_t m) {
  const int r = ys - h.sim_y0;
  if (r < 0 || r >= h.sim_rows) return m;
  const int cut = __ldg(&A.sim_rows[h.sim_base + (uint32_t)r]).x;
  if (cut == INT_MAX) return m;
  const long long cf = (long long)cut - (long long)sx0;
  if (cf >= (long long)ncols) return m;
  if (cf <= 0) return 0ull;
  return m & ~(~0ull << (int)cf);
}

Z2D_D uint32_t nibble_popc(uint32_t x) {  // per-nibble popcount (values 0..4 in each 4-bit field)
  x = x - ((x >> 1) & 0x55555555u);
  return (x & 0x33333333u) + ((x >> 2) & 0x33333333u);
}

// ---- packed integer src_over for the 32-bit formats (two 16-bit lanes per register: bytes 0,2 and bytes 1,3;
// alpha / padding is byte 3 in all four layouts).  floor(p / 255) == (p + 1 + (p >> 8)) >> 8 for 0 <= p <= 65534.
Z2D_D uint32_t div255_x2(uint32_t p) { return ((p + 0x00010001u + ((p >> 8) & 0x00ff00ffu)) >> 8) & 0x00ff00ffu; }
// source pixel at mask m, in the surface's channel layout: mask_mul16 == dst_in(src, alpha8 m) (surface.zig:573-576)
Z2D_D uint2 src_lanes(const Fmt32& f, RGBA16 s, int m, bool masked) {
  const uint32_t c0 = f.rs == 0 ? (uint32_t)s.r : (uint32_t)s.b, c2 = f.rs == 0 ? (uint32_t)s.b : (uint32_t)s.r;
  uint32_t lo = c0 | (c2 << 16), hi = (uint32_t)s.g | ((uint32_t)s.a << 16);
  if (masked) {
    lo = div255_x2(lo * (uint32_t)m);
    hi = div255_x2(hi * (uint32_t)m);
  }
  return make_uint2(lo, hi);
}
// IntegerOps.src_over: colour = sc + dc * (255 - sa) / 255, alpha = sa + da - sa * da / 255  (compositor.zig:1216-1231)
// The alpha lane uses sa * da / 255 == da - ceil(da * (255 - sa) / 255), i.e. sa + (da * inv + 254) / 255.
Z2D_D uint32_t src_over_x4(uint32_t raw, uint2 s, uint32_t amask) {
  const uint32_t inv = 255u - (s.y >> 16);
  const uint32_t dlo = raw & 0x00ff00ffu, dhi = ((raw | ~amask) >> 8) & 0x00ff00ffu;  // no alpha channel: da = 255
  const uint32_t lo = s.x + div255_x2(dlo * inv);
  const uint32_t hi = s.y + div255_x2(dhi * inv + (254u << 16));
  return (lo | (hi << 8)) & amask;
}

// ---- table-driven form of the same blend for the tile kernel.  Entry c of the per-(draw, tile) table holds the source at
// coverage level c in blend-ready form: {C_lo, C_hi, 255 - sa}, C = (source lanes << 8) + the rounding terms of div255_x2
// so that per pixel, with p = d * inv (+ 254 << 16 in the alpha lane, the IMAD's addend),
//     q = p + ((p >> 8) & 0x00ff00ff) + C                      (one IMAD, one PRMT, one IADD3 per 16-bit lane pair)
// and byte 1 / byte 3 of q are the blended channels: s + floor(d * inv / 255) <= 255 for premultiplied sources, so the sum
// never carries into the neighbouring lane.  Identical results to src_over_x4.  Formats without an alpha channel keep the
// destination alpha lane at 0 (hmask) and C_hi's alpha lane at 1: a touched pixel gets padding byte 0, as pack32 writes it.
Z2D_D uint4 blend_entry(const Fmt32& f, RGBA16 s, int m, bool masked) {
  const uint2 sl = src_lanes(f, s, m, masked);
  const uint32_t inv = 255u - (sl.y >> 16);
  const uint32_t hi = f.has_a ? (0x00010001u + (sl.y << 8)) : (0x00010001u + ((sl.y & 0xffu) << 8));
  return make_uint4(0x00010001u + (sl.x << 8), hi, inv, f.has_a ? (254u << 16) : 0u);  // .w: addend of the alpha lane's product
}
Z2D_D uint32_t blend_tab_px(uint32_t raw, const uint4 e, uint32_t hmask) {
  const uint32_t plo = (raw & 0x00ff00ffu) * e.z, phi = ((raw >> 8) & hmask) * e.z + e.w;
  const uint32_t qlo = plo + __byte_perm(plo, 0u, 0x4341) + e.x;  // + ((p >> 8) & 0x00ff00ff)
  const uint32_t qhi = phi + __byte_perm(phi, 0u, 0x4341) + e.y;
  return __byte_perm(qlo, qhi, 0x7351);  // bytes: qlo.1, qhi.1, qlo.3, qhi.3
}
// the lane's 8 pixels (two uint4); cov_e / cov_o: coverage level of the even / odd pixels, one byte each (0: untouched)
Z2D_D void blend_fast8(uint4& v0, uint4& v1, uint32_t cov_e, uint32_t cov_o, const uint4* __restrict__ tab, uint32_t hmask) {
#define Z2D_BL(dst, covw, sel)                                                   \
  {                                                                              \
    const uint32_t c = __byte_perm(covw, 0u, sel);                               \
    const uint32_t r = blend_tab_px(dst, tab[c], hmask);                         \
    if (c) dst = r;                                                              \
  }
  Z2D_BL(v0.x, cov_e, 0x4440) Z2D_BL(v0.y, cov_o, 0x4440) Z2D_BL(v0.z, cov_e, 0x4441) Z2D_BL(v0.w, cov_o, 0x4441)
  Z2D_BL(v1.x, cov_e, 0x4442) Z2D_BL(v1.y, cov_o, 0x4442) Z2D_BL(v1.z, cov_e, 0x4443) Z2D_BL(v1.w, cov_o, 0x4443)
#undef Z2D_BL
}
Z2D_D uint32_t nonzero_bytes(uint32_t x) { return (uint32_t)__popc(((x & 0x7f7f7f7fu) + 0x7f7f7f7fu | x) & 0x80808080u); }

// Gradient / dither source on one (draw, tile) pair: a real function (values and plain pointers only), so that its registers --
// the f64 row terms of the offset arithmetic, the 28-operator float pipeline -- do not weigh on the allocation of the kernel's hot
// path.  Stages the gradient, its stops and the source record in shared memory once per pair and samples with the row-invariant
// part of the offset arithmetic hoisted (a lane's 8 pixels share a row).  Returns the number of pixels composited by this lane.
// flags: 1 unbounded MSAA pre-clear, 2 every pixel of the region is composited (supersample), 4 the lane's row is on the surface.
static __device__ __noinline__ uint32_t pattern_tile(const DevDraw* dp, const DrawHot* hp, GradTables T, uint32_t* px, DevGrad* sg, float* soff,
                                              float4* scol, DevSrc* psrc, uint32_t cov_e, uint32_t cov_o, int px0, int py, int sfc_w,
                                              uint32_t fmt, uint32_t flags) {
  const int lane = (int)(threadIdx.x & 31u);
  const DrawHot& h = *hp;
  const bool pre = (flags & 1u) != 0u, all_px = (flags & 2u) != 0u, row_ok = (flags & 4u) != 0u;
  TileFmt tf;
  tf.fmt = fmt;
  tf.is32 = fmt <= Z2D_FMT_RGBA;
  tf.f = fmt32_of(fmt);
  {
    const uint32_t* gw = reinterpret_cast<const uint32_t*>(&dp->src);
    uint32_t* sw = reinterpret_cast<uint32_t*>(psrc);
    for (int k = lane; k < (int)(sizeof(DevSrc) / 4); k += 32) sw[k] = gw[k];
  }
  const GradTables Tp = pattern_stage(dp->src, T, sg, soff, scol, lane, 32);
  __syncwarp();
  PatternSampler ps;
  ps.init(psrc, sg, Tp);
  ps.set_row(py);
  uint32_t n_cov = 0;
#pragma unroll 1
  for (int i = 0; i < 8; i++) {
    const int x = px0 + i;
    const int cov = (int)(((i & 1) ? cov_o : cov_e) >> (8 * (i >> 1))) & 0xff;
    const bool in_sfc = x < sfc_w && row_ok;
    const bool in_reg = in_sfc && x >= h.rx0 && x < h.rx1 && py >= h.ry0 && py < h.ry1;
    if (!(in_sfc && (pre || (in_reg && (all_px || cov != 0))))) continue;
    const int pi = (((i) >> 2) << 7) + (lane << 2) + ((i) & 3);
    uint32_t raw = px[pi];
    if (pre) {  // multisample.zig:96-110
      if (py < h.pre_y0 || (py > h.pre_y1 && py < h.pre_rows) || (py >= h.pre_y0 && py <= h.pre_y1 && x < h.pre_x)) raw = 0u;
    }
    if (in_reg) {
      n_cov += cov > 0;
      raw = composite_cov_pattern(h, ps, tf, raw, cov, x, py);
    }
    px[pi] = raw;
  }
  __syncwarp();
  return n_cov;
}

// One warp per tile.  The default is ONE WARP PER CTA (32 threads): everything derived from blockIdx is then uniform for the
// compiler (uniform registers and datapath instead of one copy per lane), which measured 3.03 ms against 3.34 ms (128 threads)
// and 3.9 ms (256 threads) on the 100 k-path scene.
// Two instantiations, because the hot path is sensitive to everything that shares its register allocation:
//   LEAN  batches of fills with single-pixel sources (config 2): 64 registers, 32 CTAs per SM;
//   RICH  batches with strokes or gradient / dither sources: adds the compaction of crossing edges for tile rows with more
//         than 32 binned edges (tile_cover_big) and the staged pattern path (pattern_tile); 80 registers, 24 CTAs per SM.
//   Same box, raster ms (config 2 / config 3 / 256 config-5 scenes): LEAN-only code 2.37 / 10.8 / 10.4, RICH code for
//   everything 2.69 / 7.5 / 8.6.
#ifndef Z2D_RASTER_MIN_CTAS
#define Z2D_RASTER_MIN_CTAS (2048 / Z2D_RASTER_THREADS > 32 ? 32 : 2048 / Z2D_RASTER_THREADS)
#endif
#ifndef Z2D_RASTER_RICH_MIN_CTAS
#define Z2D_RASTER_RICH_MIN_CTAS (Z2D_RASTER_MIN_CTAS * 7 / 8)  // 28 CTAs, 72 registers (24 / 80: config 3 raster 6.42 ms, this 6.26)
#endif
template <bool RICH>
Z2D_D void raster_tiles_body(const RasterArgs& A) {
  __shared__ __align__(16) uint32_t tile_px[kRasterThreads / 32][8 * 32];
  __shared__ uint4 blend_tab[kRasterThreads / 32][17];
  __shared__ __align__(16) uint2 diff_s[kRasterThreads / 32][66];  // per warp: row difference array of tile_cover32 (int[132] view: tile_cover)
  __shared__ __align__(16) uint4 edge_s[kRasterThreads / 32][2][32];  // per warp: the (<= 32) crossing edges of the current pair
  __shared__ uint32_t meta_s[kRasterThreads / 32][32];
  __shared__ __align__(16) DevGrad grad_s[kRasterThreads / 32];      // per warp: gradient / stops / source of the current pattern draw
  __shared__ float soff_s[kRasterThreads / 32][kGenMaxStops];
  __shared__ float4 scol_s[kRasterThreads / 32][kGenMaxStops];
  __shared__ DevSrc psrc_s[kRasterThreads / 32];                // per warp: row range | direction of compacted crossing edges
  const int warp = kRasterThreads == 32 ? 0 : (int)(threadIdx.x >> 5), lane = threadIdx.x & 31;
  const uint32_t gt = blockIdx.x * (kRasterThreads / 32) + warp;
  if (gt >= A.n_tiles) return;
  if (A.abort && *A.abort) return;  // (small-batch path: the prepare kernel gave up; the sized pipeline redoes the batch)
  // tile -> surface, tx, ty
  uint32_t si;
  {
    uint32_t lo = 0, hi = A.n_sfc;
    while (hi - lo > 1) {
      uint32_t mid = (lo + hi) >> 1;
      if (A.sfcs[mid].tile_base <= gt) lo = mid; else hi = mid;
    }
    si = lo;
  }
  const DevSurface S = A.sfcs[si];
  const uint32_t lt = gt - S.tile_base;
  const int ty_local = (int)(lt / (uint32_t)S.tiles_x), tx = (int)(lt % (uint32_t)S.tiles_x);
  const int ty = ty_local + (S.y0 >> kTileShift);  // canvas tile row (band surfaces start at row S.y0)
  // ordered draw list of this tile-row
  const uint32_t n_draws_s = S.draw_end - S.draw_begin;
  const uint32_t chunks = (n_draws_s + kDrawChunk - 1) / kDrawChunk;
  const uint32_t w0 = A.work_base[si] + (uint32_t)ty_local * chunks;
  const uint32_t lb = A.list_off[w0], le = A.list_off[w0 + chunks];
  if (lb == le) return;

  uint32_t* px = tile_px[warp];  // pixel i of lane l lives at word (i >> 2) * 128 + l * 4 + (i & 3): 128-bit, conflict-free
#define Z2D_PXI(i) ((((i) >> 2) << 7) + (lane << 2) + ((i) & 3))
  const int row = lane >> 1, half = lane & 1;
  const int py = ty * kTile + row;
  const int px0 = tx * kTile + half * 8;
  bool loaded = false, dirty = false;
  uint32_t n_cov = 0, n_eval = 0, n_pairs = 0;  // statistics: composited pixels, (edge, sub-scanline) crossings, (draw, tile) pairs
  const size_t row_idx = (size_t)(py - S.y0) * (size_t)S.w;
  const bool row_ok = py - S.y0 < S.h;
  TileFmt tf;
  tf.fmt = S.fmt;
  tf.is32 = S.fmt <= Z2D_FMT_RGBA;
  tf.f = fmt32_of(S.fmt);
  const bool vec_ok = tf.is32 && row_ok && px0 + 8 <= S.w && ((row_idx + (size_t)px0) & 3) == 0;  // 2 x 128-bit global access

  for (uint32_t base = lb; base < le; base += 32) {
    uint4 it = make_uint4(0u, 0u, 0u, 0u);
    bool hit = false;
    if (base + lane < le) {
      it = __ldg(A.list_items + base + lane);
      const int itx0 = (int)(it.y & 0xffffu), itx1 = (int)(it.y >> 16);
      hit = tx >= itx0 && tx <= itx1;
    }
    uint32_t hits = __ballot_sync(0xffffffffu, hit);
    n_pairs += (uint32_t)__popc(hits);
    while (hits) {
      const int src_lane = __ffs(hits) - 1;
      hits &= hits - 1;
      const uint32_t di = __shfl_sync(0xffffffffu, it.x, src_lane);
      const uint32_t iw = __shfl_sync(0xffffffffu, it.w, src_lane);
      const uint32_t ifl = iw >> 24;
      const DrawHot& h = A.hots[di];  // read on demand (uniform): region, source pixel; the list item carries the rest
      asm volatile("prefetch.global.L1 [%0];" ::"l"(&h));

      // ---- coverage
      const int aa = (int)(ifl & kItemAaMask);
      const uint32_t rule = (ifl & kItemEvenOdd) ? (uint32_t)Z2D_FILL_EVEN_ODD : (uint32_t)Z2D_FILL_NON_ZERO;
      const bool special = (ifl & kItemSpecial) != 0u;
      uint32_t cov_e = 0, cov_o = 0;  // per-pixel coverage bytes: even pixels in cov_e, odd in cov_o
      const bool in_rows = (ifl & kItemInRows) != 0u;
      const bool pre = special && h.unbounded && aa == Z2D_AA_MULTISAMPLE_4X;
      const bool all_px = aa == Z2D_AA_SUPERSAMPLE_4X;  // every pixel of the region is composited, even at coverage 0
      const bool rowrec = special && (h.flags & kDrawRowRecords) != 0u;
      if (in_rows) {
        uint32_t eb = __shfl_sync(0xffffffffu, it.z, src_lane), nbe = iw & 0xffffffu;
        if (nbe == 0xffffffu) {  // (count does not fit the item)
          const uint32_t bslot = h.band_base + (uint32_t)(ty - h.ey0);
          nbe = A.band_off[bslot + 1] - eb;
        }
        const DevEdge* be = A.band_edges + eb;
        const int4* hd = A.band_hdr + eb;
        const bool unpaired = special && (h.flags & kDrawUnpaired) != 0u;
        uint64_t m0 = 0, m1 = 0;
        if (aa != Z2D_AA_NONE) {
          const int sx0 = tx * kTile * 4;
          if ((!RICH && nbe > 32u) ||
              !tile_cover32<true, RICH>(be, hd, nbe, ty * kTile * 4, ty * kTile * 4 + lane * 2, sx0, rule, diff_s[warp], edge_s[warp][0], meta_s[warp], m0, m1, n_eval))
            tile_cover(be, hd, nbe, ty * kTile * 4, ty * kTile * 4 + lane * 2, true, sx0, 64, rule, reinterpret_cast<int*>(diff_s[warp]), m0, m1, n_eval);
          if (unpaired) {
            m0 = cut_open_tail(A, h, ty * kTile * 4 + lane * 2, sx0, 64, m0);
            m1 = cut_open_tail(A, h, ty * kTile * 4 + lane * 2 + 1, sx0, 64, m1);
          }
          // pixel row `row` needs sub-scanlines 4*row .. 4*row+3: this lane's two and its partner's two
          const uint64_t q0 = __shfl_xor_sync(0xffffffffu, m0, 1), q1 = __shfl_xor_sync(0xffffffffu, m1, 1);
          const int sh = half * 32;
          const uint32_t a = nibble_popc((uint32_t)(m0 >> sh)), b = nibble_popc((uint32_t)(m1 >> sh));
          const uint32_t c = nibble_popc((uint32_t)(q0 >> sh)), e2 = nibble_popc((uint32_t)(q1 >> sh));
          cov_e = (a & 0x0f0f0f0fu) + (b & 0x0f0f0f0fu) + (c & 0x0f0f0f0fu) + (e2 & 0x0f0f0f0fu);
          cov_o = ((a >> 4) & 0x0f0f0f0fu) + ((b >> 4) & 0x0f0f0f0fu) + ((c >> 4) & 0x0f0f0f0fu) + ((e2 >> 4) & 0x0f0f0f0fu);
        } else {
          const int sx0 = tx * kTile;
          if ((!RICH && nbe > 32u) ||
              !tile_cover32<false, RICH>(be, hd, nbe, ty * kTile, ty * kTile + row, sx0, rule, diff_s[warp], edge_s[warp][0], meta_s[warp], m0, m1, n_eval))
            tile_cover(be, hd, nbe, ty * kTile, ty * kTile + row, false, sx0, 16, rule, reinterpret_cast<int*>(diff_s[warp]), m0, m1, n_eval);
          if (unpaired) m0 = cut_open_tail(A, h, ty * kTile + row, sx0, 16, m0);
          const uint32_t bits = ((uint32_t)m0 >> (half * 8)) & 0xffu;
          for (int i = 0; i < 8; i += 2) {
            cov_e |= ((bits >> i) & 1u) << (4 * i);  // byte i/2
            cov_o |= ((bits >> (i + 1)) & 1u) << (4 * i);
          }
        }
      }
      // nothing of this draw lands in the tile?
      const bool lane_work = row_ok && (pre || rowrec || (py >= h.ry0 && py < h.ry1 && (all_px || (cov_e | cov_o) != 0u)));
      if (!__any_sync(0xffffffffu, lane_work)) continue;

      if (!loaded) {  // lazy tile load: 8 pixels per lane
        if (vec_ok) {
          const uint4* g = reinterpret_cast<const uint4*>(S.data) + ((row_idx + (size_t)px0) >> 2);
          reinterpret_cast<uint4*>(px)[lane] = g[0];
          reinterpret_cast<uint4*>(px)[32 + lane] = g[1];
        } else {
          for (int i = 0; i < 8; i++) {
            const int x = px0 + i;
            px[Z2D_PXI(i)] = (x < S.w && row_ok) ? load_raw(S.data, S.fmt, row_idx + (size_t)x) : 0u;
          }
        }
        loaded = true;
      }

      // ---- composite the lane's 8 pixels
      const RGBA16 spx = unpack_rgba(h.px_rgba);
      if (rowrec) {
        // direct.zig with an unbounded operator (anti-aliasing none): every row of the surface is rewritten from its row
        // record {start, end of the last processed span pair, pairs processed, filtered crossings} (k_edge_sim; direct.zig:88-124)
        const int rr = py - h.sim_y0;
        if (row_ok && rr >= 0 && rr < h.sim_rows) {
          const int4 rec = __ldg(&A.sim_rows[h.sim_base + (uint32_t)rr]);
          if (rec.w == 0 || rec.z > 0) {  // (crossings but no pair processed: the row is left untouched)
            const DevDraw& dd = A.draws[di];
#pragma unroll 1
            for (int i = 0; i < 8; i++) {
              const int x = px0 + i;
              if (x >= S.w) break;
              uint32_t raw = 0u;  // cleared: a row without crossings, or outside the surviving pair
              if (rec.w != 0 && x >= rec.x && x < rec.y) {
                if (rec.z == 1) raw = px[Z2D_PXI(i)];  // (a later pair composites over what the previous pair's tail clear left)
                n_cov++;
                raw = composite_cov(h, dd, A.T, tf, spx, raw, 1, x, py);
              }
              px[Z2D_PXI(i)] = raw;
            }
          }
        }
        __syncwarp();
        dirty = true;
        continue;
      }
      if (ifl & kItemFastBlend) {
        // fast path: single-pixel source, integer src_over.  Lanes 1..16 build the source at each coverage level
        // (multisample.zig:223: alpha 16 * cov - 1; full coverage: unmasked) in blend-ready form, then every lane blends
        // its 8 pixels straight-line (no divergence between fully and partially covered pixels, no register indexing).
        uint4* bt = blend_tab[warp];
        const int full = aa == Z2D_AA_NONE ? 1 : 16;
        if (lane >= 1 && lane <= full) bt[lane] = blend_entry(tf.f, spx, 16 * lane - 1, lane < full);
        __syncwarp();
        if (!(row_ok && py >= h.ry0 && py < h.ry1)) cov_e = cov_o = 0u;
        {
          const int lo = max(h.rx0 - px0, 0), hi = min(min(h.rx1, S.w) - px0, 8);
          if (lo > 0 || hi < 8) {  // region / surface edge inside this lane's run of 8 pixels
            uint32_t keep_e = 0u, keep_o = 0u;
            for (int i = 0; i < 8; i++)
              if (i >= lo && i < hi) ((i & 1) ? keep_o : keep_e) |= 0xffu << (8 * (i >> 1));
            cov_e &= keep_e;
            cov_o &= keep_o;
          }
        }
        n_cov += nonzero_bytes(cov_e) + nonzero_bytes(cov_o);
        if ((cov_e | cov_o) != 0u) {
          uint4 v0 = reinterpret_cast<uint4*>(px)[lane], v1 = reinterpret_cast<uint4*>(px)[32 + lane];
          blend_fast8(v0, v1, cov_e, cov_o, bt, tf.f.has_a ? 0x00ff00ffu : 0x000000ffu);
          reinterpret_cast<uint4*>(px)[lane] = v0;
          reinterpret_cast<uint4*>(px)[32 + lane] = v1;
        }
        __syncwarp();
        dirty = true;
        continue;
      }
      const DevDraw& d = A.draws[di];
      if (RICH && (d.src.kind == Z2D_PARAM_GRADIENT || d.src.kind == Z2D_PARAM_DITHER) &&
          (!(d.src.kind == Z2D_PARAM_GRADIENT || d.src.dither_source == Z2D_DITHER_SRC_GRADIENT) || A.T.grads[d.src.grad].n_stops <= (uint32_t)kGenMaxStops)) {
        n_cov += pattern_tile(&d, &h, A.T, px, &grad_s[warp], soff_s[warp], scol_s[warp], &psrc_s[warp], cov_e, cov_o, px0, py, S.w, S.fmt,
                              (pre ? 1u : 0u) | (all_px ? 2u : 0u) | (row_ok ? 4u : 0u));
        dirty = true;
        continue;
      }
#pragma unroll 1  // one copy of the generic compositor (28 operators x sources x formats) instead of eight
      for (int i = 0; i < 8; i++) {
        const int x = px0 + i;
        const int cov = (int)(((i & 1) ? cov_o : cov_e) >> (8 * (i >> 1))) & 0xff;
        const bool in_sfc = x < S.w && row_ok;
        const bool in_reg = in_sfc && x >= h.rx0 && x < h.rx1 && py >= h.ry0 && py < h.ry1;
        const bool work = in_sfc && (pre || (in_reg && (all_px || cov != 0)));
        if (!__any_sync(0xffffffffu, work)) continue;
        if (!work) continue;
        uint32_t raw = px[Z2D_PXI(i)];
        if (pre) {  // multisample.zig:96-110
          if (py < h.pre_y0 || (py > h.pre_y1 && py < h.pre_rows) || (py >= h.pre_y0 && py <= h.pre_y1 && x < h.pre_x)) raw = 0u;
        }
        if (in_reg) {
          n_cov += cov > 0;
          raw = composite_cov(h, d, A.T, tf, spx, raw, cov, x, py);
        }
        px[Z2D_PXI(i)] = raw;
      }
      dirty = true;
    }
  }
  if (dirty) {
    if (vec_ok) {
      uint4* g = reinterpret_cast<uint4*>(S.data) + ((row_idx + (size_t)px0) >> 2);
      g[0] = reinterpret_cast<uint4*>(px)[lane];
      g[1] = reinterpret_cast<uint4*>(px)[32 + lane];
    } else {
      for (int i = 0; i < 8; i++) {
        const int x = px0 + i;
        if (x < S.w && row_ok) store_raw(S.data, S.fmt, row_idx + (size_t)x, px[Z2D_PXI(i)]);
      }
    }
  }
#undef Z2D_PXI
  if (A.counters) {
    n_cov = __reduce_add_sync(0xffffffffu, n_cov);
    n_eval = __reduce_add_sync(0xffffffffu, n_eval);
    if (lane == 0) {
      if (n_cov) atomicAdd(&A.counters[0], (unsigned long long)n_cov);
      if (n_pairs) atomicAdd(&A.counters[2], (unsigned long long)n_pairs);
      if (n_eval) atomicAdd(&A.counters[3], (unsigned long long)n_eval);
    }
  }
}

#ifdef Z2D_RASTER_TU  // the kernels are instantiated by raster.cu only; kernels.cu includes this file for the blend helpers
__global__ void __launch_bounds__(kRasterThreads, Z2D_RASTER_MIN_CTAS) k_raster_tiles(const __grid_constant__ RasterArgs A) { raster_tiles_body<false>(A); }
__global__ void __launch_bounds__(kRasterThreads, Z2D_RASTER_RICH_MIN_CTAS) k_raster_tiles_rich(const __grid_constant__ RasterArgs A) {
  raster_tiles_body<true>(A);
}
#endif

}  // namespace z2d
