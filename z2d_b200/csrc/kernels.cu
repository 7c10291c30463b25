// Hand-written sm_100a kernels of the fill / composite pipeline (see
// z2d_batch.cuh for the stage list).  No tensor cores: nothing here is a dense
// contraction; the work is f64 edge evaluation, bit-sliced integer winding and
// byte-wise compositing, bounded by HBM for the compositor and by the f64 /
// integer pipes for coverage.
#include "kernels.cuh"

namespace z2d {

// =====================================================================================
// exclusive scan (u32): 256 threads x 8 items per block, recursive over block sums
// =====================================================================================
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanBlock = kScanThreads * kScanItems;

__global__ void __launch_bounds__(kScanThreads) k_scan_block(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                                                             uint32_t* __restrict__ sums, uint32_t n) {
  __shared__ uint32_t sh[kScanThreads];
  const uint32_t base = blockIdx.x * kScanBlock + threadIdx.x * kScanItems;
  uint32_t v[kScanItems];
  uint32_t local = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; i++) {
    v[i] = (base + i < n) ? in[base + i] : 0u;
    local += v[i];
  }
  sh[threadIdx.x] = local;
  __syncthreads();
  for (int off = 1; off < kScanThreads; off <<= 1) {  // Hillis-Steele inclusive scan of the thread sums
    uint32_t t = (threadIdx.x >= (unsigned)off) ? sh[threadIdx.x - off] : 0u;
    __syncthreads();
    sh[threadIdx.x] += t;
    __syncthreads();
  }
  uint32_t run = sh[threadIdx.x] - local;  // exclusive prefix of this thread
#pragma unroll
  for (int i = 0; i < kScanItems; i++) {
    if (base + i < n) out[base + i] = run;
    run += v[i];
  }
  if (threadIdx.x == kScanThreads - 1) sums[blockIdx.x] = sh[threadIdx.x];
}

__global__ void k_scan_add(uint32_t* __restrict__ out, const uint32_t* __restrict__ block_off, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] += block_off[i / kScanBlock];
}

// out[0..n) = exclusive scan of in[0..n); out[n] = total.  `tmp` must hold scan_tmp_len(n) words.
size_t scan_tmp_len(uint32_t n) {
  size_t total = 0;
  while (n > 1) {
    uint32_t nb = (n + kScanBlock - 1) / kScanBlock;
    total += (size_t)nb * 2 + 2;
    if (nb == 1) break;
    n = nb;
  }
  return total + 4;
}

static void scan_rec(const uint32_t* in, uint32_t* out, uint32_t n, uint32_t* total, uint32_t* tmp, cudaStream_t st) {
  uint32_t nb = (n + kScanBlock - 1) / kScanBlock;
  uint32_t* sums = tmp;
  uint32_t* sums_scanned = tmp + nb + 1;
  k_scan_block<<<nb, kScanThreads, 0, st>>>(in, out, sums, n);
  if (nb == 1) {
    cudaMemcpyAsync(total, sums, 4, cudaMemcpyDeviceToDevice, st);
    return;
  }
  scan_rec(sums, sums_scanned, nb, total, tmp + 2 * (size_t)nb + 2, st);
  k_scan_add<<<(n + 255) / 256, 256, 0, st>>>(out, sums_scanned, n);
}

void exclusive_scan(const uint32_t* in, uint32_t* out, uint32_t n, uint32_t* tmp, cudaStream_t st) {
  if (n == 0) {
    cudaMemsetAsync(out, 0, 4, st);
    return;
  }
  scan_rec(in, out, n, out + n, tmp, st);
}

// =====================================================================================
// K1: flatten (fill) -- thread per sub-path
// =====================================================================================
struct Pt {
  double x, y;
};
Z2D_D bool pt_eq(Pt a, Pt b) { return a.x == b.x && a.y == b.y; }

struct Knots {
  Pt a, b, c, d;
};

Z2D_D double knots_error_sq(const Knots& k) {  // tess/Spline.zig:83-123
  double bx = k.b.x - k.a.x, by = k.b.y - k.a.y, cx = k.c.x - k.a.x, cy = k.c.y - k.a.y;
  if (k.a.x != k.d.x || k.a.y != k.d.y) {
    double dx = k.d.x - k.a.x, dy = k.d.y - k.a.y;
    double dd = dx * dx + dy * dy;
    double bd = bx * dx + by * dy;
    if (bd >= dd) {
      bx -= dx;
      by -= dy;
    } else {
      bx -= bd / dd * dx;
      by -= bd / dd * dy;
    }
    double cd = cx * dx + cy * dy;
    if (cd >= dd) {
      cx -= dx;
      cy -= dy;
    } else {
      cx -= cd / dd * dx;
      cy -= cd / dd * dy;
    }
  }
  double be = bx * bx + by * by, ce = cx * cx + cy * cy;
  return be > ce ? be : ce;
}
Z2D_D Pt lerp_half(Pt a, Pt b) { return {a.x + ((b.x - a.x) / 2), a.y + ((b.y - a.y) / 2)}; }
Z2D_D Knots knots_split(Knots& k) {  // tess/Spline.zig:128-151 (k becomes the first half)
  Pt ab = lerp_half(k.a, k.b), bc = lerp_half(k.b, k.c), cd = lerp_half(k.c, k.d);
  Pt abbc = lerp_half(ab, bc), bccd = lerp_half(bc, cd);
  Pt fin = lerp_half(abbc, bccd);
  Knots r{fin, bccd, cd, k.d};
  k.b = ab;
  k.c = abbc;
  k.d = fin;
  return r;
}

// Edge sink: applies Polygon.addEdge (tess/Polygon.zig:61-109).  EMIT=false counts and tracks extents.
template <bool EMIT>
struct EdgeSink {
  double scale;
  uint32_t n = 0;
  double top, bottom, left, right;
  DevEdge* out = nullptr;
  uint32_t* out_draw = nullptr;
  uint32_t draw = 0;
  Z2D_D void add(Pt p0, Pt p1) {
    double ax = p0.x * scale, ay = p0.y * scale, bx = p1.x * scale, by = p1.y * scale;
    DevEdge e;
    if (ay < by) {
      e = {ay, by, ax, (bx - ax) / (by - ay)};
    } else if (ay > by) {
      e = {ay, by, bx, (ax - bx) / (ay - by)};
    } else {
      return;
    }
    if (EMIT) {
      out[n] = e;
      out_draw[n] = draw;
    } else {
      double t = ay < by ? ay : by, b = ay < by ? by : ay;
      double l = ax < bx ? ax : bx, r = ax < bx ? bx : ax;
      if (n == 0) {
        top = t; bottom = b; left = l; right = r;
      } else {
        if (t < top) top = t;
        if (b > bottom) bottom = b;
        if (l < left) left = l;
        if (r > right) right = r;
      }
    }
    n++;
  }
};

constexpr int kSplineStack = 48;

// fill_plotter.plot (tess/fill_plotter.zig:21-97) restricted to one sub-path
// (the plotter state resets at every move_to).
template <bool EMIT>
Z2D_D void fill_subpath(const z2d_node* __restrict__ nodes, uint32_t begin, uint32_t end, double tol, EdgeSink<EMIT>& sink) {
  Pt first{0, 0}, last{0, 0};
  int len = 0;  // PointBuffer(1,3): first point + sliding window; only first/last/len matter
  auto add_pt = [&](Pt p) {
    if (len == 0) first = p;
    if (len < 3) len++;
    last = p;
  };
  auto line_to = [&](Pt p) {
    if (!pt_eq(last, p)) {
      sink.add(last, p);
      add_pt(p);
    }
  };
  const double tol_sq = tol * tol;
  for (uint32_t i = begin; i < end; i++) {
    const z2d_node nd = nodes[i];
    switch (nd.tag) {
      case Z2D_NODE_MOVE_TO:
        len = 0;
        add_pt({nd.p[0], nd.p[1]});
        break;
      case Z2D_NODE_LINE_TO:
        if (len > 0) line_to({nd.p[0], nd.p[1]});
        break;
      case Z2D_NODE_CURVE_TO: {
        if (len == 0) break;
        const Pt a = last, b{nd.p[0], nd.p[1]}, c{nd.p[2], nd.p[3]}, d{nd.p[4], nd.p[5]};
        if (pt_eq(a, b) && pt_eq(c, d)) {  // Spline.zig:39-42
          line_to(d);
          break;
        }
        Knots stack[kSplineStack];
        int sp = 0;
        stack[sp++] = Knots{a, b, c, d};
        while (sp > 0) {  // Spline.zig:56-71, depth first, left half first
          Knots k = stack[--sp];
          if (knots_error_sq(k) < tol_sq || sp >= kSplineStack - 2) {
            if (!pt_eq(k.a, a)) line_to(k.a);
            continue;
          }
          Knots s2 = knots_split(k);
          stack[sp++] = s2;
          stack[sp++] = k;
        }
        line_to(d);
        break;
      }
      default:  // close_path (fill_plotter.zig:72-92)
        if (len >= 3) {
          if (pt_eq(last, first)) break;
          sink.add(last, first);
          add_pt(first);
        }
    }
  }
}

__global__ void k_flatten_count(const DevSubPath* __restrict__ sps, uint32_t n_sp, const z2d_node* __restrict__ nodes,
                                DevDraw* __restrict__ draws, uint32_t* __restrict__ sp_count) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_sp) return;
  const DevSubPath sp = sps[i];
  DevDraw& d = draws[sp.draw];
  EdgeSink<false> sink;
  sink.scale = d.scale;
  if (d.kind == 0) fill_subpath<false>(nodes, sp.node_begin, sp.node_end, d.tolerance, sink);
  sp_count[i] = sink.n;
  if (sink.n > 0) {
    atomicMin(&d.ext[0], f64_order(sink.top));
    atomicMax(&d.ext[1], f64_order(sink.bottom));
    atomicMin(&d.ext[2], f64_order(sink.left));
    atomicMax(&d.ext[3], f64_order(sink.right));
    atomicAdd(&d.n_edges, sink.n);
  }
}

__global__ void k_flatten_emit(const DevSubPath* __restrict__ sps, uint32_t n_sp, const z2d_node* __restrict__ nodes,
                               const DevDraw* __restrict__ draws, const uint32_t* __restrict__ sp_off,
                               DevEdge* __restrict__ edges, uint32_t* __restrict__ edge_draw) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_sp) return;
  const DevSubPath sp = sps[i];
  const DevDraw& d = draws[sp.draw];
  EdgeSink<true> sink;
  sink.scale = d.scale;
  sink.out = edges + sp_off[i];
  sink.out_draw = edge_draw + sp_off[i];
  sink.draw = sp.draw;
  if (d.kind == 0) fill_subpath<true>(nodes, sp.node_begin, sp.node_end, d.tolerance, sink);
}

// =====================================================================================
// K2: per-draw setup (regions, tile ranges)
// =====================================================================================
Z2D_D int clampi(int v, int lo, int hi) { return max(lo, min(v, hi)); }

__global__ void k_reset_draws(DevDraw* __restrict__ draws, uint32_t n_draws) {  // replay: undo what the pipeline wrote
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_draws) return;
  DevDraw& d = draws[i];
  d.ext[0] = f64_order(INFINITY);
  d.ext[1] = f64_order(-INFINITY);
  d.ext[2] = f64_order(INFINITY);
  d.ext[3] = f64_order(-INFINITY);
  d.n_edges = 0;
  d.valid = 0;
}

__global__ void k_setup_draws(DevDraw* __restrict__ draws, uint32_t n_draws, const DevSurface* __restrict__ sfcs,
                              uint32_t* __restrict__ draw_bands, unsigned long long* __restrict__ counters) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_draws) return;
  DevDraw& d = draws[i];
  const DevSurface s = sfcs[d.surface];
  d.valid = 0;
  draw_bands[i] = 0;
  if (d.n_edges == 0) return;
  const double top = f64_unorder(d.ext[0]), bottom = f64_unorder(d.ext[1]);
  const double left = f64_unorder(d.ext[2]), right = f64_unorder(d.ext[3]);
  const int W = s.w, H = s.h;
  const double sc = d.scale;
  // Polygon.inBox (tess/Polygon.zig:142-201)
  if (right < 0.0 || bottom < 0.0) return;
  const int psx = (int)floor(left / sc), psy = (int)floor(top / sc);
  const int pex = (int)ceil(right / sc), pey = (int)ceil(bottom / sc);
  const int pw = pex - psx, ph = pey - psy;
  if (pw == 0 || ph == 0) return;
  if (psx + pw < 0 || psy + ph < 0) return;
  if (psx >= W || psy >= H) return;

  const bool unbounded = !op_is_bounded(d.op);
  d.unbounded = unbounded ? 1u : 0u;
  int rx0, rx1, ry0, ry1;
  if (d.aa == Z2D_AA_MULTISAMPLE_4X) {  // raster/multisample.zig:46-78
    const int y0 = clampi(psy, 0, H - 1), y1 = clampi(pey, y0, H - 1);
    const int x0 = clampi(psx, 0, W - 1), x1 = clampi(pex, x0, W);
    if (x1 - x0 < 1) return;
    rx0 = x0; rx1 = x1; ry0 = y0; ry1 = y1 + 1;
    d.pre_y0 = y0;
    d.pre_y1 = y1;
    d.pre_x = max(x0, (x1 < W) ? (W - x1) : 0);  // sic: both side clears start at x = 0 (multisample.zig:104-108)
    d.pre_rows = min(W, H);                      // sic: rows after the box are cleared up to sfc_width (multisample.zig:101)
  } else if (d.aa == Z2D_AA_SUPERSAMPLE_4X) {  // raster/supersample.zig:48-72, surface.zig:447-469
    const int x0 = unbounded ? 0 : psx, y0 = unbounded ? 0 : psy;
    const int x1 = unbounded ? W : pex, y1 = unbounded ? H : pey;
    rx0 = max(0, x0); ry0 = max(0, y0);
    rx1 = min(x1, W); ry1 = min(y1, H);
    if (rx1 <= rx0 || ry1 <= ry0) return;
  } else {  // raster/direct.zig:44-49 (bounded operators; unbounded ones take the scanline path)
    const int y0 = clampi((int)floor(top), 0, H - 1), y1 = clampi((int)ceil(bottom), y0, H - 1);
    ry0 = y0; ry1 = y1 + 1;
    rx0 = clampi((int)floor(left) - 1, 0, W);
    rx1 = clampi((int)ceil(right) + 1, rx0, W);
    if (rx1 <= rx0) return;
  }
  d.rx0 = rx0; d.rx1 = rx1; d.ry0 = ry0; d.ry1 = ry1;
  d.ey0 = ry0 >> kTileShift;
  d.ey1 = (ry1 - 1) >> kTileShift;
  if (unbounded && d.aa != Z2D_AA_NONE) {  // the whole surface is touched
    d.tx0 = 0; d.tx1 = s.tiles_x - 1; d.ty0 = 0; d.ty1 = s.tiles_y - 1;
  } else {
    d.tx0 = rx0 >> kTileShift; d.tx1 = (rx1 - 1) >> kTileShift;
    d.ty0 = d.ey0; d.ty1 = d.ey1;
  }
  d.valid = 1;
  draw_bands[i] = (uint32_t)(d.ey1 - d.ey0 + 1);
  if (counters) atomicAdd(&counters[1], (unsigned long long)(rx1 - rx0) * (unsigned long long)(ry1 - ry0));
}

__global__ void k_assign_band_base(DevDraw* __restrict__ draws, uint32_t n_draws, const uint32_t* __restrict__ band_off) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_draws) draws[i].band_base = band_off[i];
}

// =====================================================================================
// K3a: bin edges into (draw, tile-row) lists.  An edge is active on sub-scanline ys iff
// top < ys + 0.5 <= bottom (tess/Polygon.zig:284-285).
// =====================================================================================
Z2D_D bool edge_band_range(const DevEdge& e, const DevDraw& d, int& t0, int& t1) {
  const double top = e.y0 < e.y1 ? e.y0 : e.y1, bottom = e.y0 < e.y1 ? e.y1 : e.y0;
  const double ys_min_f = floor(top - 0.5) + 1.0, ys_max_f = floor(bottom - 0.5);
  if (ys_max_f < ys_min_f) return false;
  const int S = (d.aa == Z2D_AA_NONE) ? 1 : 4;
  // clamp in f64 first (coordinates may be far outside the surface)
  const double lo = (double)d.ry0 * S, hi = (double)d.ry1 * S - 1.0;
  const double a = ys_min_f > lo ? ys_min_f : lo, b = ys_max_f < hi ? ys_max_f : hi;
  if (b < a) return false;
  const int py0 = (int)a / S, py1 = (int)b / S;  // non-negative
  t0 = py0 >> kTileShift;
  t1 = py1 >> kTileShift;
  return true;
}

__global__ void k_bin_count(const DevEdge* __restrict__ edges, const uint32_t* __restrict__ edge_draw, uint32_t n_edges,
                            const DevDraw* __restrict__ draws, uint32_t* __restrict__ band_count) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_edges) return;
  const DevDraw& d = draws[edge_draw[i]];
  if (!d.valid) return;
  int t0, t1;
  if (!edge_band_range(edges[i], d, t0, t1)) return;
  for (int t = t0; t <= t1; t++) atomicAdd(&band_count[d.band_base + (uint32_t)(t - d.ey0)], 1u);
}

__global__ void k_bin_scatter(const DevEdge* __restrict__ edges, const uint32_t* __restrict__ edge_draw, uint32_t n_edges,
                              const DevDraw* __restrict__ draws, const uint32_t* __restrict__ band_off,
                              uint32_t* __restrict__ band_cursor, DevEdge* __restrict__ band_edges) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_edges) return;
  const DevDraw& d = draws[edge_draw[i]];
  if (!d.valid) return;
  const DevEdge e = edges[i];
  int t0, t1;
  if (!edge_band_range(e, d, t0, t1)) return;
  for (int t = t0; t <= t1; t++) {
    const uint32_t b = d.band_base + (uint32_t)(t - d.ey0);
    const uint32_t slot = band_off[b] + atomicAdd(&band_cursor[b], 1u);
    band_edges[slot] = e;
  }
}

// =====================================================================================
// K3b: per surface tile-row, the ordered list of draws that touch it.
// Work item w = (surface, tile-row, chunk of kDrawChunk consecutive draws), laid out
// row-major so that after the scan each tile-row's list is contiguous and in draw order.
// =====================================================================================
Z2D_D uint32_t find_surface_by(const DevSurface* sfcs, uint32_t n_sfc, const uint32_t* bases, uint32_t v) {
  uint32_t lo = 0, hi = n_sfc;  // last s with bases[s] <= v
  while (hi - lo > 1) {
    uint32_t mid = (lo + hi) >> 1;
    if (bases[mid] <= v) lo = mid; else hi = mid;
  }
  return lo;
}

template <bool WRITE>
__global__ void k_band_lists(const DevSurface* __restrict__ sfcs, uint32_t n_sfc, const uint32_t* __restrict__ work_base,
                             uint32_t n_work, const DevDraw* __restrict__ draws, uint32_t* __restrict__ cnt,
                             const uint32_t* __restrict__ off, uint2* __restrict__ items) {
  uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n_work) return;
  const uint32_t si = find_surface_by(sfcs, n_sfc, work_base, w);
  const DevSurface s = sfcs[si];
  const uint32_t n_draws = s.draw_end - s.draw_begin;
  const uint32_t chunks = (n_draws + kDrawChunk - 1) / kDrawChunk;
  const uint32_t local = w - work_base[si];
  const int band = (int)(local / chunks);
  const uint32_t chunk = local % chunks;
  const uint32_t b = s.draw_begin + chunk * kDrawChunk;
  const uint32_t e = min(b + kDrawChunk, s.draw_end);
  uint32_t n = 0;
  uint32_t o = WRITE ? off[w] : 0u;
  for (uint32_t i = b; i < e; i++) {
    const DevDraw& d = draws[i];
    if (d.valid && band >= d.ty0 && band <= d.ty1) {
      if (WRITE) items[o + n] = make_uint2(i, (uint32_t)d.tx0 | ((uint32_t)d.tx1 << 16));
      n++;
    }
  }
  if (!WRITE) cnt[w] = n;
}
template __global__ void k_band_lists<false>(const DevSurface*, uint32_t, const uint32_t*, uint32_t, const DevDraw*, uint32_t*,
                                             const uint32_t*, uint2*);
template __global__ void k_band_lists<true>(const DevSurface*, uint32_t, const uint32_t*, uint32_t, const DevDraw*, uint32_t*,
                                            const uint32_t*, uint2*);

// =====================================================================================
// per-pixel compositing for a rasterised draw (raster/shared.zig, multisample.zig:195-227,
// supersample.zig:159-184, direct.zig:88-124)
// =====================================================================================
Z2D_D RGBA16 mask_mul16(RGBA16 s, int m) { return {iM(s.r, m), iM(s.g, m), iM(s.b, m), iM(s.a, m)}; }  // dst_in(dst:=s, src:=alpha8 m)

// generic StrideCompositor batch: [dst_in(pattern, mask)]? ; op        (shared.zig:24-45, 78-102)
__device__ __noinline__ uint32_t composite_generic(const DevDraw& d, const GradTables& T, uint32_t fmt, uint32_t raw, int mask8,
                                                   bool use_mask, int x, int y) {
  if (d.precision == Z2D_PRECISION_INTEGER) {
    RGBA16 s = src_int(d.src, T, x, y, 0);
    if (use_mask) s = mask_mul16(s, mask8);
    return rgba16_to_raw(fmt, int_op(d.op, raw_to_rgba16(fmt, raw), s));
  }
  RGBAF s = src_float(d.src, T, x, y, 0);
  if (use_mask) {
    const float ma = (float)mask8 / 255.0f;
    s = {s.r * ma, s.g * ma, s.b * ma, ma * s.a};
  }
  RGBAF r = float_op(d.op, decode_raw(raw_to_rgba16(fmt, raw)), s);
  return rgba16_to_raw(fmt, encode_raw(r));
}

// cov: number of covered samples (MSAA/SSAA: 0..16, none: 0..1).  Returns the new raw pixel.
Z2D_D uint32_t composite_cov(const DevDraw& d, const GradTables& T, uint32_t fmt, uint32_t raw, int cov, int x, int y) {
  if (d.aa == Z2D_AA_SUPERSAMPLE_4X) {
    // mask = box average of 16 samples in the mask surface's own format (supersample.zig:82-91, pixel.zig:435-464,633-646)
    int m8;
    if (fmt == Z2D_FMT_ALPHA4 || fmt == Z2D_FMT_ALPHA2 || fmt == Z2D_FMT_ALPHA1) {
      const int bits = fmt_bits(fmt);
      m8 = scale_alpha((((1 << bits) - 1) * cov) / 16, bits, 8);
    } else {
      m8 = (255 * cov) / 16;
    }
    return composite_generic(d, T, fmt, raw, m8, true, x, y);
  }
  if (cov == 0) return raw;
  if (d.op == Z2D_OP_CLEAR) return 0u;  // shared.zig:18,60 (also at partial coverage)
  const bool full = (d.aa == Z2D_AA_NONE) || cov == 16;
  if (full) {
    if (d.reduces) return d.paint_raw;
    return composite_generic(d, T, fmt, raw, 255, false, x, y);
  }
  const int o = 16 * cov - 1;  // multisample.zig:223
  if (d.reduces) {             // surface.zig:557-581 compositeStride (integer only)
    RGBA16 s = mask_mul16(unpack_rgba(d.src.px_rgba), o);
    return rgba16_to_raw(fmt, int_op(d.op, raw_to_rgba16(fmt, raw), s));
  }
  return composite_generic(d, T, fmt, raw, o, true, x, y);
}

// =====================================================================================
// K4: fused coverage + compositing, one warp per 16x16-pixel tile
// =====================================================================================
template <int W>
Z2D_D void wind_add(uint64_t (&p)[W], uint64_t mask, int dir) {
  uint64_t c = mask;
  if (dir > 0) {
#pragma unroll
    for (int k = 0; k < W; k++) {
      const uint64_t t = p[k] & c;
      p[k] ^= c;
      c = t;
    }
  } else {
#pragma unroll
    for (int k = 0; k < W; k++) {
      const uint64_t t = ~p[k] & c;
      p[k] ^= c;
      c = t;
    }
  }
}

// Accumulate the inside-masks of this lane's sub-scanlines for one draw in one tile.
// Sample column s (device sample units) is inside iff the signed count of active edges
// with x_i <= s is non-zero (non_zero) / odd (even_odd)  -- Polygon.zig:302-353.
Z2D_D double4 ld_edge(const DevEdge* e) {  // 2 x 128-bit read-only loads
  const double2* q = reinterpret_cast<const double2*>(e);
  const double2 a = __ldg(q), b = __ldg(q + 1);
  return make_double4(a.x, a.y, b.x, b.y);
}

// Column (relative to the tile's first sample) at which edge `ev` crosses sub-scanline
// centre ym, or -1 when the edge is inactive there / crosses right of the tile.
Z2D_D int edge_col(const double4& ev, double top, double bottom, double ym, int sx0, int ncols) {
  if (!(top < ym && ym <= bottom)) return -1;
  const double xf = round_half_away(ev.z + (ev.w * (ym - top)));  // Polygon.zig:305
  const double cf = xf - (double)sx0;
  if (!(cf < (double)ncols)) return -1;
  return cf < 0.0 ? 0 : (int)cf;
}

// Inside-mask of ONE sub-scanline of a tile, bit-sliced winding in W registers (|winding| < 2^(W-1)).
template <int W>
Z2D_D uint64_t row_mask(const DevEdge* __restrict__ be, uint32_t n_be, int ys, int sx0, int ncols, uint32_t rule) {
  uint64_t p[W];
#pragma unroll
  for (int k = 0; k < W; k++) p[k] = 0ull;
  const double ym = (double)ys + 0.5;
  const double xlim = (double)(sx0 + ncols) + 1.0;
  for (uint32_t i = 0; i < n_be; i++) {
    const double4 ev = ld_edge(be + i);  // y0,y1,x_start,x_inc (warp-uniform address)
    const bool down = ev.x < ev.y;
    const double top = down ? ev.x : ev.y, bottom = down ? ev.y : ev.x;
    // warp-uniform cull: an edge entirely to the right of the tile contributes nothing
    const double xe = ev.z + ev.w * (bottom - top);
    if ((ev.z < xe ? ev.z : xe) > xlim) continue;
    const int col = edge_col(ev, top, bottom, ym, sx0, ncols);
    if (col >= 0) {
      const uint64_t mask = ~0ull << col;
      if (rule == Z2D_FILL_EVEN_ODD) p[0] ^= mask; else wind_add<W>(p, mask, down ? -1 : 1);
    }
  }
  if (rule == Z2D_FILL_EVEN_ODD) return p[0];
  uint64_t a = 0;
#pragma unroll
  for (int k = 0; k < W; k++) a |= p[k];
  return a;
}

// Same, for tile-rows holding so many edges of one draw that the winding number could
// exceed the register-resident counter: 32 bit-planes in local memory (rare).
__device__ __noinline__ uint64_t row_mask_wide(const DevEdge* __restrict__ be, uint32_t n_be, int ys, int sx0, int ncols, uint32_t rule) {
  uint64_t p[32];
  for (int k = 0; k < 32; k++) p[k] = 0ull;
  const double ym = (double)ys + 0.5;
  for (uint32_t i = 0; i < n_be; i++) {
    const double4 ev = ld_edge(be + i);
    const bool down = ev.x < ev.y;
    const double top = down ? ev.x : ev.y, bottom = down ? ev.y : ev.x;
    const int col = edge_col(ev, top, bottom, ym, sx0, ncols);
    if (col < 0) continue;
    uint64_t c = ~0ull << col;
    if (rule == Z2D_FILL_EVEN_ODD) {
      p[0] ^= c;
    } else if (!down) {
      for (int k = 0; k < 32 && c; k++) { const uint64_t t = p[k] & c; p[k] ^= c; c = t; }
    } else {
      for (int k = 0; k < 32 && c; k++) { const uint64_t t = ~p[k] & c; p[k] ^= c; c = t; }
    }
  }
  if (rule == Z2D_FILL_EVEN_ODD) return p[0];
  uint64_t a = 0;
  for (int k = 0; k < 32; k++) a |= p[k];
  return a;
}

Z2D_D uint64_t row_mask_any(const DevEdge* __restrict__ be, uint32_t n_be, int ys, int sx0, int ncols, uint32_t rule) {
  if (n_be < 120) return row_mask<8>(be, n_be, ys, sx0, ncols, rule);
  return row_mask_wide(be, n_be, ys, sx0, ncols, rule);
}

Z2D_D uint32_t nibble_popc(uint32_t x) {  // per-nibble popcount (values 0..4 in each 4-bit field)
  x = x - ((x >> 1) & 0x55555555u);
  return (x & 0x33333333u) + ((x >> 2) & 0x33333333u);
}

__global__ void __launch_bounds__(kRasterThreads) k_raster_tiles(RasterArgs A) {
  __shared__ uint32_t tile_px[kRasterThreads / 32][8 * 32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t gt = blockIdx.x * (kRasterThreads / 32) + warp;
  if (gt >= A.n_tiles) return;
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
  const int ty = (int)(lt / (uint32_t)S.tiles_x), tx = (int)(lt % (uint32_t)S.tiles_x);
  // ordered draw list of this tile-row
  const uint32_t n_draws_s = S.draw_end - S.draw_begin;
  const uint32_t chunks = (n_draws_s + kDrawChunk - 1) / kDrawChunk;
  const uint32_t w0 = A.work_base[si] + (uint32_t)ty * chunks;
  const uint32_t lb = A.list_off[w0], le = A.list_off[w0 + chunks];
  if (lb == le) return;

  uint32_t* px = tile_px[warp];
  const int row = lane >> 1, half = lane & 1;
  const int py = ty * kTile + row;
  const int px0 = tx * kTile + half * 8;
  bool loaded = false, dirty = false;
  uint32_t n_cov = 0;
  const size_t row_idx = (size_t)py * (size_t)S.w;

  for (uint32_t base = lb; base < le; base += 32) {
    uint2 it = make_uint2(0, 0);
    bool hit = false;
    if (base + lane < le) {
      it = A.list_items[base + lane];
      const int itx0 = (int)(it.y & 0xffffu), itx1 = (int)(it.y >> 16);
      hit = tx >= itx0 && tx <= itx1;
    }
    uint32_t hits = __ballot_sync(0xffffffffu, hit);
    while (hits) {
      const int src_lane = __ffs(hits) - 1;
      hits &= hits - 1;
      const uint32_t di = __shfl_sync(0xffffffffu, it.x, src_lane);
      const DevDraw& d = A.draws[di];

      if (!loaded) {  // lazy tile load: 8 pixels per lane
        for (int i = 0; i < 8; i++) {
          const int x = px0 + i;
          px[i * 32 + lane] = (x < S.w && py < S.h) ? load_raw(S.data, S.fmt, row_idx + (size_t)x) : 0u;
        }
        loaded = true;
      }

      // ---- coverage
      const int aa = (int)d.aa;
      const int Sc = (aa == Z2D_AA_NONE) ? 1 : 4;
      uint32_t cov_e = 0, cov_o = 0;  // per-pixel coverage bytes: even pixels in cov_e, odd in cov_o
      const bool in_rows = ty >= d.ey0 && ty <= d.ey1;
      if (in_rows) {
        const uint32_t bslot = d.band_base + (uint32_t)(ty - d.ey0);
        const uint32_t eb = A.band_off[bslot], ee = A.band_off[bslot + 1];
        const DevEdge* be = A.band_edges + eb;
        const uint32_t nbe = ee - eb;
        uint64_t m0 = 0, m1 = 0;
        const int sx0 = tx * kTile * Sc;
        if (Sc == 4) {
          const int ys0 = ty * kTile * 4 + lane * 2;
          m0 = row_mask_any(be, nbe, ys0, sx0, 64, d.rule);
          m1 = row_mask_any(be, nbe, ys0 + 1, sx0, 64, d.rule);
          // pixel row `row` needs sub-scanlines 4*row .. 4*row+3: this lane's two and its partner's two
          const uint64_t q0 = __shfl_xor_sync(0xffffffffu, m0, 1), q1 = __shfl_xor_sync(0xffffffffu, m1, 1);
          const int sh = half * 32;
          const uint32_t a = nibble_popc((uint32_t)(m0 >> sh)), b = nibble_popc((uint32_t)(m1 >> sh));
          const uint32_t c = nibble_popc((uint32_t)(q0 >> sh)), e2 = nibble_popc((uint32_t)(q1 >> sh));
          cov_e = (a & 0x0f0f0f0fu) + (b & 0x0f0f0f0fu) + (c & 0x0f0f0f0fu) + (e2 & 0x0f0f0f0fu);
          cov_o = ((a >> 4) & 0x0f0f0f0fu) + ((b >> 4) & 0x0f0f0f0fu) + ((c >> 4) & 0x0f0f0f0fu) + ((e2 >> 4) & 0x0f0f0f0fu);
        } else {
          m0 = row_mask_any(be, nbe, ty * kTile + row, sx0, 16, d.rule);
          const uint32_t bits = ((uint32_t)m0 >> (half * 8)) & 0xffu;
          for (int i = 0; i < 8; i += 2) {
            cov_e |= ((bits >> i) & 1u) << (4 * i);        // byte i/2
            cov_o |= ((bits >> (i + 1)) & 1u) << (4 * i);
          }
        }
      }

      // ---- composite the lane's 8 pixels
      const bool pre = d.unbounded && aa == Z2D_AA_MULTISAMPLE_4X;
      for (int i = 0; i < 8; i++) {
        const int x = px0 + i;
        if (x >= S.w || py >= S.h) continue;
        uint32_t raw = px[i * 32 + lane];
        if (pre) {  // multisample.zig:96-110
          if (py < d.pre_y0 || (py > d.pre_y1 && py < d.pre_rows) || (py >= d.pre_y0 && py <= d.pre_y1 && x < d.pre_x)) raw = 0u;
        }
        if (x >= d.rx0 && x < d.rx1 && py >= d.ry0 && py < d.ry1) {
          const int cov = (int)(((i & 1) ? cov_o : cov_e) >> (8 * (i >> 1))) & 0xff;
          n_cov += cov > 0;
          raw = composite_cov(d, A.T, S.fmt, raw, cov, x, py);
        }
        px[i * 32 + lane] = raw;
      }
      dirty = true;
    }
  }
  if (dirty) {
    for (int i = 0; i < 8; i++) {
      const int x = px0 + i;
      if (x < S.w && py < S.h) store_raw(S.data, S.fmt, row_idx + (size_t)x, px[i * 32 + lane]);
    }
  }
  if (A.counters) {
    n_cov = __reduce_add_sync(0xffffffffu, n_cov);
    if (lane == 0 && n_cov) atomicAdd(&A.counters[0], (unsigned long long)n_cov);
  }
}

// =====================================================================================
// K5: surface-level compositor (SurfaceCompositor.run / StrideCompositor.run)
// =====================================================================================
Z2D_D uint32_t comp_pixel(const CompArgs& A, uint32_t raw, int dx, int dy, int sxp, int syp) {
  // (dx,dy): destination pixel; (sxp,syp): matching source-space pixel (compositor.zig:389-436)
  if (A.precision == Z2D_PRECISION_INTEGER) {
    RGBA16 d{0, 0, 0, 0}, s{0, 0, 0, 0};
    for (uint32_t k = 0; k < A.n_ops; k++) {
      const CompOp& o = A.ops[k];
      s = o.has_src ? src_int(o.src, A.T, dx, dy, (size_t)o.src.sw * (size_t)syp + (size_t)sxp) : d;
      d = o.has_dst ? src_int(o.dst, A.T, dx, dy, (size_t)o.dst.sw * (size_t)dy + (size_t)dx) : raw_to_rgba16(A.fmt, raw);
      d = int_op(o.op, d, s);
    }
    return rgba16_to_raw(A.fmt, d);
  }
  RGBAF d{0, 0, 0, 0}, s{0, 0, 0, 0};
  for (uint32_t k = 0; k < A.n_ops; k++) {
    const CompOp& o = A.ops[k];
    s = o.has_src ? src_float(o.src, A.T, dx, dy, (size_t)o.src.sw * (size_t)syp + (size_t)sxp) : d;
    d = o.has_dst ? src_float(o.dst, A.T, dx, dy, (size_t)o.dst.sw * (size_t)dy + (size_t)dx) : decode_raw(raw_to_rgba16(A.fmt, raw));
    d = float_op(o.op, d, s);
  }
  return rgba16_to_raw(A.fmt, encode_raw(d));
}

// generic: one thread per pixel of the composited rectangle
__global__ void __launch_bounds__(256) k_composite_px(const __grid_constant__ CompArgs A) {
  const size_t n = (size_t)A.scan_w * (size_t)A.rows;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / (size_t)A.scan_w), c = (int)(i % (size_t)A.scan_w);
    const int dx = A.dst_start_x + c, dy = A.dst_start_y + r;
    const int sxp = A.src_start_x + c, syp = A.src_start_y + r;
    const size_t idx = (size_t)dy * (size_t)A.w + (size_t)dx;
    const uint32_t raw = load_raw(A.data, A.fmt, idx);
    store_raw(A.data, A.fmt, idx, comp_pixel(A, raw, dx, dy, sxp, syp));
  }
}

// 32-bit formats, full-width contiguous rows: 4 pixels (16 B) per thread, 128-bit loads/stores
__global__ void __launch_bounds__(256) k_composite_v4(const __grid_constant__ CompArgs A) {
  const size_t n4 = ((size_t)A.scan_w * (size_t)A.rows) >> 2;
  uint4* base = reinterpret_cast<uint4*>(A.data + ((size_t)A.dst_start_y * (size_t)A.w) * 4);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    uint4 v = base[i];
    const size_t p = i << 2;
    const int r = (int)(p / (size_t)A.scan_w), c = (int)(p % (size_t)A.scan_w);  // scan_w % 4 == 0: the 4 pixels share a row
    const int dy = A.dst_start_y + r, syp = A.src_start_y + r;
    v.x = comp_pixel(A, v.x, c + 0, dy, A.src_start_x + c + 0, syp);
    v.y = comp_pixel(A, v.y, c + 1, dy, A.src_start_x + c + 1, syp);
    v.z = comp_pixel(A, v.z, c + 2, dy, A.src_start_x + c + 2, syp);
    v.w = comp_pixel(A, v.w, c + 3, dy, A.src_start_x + c + 3, syp);
    base[i] = v;
  }
}

// whole-surface paint (Surface.paintPixel / initPixel) and single pixel put
__global__ void k_paint(uint8_t* data, uint32_t fmt, size_t n_px, uint32_t raw) {
  if (fmt <= Z2D_FMT_RGBA) {
    uint32_t* p = (uint32_t*)data;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_px; i += (size_t)gridDim.x * blockDim.x) p[i] = raw;
    return;
  }
  // sub-32-bit formats: fill whole bytes with the replicated sample (surface.zig:889-915)
  const int bits = fmt_bits(fmt);
  uint32_t b = 0;
  for (int sh = 0; sh < 8; sh += bits) b |= raw << sh;
  const size_t nbytes = (n_px * (size_t)bits + 7) / 8;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nbytes; i += (size_t)gridDim.x * blockDim.x) data[i] = (uint8_t)b;
}

__global__ void k_put_pixel(uint8_t* data, uint32_t fmt, size_t idx, uint32_t raw) { store_raw(data, fmt, idx, raw); }

// ------------------------------------------------------------------------- launchers
static inline unsigned blocks_for(size_t n, unsigned threads) { return (unsigned)((n + threads - 1) / threads); }

void launch_flatten_count(const DevSubPath* sps, uint32_t n_sp, const z2d_node* nodes, DevDraw* draws, uint32_t* sp_count, cudaStream_t st) {
  if (n_sp) k_flatten_count<<<blocks_for(n_sp, 128), 128, 0, st>>>(sps, n_sp, nodes, draws, sp_count);
}
void launch_flatten_emit(const DevSubPath* sps, uint32_t n_sp, const z2d_node* nodes, const DevDraw* draws, const uint32_t* sp_off,
                         DevEdge* edges, uint32_t* edge_draw, cudaStream_t st) {
  if (n_sp) k_flatten_emit<<<blocks_for(n_sp, 128), 128, 0, st>>>(sps, n_sp, nodes, draws, sp_off, edges, edge_draw);
}
void launch_setup_draws(DevDraw* draws, uint32_t n, const DevSurface* sfcs, uint32_t* draw_bands, unsigned long long* counters, cudaStream_t st) {
  if (n) k_setup_draws<<<blocks_for(n, 128), 128, 0, st>>>(draws, n, sfcs, draw_bands, counters);
}
void launch_reset_draws(DevDraw* draws, uint32_t n, cudaStream_t st) {
  if (n) k_reset_draws<<<blocks_for(n, 256), 256, 0, st>>>(draws, n);
}
void launch_assign_band_base(DevDraw* draws, uint32_t n, const uint32_t* band_off, cudaStream_t st) {
  if (n) k_assign_band_base<<<blocks_for(n, 256), 256, 0, st>>>(draws, n, band_off);
}
void launch_bin_count(const DevEdge* edges, const uint32_t* edge_draw, uint32_t n, const DevDraw* draws, uint32_t* band_count, cudaStream_t st) {
  if (n) k_bin_count<<<blocks_for(n, 256), 256, 0, st>>>(edges, edge_draw, n, draws, band_count);
}
void launch_bin_scatter(const DevEdge* edges, const uint32_t* edge_draw, uint32_t n, const DevDraw* draws, const uint32_t* band_off,
                        uint32_t* band_cursor, DevEdge* band_edges, cudaStream_t st) {
  if (n) k_bin_scatter<<<blocks_for(n, 256), 256, 0, st>>>(edges, edge_draw, n, draws, band_off, band_cursor, band_edges);
}
void launch_band_lists(bool write, const DevSurface* sfcs, uint32_t n_sfc, const uint32_t* work_base, uint32_t n_work,
                       const DevDraw* draws, uint32_t* cnt, const uint32_t* off, uint2* items, cudaStream_t st) {
  if (!n_work) return;
  if (write)
    k_band_lists<true><<<blocks_for(n_work, 128), 128, 0, st>>>(sfcs, n_sfc, work_base, n_work, draws, cnt, off, items);
  else
    k_band_lists<false><<<blocks_for(n_work, 128), 128, 0, st>>>(sfcs, n_sfc, work_base, n_work, draws, cnt, off, items);
}
void launch_raster(const RasterArgs& A, cudaStream_t st) {
  if (A.n_tiles) k_raster_tiles<<<blocks_for(A.n_tiles, kRasterThreads / 32), kRasterThreads, 0, st>>>(A);
}
void launch_composite(const CompArgs& A, int sm_count, cudaStream_t st) {
  const size_t n = (size_t)A.scan_w * (size_t)A.rows;
  if (n == 0) return;
  const bool vec = A.fmt <= Z2D_FMT_RGBA && A.dst_start_x == 0 && A.scan_w == A.w && (A.scan_w & 3) == 0;
  const size_t items = vec ? (n >> 2) : n;
  unsigned blocks = (unsigned)((items + 255) / 256);
  const unsigned cap = (unsigned)sm_count * 8u * 4u;  // grid-stride: a few waves of 8 resident CTAs per SM
  if (blocks > cap) blocks = cap;
  if (vec)
    k_composite_v4<<<blocks, 256, 0, st>>>(A);
  else
    k_composite_px<<<blocks, 256, 0, st>>>(A);
}
void launch_paint(uint8_t* data, uint32_t fmt, size_t n_px, uint32_t raw, cudaStream_t st) {
  unsigned blocks = (unsigned)((n_px + 255) / 256);
  if (blocks > 148u * 16u) blocks = 148u * 16u;
  if (blocks == 0) blocks = 1;
  k_paint<<<blocks, 256, 0, st>>>(data, fmt, n_px, raw);
}
void launch_put_pixel(uint8_t* data, uint32_t fmt, size_t idx, uint32_t raw, cudaStream_t st) { k_put_pixel<<<1, 1, 0, st>>>(data, fmt, idx, raw); }

}  // namespace z2d
