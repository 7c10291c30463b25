// Hand-written sm_100a kernels of the fill / composite pipeline (see
// z2d_batch.cuh for the stage list).  No tensor cores: nothing here is a dense
// contraction; the work is f64 edge evaluation, bit-sliced integer winding and
// byte-wise compositing, bounded by HBM for the compositor and by the f64 /
// integer pipes for coverage.
#include <algorithm>

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
}  // namespace z2d
#include "geom.cuh"
namespace z2d {

// Extents of a draw's edges, gathered from many threads: most of them lie inside what others already reported, so the
// current value is looked at before paying for an atomic on a word that hundreds of threads share.
Z2D_D void ext_commit(DevDraw& d, double top, double bottom, double left, double right) {
  const volatile long long* ext = d.ext;
  const long long t = f64_order(top), b = f64_order(bottom), l = f64_order(left), r = f64_order(right);
  if (t < ext[0]) atomicMin(&d.ext[0], t);
  if (b > ext[1]) atomicMax(&d.ext[1], b);
  if (l < ext[2]) atomicMin(&d.ext[2], l);
  if (r > ext[3]) atomicMax(&d.ext[3], r);
}

// fill_plotter.plot (tess/fill_plotter.zig:21-97) restricted to one sub-path
// (the plotter state resets at every move_to).
template <bool EMIT>
Z2D_D void fill_subpath(const z2d_node* __restrict__ nodes, uint32_t begin, uint32_t end, double tol, EdgeSink<EMIT>& sink) {
  Pt first{0, 0}, last{0, 0};
  int len = 0;  // PointBuffer(1,3): first point + sliding window; only first/last/len matter
  auto add_pt = [&](Pt p) Z2D_LAMBDA {
    if (len == 0) first = p;
    if (len < 3) len++;
    last = p;
  };
  auto line_to = [&](Pt p) Z2D_LAMBDA {
    if (!pt_eq(last, p)) {
      sink.add(last, p);
      add_pt(p);
    }
  };
  const double tol_sq = tol * tol;
  #pragma unroll 1
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
        spline_decompose(last, {nd.p[0], nd.p[1]}, {nd.p[2], nd.p[3]}, {nd.p[4], nd.p[5]}, tol_sq, line_to);
        break;
      }
      default:  // close_path (fill_plotter.zig:72-92)
        if (len >= 3) {
          if (pt_eq(last, first)) break;
          sink.add(last, first);
          add_pt(first);
        } else if (len == 2) {
          sink.unpaired = true;  // a lone edge that is never closed: rows crossing it have an unpaired crossing
        }
    }
  }
}

}  // namespace z2d
#include "stroke.cuh"
#include "stroke_units.cuh"
namespace z2d {

#ifndef Z2D_FLATTEN_THREADS
#define Z2D_FLATTEN_THREADS 64
#endif
#ifndef Z2D_FLATTEN_MIN_CTAS
#define Z2D_FLATTEN_MIN_CTAS 1
#endif
__global__ void __launch_bounds__(Z2D_FLATTEN_THREADS, Z2D_FLATTEN_MIN_CTAS) k_flatten_count(const DevSubPath* __restrict__ sps, uint32_t n_sp, const z2d_node* __restrict__ nodes,
                                DevDraw* __restrict__ draws, uint32_t* __restrict__ sp_count, const PenV* __restrict__ pens,
                                const double* __restrict__ dashes, const uint32_t* __restrict__ order) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_sp) return;
  if (order) i = order[i];
  const DevSubPath sp = sps[i];
  if (i == 0 || sps[i - 1].draw != sp.draw) draws[sp.draw].sp_first = i;  // sub-paths of a draw are consecutive
  if (sp.flags & (kSpNodeParallel | kSpStrokeUnits)) {
    sp_count[i] = 0;
    return;
  }
  DevDraw& d = draws[sp.draw];
  EdgeSink<false> sink;
  sink.scale = d.scale;
  if (d.kind == 0) fill_subpath<false>(nodes, sp.node_begin, sp.node_end, d.tolerance, sink);
  else stroke_subpath<false>(nodes, sp.node_begin, sp.node_end, d, pens, dashes, sink);
  sp_count[i] = sink.n;
  if (sink.unpaired && sink.n > 0) atomicOr(&d.flags, kDrawUnpaired);
  if (sink.n > 0) {
    ext_commit(d, sink.top, sink.bottom, sink.left, sink.right);
    atomicAdd(&d.n_edges, sink.n);
  }
}

__global__ void __launch_bounds__(Z2D_FLATTEN_THREADS, Z2D_FLATTEN_MIN_CTAS) k_flatten_emit(const DevSubPath* __restrict__ sps, uint32_t n_sp, const z2d_node* __restrict__ nodes,
                               const DevDraw* __restrict__ draws, const uint32_t* __restrict__ sp_off,
                               DevEdge* __restrict__ edges, uint32_t* __restrict__ edge_draw, const PenV* __restrict__ pens,
                               const double* __restrict__ dashes, const uint32_t* __restrict__ order) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_sp) return;
  if (order) i = order[i];
  const DevSubPath sp = sps[i];
  if (sp.flags & (kSpNodeParallel | kSpStrokeUnits)) return;
  const DevDraw& d = draws[sp.draw];
  EdgeSink<true> sink;
  sink.scale = d.scale;
  sink.out = edges + sp_off[i];
  sink.out_draw = edge_draw + sp_off[i];
  sink.draw = sp.draw;
  sink.limit = sp_off[i + 1] - sp_off[i];
  if (d.kind == 0) fill_subpath<true>(nodes, sp.node_begin, sp.node_end, d.tolerance, sink);
  else stroke_subpath<true>(nodes, sp.node_begin, sp.node_end, d, pens, dashes, sink);
}

// ---- thread -> sub-path order for the sequential plotters.  A warp whose 32 lanes walk 32 differently styled strokes runs
// with 2.4 active lanes (ncu, profiles/r01_stroke_flatten_ncu.json); lanes that plot the same kind of stroke follow the same
// branches.  Sub-paths are bucketed by (dashed, join, cap, pen size class, node count) with a counting sort; the heaviest
// buckets (dashed, many nodes) come first so they do not form the tail of the launch.  Only the thread mapping changes: every
// sub-path still writes its own count slot and its own edge range.
constexpr int kSpKeys = 1024;
Z2D_D uint32_t sp_key(const DevSubPath& sp, const DevDraw* __restrict__ draws) {
  if (sp.flags & kSpNodeParallel) return 0;  // returns at once in the sequential kernels: keep them together at the end
  // (kSpStrokeUnits sub-paths are keyed like every stroke: the walker's lanes diverge on dashed / node count as well)
  const DevDraw& d = draws[sp.draw];
  const uint32_t nn = min(sp.node_end - sp.node_begin, 15u);
  if (d.kind == 0) return 1 + nn;  // irregular fills
  const uint32_t lg = 31u - (uint32_t)__clz(d.pen_count | 1u);           // pen vertices: 4-7, 8-15, 16-31, 32+
  const uint32_t pen = lg < 2u ? 0u : min(lg - 2u, 3u);
  return ((d.dash_count ? 1u : 0u) << 9) | (nn << 5) | (pen << 3) | ((d.join == Z2D_JOIN_ROUND ? 1u : 0u) << 2) | min(d.cap, 3u);
}
__global__ void k_sp_hist(const DevSubPath* __restrict__ sps, uint32_t n_sp, const DevDraw* __restrict__ draws, uint32_t* __restrict__ hist) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_sp) atomicAdd(&hist[sp_key(sps[i], draws)], 1u);
}
__global__ void __launch_bounds__(kSpKeys) k_sp_cursor(uint32_t* __restrict__ hist) {  // descending exclusive scan, in place
  __shared__ uint32_t s[kSpKeys];
  const int t = threadIdx.x;
  const uint32_t mine = hist[kSpKeys - 1 - t];  // t = 0 is the largest key
  s[t] = mine;
  __syncthreads();
  for (int off = 1; off < kSpKeys; off <<= 1) {
    const uint32_t v = t >= off ? s[t - off] : 0u;
    __syncthreads();
    s[t] += v;
    __syncthreads();
  }
  hist[kSpKeys - 1 - t] = s[t] - mine;
}
__global__ void k_sp_scatter(const DevSubPath* __restrict__ sps, uint32_t n_sp, const DevDraw* __restrict__ draws, uint32_t* __restrict__ cursor,
                             uint32_t* __restrict__ order) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_sp) order[atomicAdd(&cursor[sp_key(sps[i], draws)], 1u)] = i;
}
void launch_sp_order(const DevSubPath* sps, uint32_t n_sp, const DevDraw* draws, uint32_t* keys_scratch, uint32_t* order, cudaStream_t st) {
  if (!n_sp) return;
  cudaMemsetAsync(keys_scratch, 0, kSpKeys * 4, st);
  k_sp_hist<<<(n_sp + 255) / 256, 256, 0, st>>>(sps, n_sp, draws, keys_scratch);
  k_sp_cursor<<<1, kSpKeys, 0, st>>>(keys_scratch);
  k_sp_scatter<<<(n_sp + 255) / 256, 256, 0, st>>>(sps, n_sp, draws, keys_scratch, order);
}

// ---- node-parallel flattening of simple fill sub-paths (move_to, segments..., close_path; the host guarantees at least two
// segments that move, so the implicit close applies, fill_plotter.zig:82-91).  The plotter's current point before a node is
// the previous node's end point (a point equal to the current one is never added, fill_plotter.zig:46,107), so every node can
// be flattened on its own: one thread per NODE instead of per sub-path -- 4-7x more threads for curve-heavy scenes and no
// thread walks several curves back to back.
__global__ void k_mark_nodes(const DevSubPath* __restrict__ sps, uint32_t n_sp, uint32_t* __restrict__ node_sp) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_sp) return;
  const DevSubPath sp = sps[i];
  if (!(sp.flags & kSpNodeParallel)) return;
  for (uint32_t k = sp.node_begin; k < sp.node_end; k++) node_sp[k] = i;
}

// The edges of one node of a node-parallel sub-path (line_to / curve_to / close_path) into `sink`.
template <class Sink>
Z2D_D void flatten_node_into(uint32_t i, const DevSubPath& sp, const z2d_node& nd, const z2d_node* __restrict__ nodes, double tolerance, Sink& sink) {
  const z2d_node pv = nodes[i - 1];  // move_to, line_to or curve_to
  Pt last = pv.tag == Z2D_NODE_CURVE_TO ? Pt{pv.p[4], pv.p[5]} : Pt{pv.p[0], pv.p[1]};
  auto line_to = [&](Pt p) Z2D_LAMBDA {
    if (!pt_eq(last, p)) {
      sink.add(last, p);
      last = p;
    }
  };
  if (nd.tag == Z2D_NODE_LINE_TO) {
    line_to({nd.p[0], nd.p[1]});
  } else if (nd.tag == Z2D_NODE_CURVE_TO) {
    // Spline.decompose as ONE loop with one flatness test per trip, so the lanes of a warp stay converged on the
    // expensive part (Knots.errorSq) whatever their subdivision depth
    const Pt a = last, b{nd.p[0], nd.p[1]}, c{nd.p[2], nd.p[3]}, e{nd.p[4], nd.p[5]};
    if (pt_eq(a, b) && pt_eq(c, e)) {  // Spline.zig:39-42
      line_to(e);
    } else {
      const double tol_sq = tolerance * tolerance;
      Knots stack[kSplineStack];
      int sp_n = 0;
      Knots k{a, b, c, e};
      for (;;) {
        if (!(knots_error_sq(k) < tol_sq || sp_n >= kSplineStack - 2)) {
          stack[sp_n++] = knots_split(k);  // k becomes the left half
        } else {
          if (!pt_eq(k.a, a)) line_to(k.a);
          if (sp_n == 0) break;
          k = stack[--sp_n];
        }
      }
      line_to(e);
    }
  } else {  // close_path: the closing edge back to the sub-path's first point
    const z2d_node mv = nodes[sp.node_begin];
    line_to({mv.p[0], mv.p[1]});
  }
}

// One node: count (EMIT = false: also extents) or emit its edges at the offset the scan of the counts gave it.
template <bool EMIT>
Z2D_D void flatten_node(uint32_t i, const DevSubPath& sp, const z2d_node& nd, const z2d_node* __restrict__ nodes, DevDraw* __restrict__ draws,
                        uint32_t* __restrict__ counts, const uint32_t* __restrict__ offs, DevEdge* __restrict__ edges,
                        uint32_t* __restrict__ edge_draw) {
  DevDraw& d = draws[sp.draw];
  EdgeSink<EMIT> sink;
  sink.scale = d.scale;
  if (EMIT) {
    sink.out = edges + offs[i];
    sink.out_draw = edge_draw + offs[i];
    sink.draw = sp.draw;
  }
  flatten_node_into(i, sp, nd, nodes, d.tolerance, sink);
  if (!EMIT) {
    counts[i] = sink.n;
    if (sink.n > 0) {
      ext_commit(d, sink.top, sink.bottom, sink.left, sink.right);
      atomicAdd(&d.n_edges, sink.n);
    }
  }
}

// Thread per node for line_to / close_path (one edge at most); curve_to nodes are collected into a dense list (count pass)
// and flattened by k_flatten_curves so that every lane of its warps runs the subdivision loop.
template <bool EMIT, bool CURVES>
__global__ void k_flatten_nodes(const DevSubPath* __restrict__ sps, const uint32_t* __restrict__ node_sp, uint32_t n_nodes,
                                const z2d_node* __restrict__ nodes, DevDraw* __restrict__ draws, uint32_t* __restrict__ counts,
                                const uint32_t* __restrict__ offs, DevEdge* __restrict__ edges, uint32_t* __restrict__ edge_draw,
                                uint32_t* __restrict__ curve_list, uint32_t* __restrict__ n_curves) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (CURVES) {  // i indexes the curve list
    if (i >= *n_curves) return;
    i = curve_list[i];
  } else if (i >= n_nodes) {
    return;
  }
  const uint32_t spi = node_sp[i];
  if (spi == 0xffffffffu) {
    if (!EMIT) counts[i] = 0;
    return;
  }
  const DevSubPath sp = sps[spi];
  const z2d_node nd = nodes[i];
  if (nd.tag == Z2D_NODE_MOVE_TO) {
    if (!EMIT) counts[i] = 0;
    return;
  }
  if (!CURVES && nd.tag == Z2D_NODE_CURVE_TO) {
    if (!EMIT) curve_list[atomicAdd(n_curves, 1u)] = i;  // (nvcc aggregates the increment per warp)
    return;
  }
  flatten_node<EMIT>(i, sp, nd, nodes, draws, counts, offs, edges, edge_draw);
}

// ---- node-parallel flattening in ONE pass.  The count pass above runs the whole subdivision just to learn how many edges
// a curve yields (0.4 of config 2's 1.2 ms of flattening).  An upper bound is free: with M = max |second difference| of the
// control polygon, the reference's flatness measure (Knots.errorSq: distance of the inner control points from the chord)
// is <= M, and de Casteljau halving divides the second differences by 4 (a - 2ab + abbc = (a - 2b + c) / 4,
// ab - 2abbc + fin = (a - b - c + d) / 8), so no piece is split beyond depth k = min{k : M^2 / 16^k < tol^2} and a curve
// yields at most 2^k edges.  Nodes take ranges of that size from the edge pool (scan + one atomicAdd per batch), write their
// edges, and mark the rest of the range dead (edge_draw = ~0, skipped by binning).  Extents come from the same pass, so K2
// runs after it.  A node that would exceed its bound (NaN / infinite coordinates) raises ctr[4] and the batch is redone by
// the two-pass kernels.
__global__ void k_node_bounds(const DevSubPath* __restrict__ sps, const uint32_t* __restrict__ node_sp, uint32_t n_nodes,
                              const z2d_node* __restrict__ nodes, const DevDraw* __restrict__ draws, uint32_t* __restrict__ counts,
                              uint32_t* __restrict__ curve_list, uint32_t* __restrict__ n_curves) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_nodes) return;
  const uint32_t spi = node_sp[i];
  uint32_t n = 0;
  if (spi != 0xffffffffu) {
    const z2d_node nd = nodes[i];
    if (nd.tag == Z2D_NODE_CURVE_TO) {
      const z2d_node pv = nodes[i - 1];
      const Pt a = pv.tag == Z2D_NODE_CURVE_TO ? Pt{pv.p[4], pv.p[5]} : Pt{pv.p[0], pv.p[1]};
      n = curve_edge_bound(a, {nd.p[0], nd.p[1]}, {nd.p[2], nd.p[3]}, {nd.p[4], nd.p[5]}, draws[sps[spi].draw].tolerance);
      curve_list[atomicAdd(n_curves, 1u)] = i;
    } else if (nd.tag != Z2D_NODE_MOVE_TO) {
      n = 1;
    }
  }
  counts[i] = n;
}

// ctr[3] = first pool slot of the batch's node ranges (ctr[2] is the pool cursor the stroke kernels share)
__global__ void k_take_node_range(uint32_t* __restrict__ ctr, const uint32_t* __restrict__ node_offs, uint32_t n_nodes) {
  ctr[3] = atomicAdd(&ctr[2], node_offs[n_nodes] - node_offs[0]);
}

template <bool CURVES>
__global__ void k_flatten_nodes_pool(const DevSubPath* __restrict__ sps, const uint32_t* __restrict__ node_sp, uint32_t n_nodes,
                                     const z2d_node* __restrict__ nodes, DevDraw* __restrict__ draws, const uint32_t* __restrict__ node_offs,
                                     uint32_t* __restrict__ ctr, DevEdge* __restrict__ edges, uint32_t* __restrict__ edge_draw, uint32_t edge_cap,
                                     const uint32_t* __restrict__ curve_list, const uint32_t* __restrict__ n_curves) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (CURVES) {
    if (i >= *n_curves) return;
    i = curve_list[i];
  } else if (i >= n_nodes) {
    return;
  }
  const uint32_t spi = node_sp[i];
  if (spi == 0xffffffffu) return;
  const z2d_node nd = nodes[i];
  if (nd.tag == Z2D_NODE_MOVE_TO || (!CURVES && nd.tag == Z2D_NODE_CURVE_TO)) return;
  const DevSubPath sp = sps[spi];
  const uint32_t first = ctr[3] + (node_offs[i] - node_offs[0]), bound = node_offs[i + 1] - node_offs[i];
  const uint32_t end = first + bound;
  PoolSink sink{edges, edge_draw, sp.draw, first, end < edge_cap ? end : edge_cap};
  sink.scale = draws[sp.draw].scale;
  flatten_node_into(i, sp, nd, nodes, draws[sp.draw].tolerance, sink);
  if (sink.pos > end) ctr[4] = 1u;  // more edges than the bound allows: the batch is redone by the two-pass kernels
  for (uint32_t p = sink.pos; p < sink.cap; p++) edge_draw[p] = 0xffffffffu;  // the rest of the range stays dead
  sink.commit(draws);
}

// =====================================================================================
// K2: per-draw setup (regions, tile ranges)
// =====================================================================================
Z2D_D int clampi(int v, int lo, int hi) { return max(lo, min(v, hi)); }

// DrawIn (+ side tables) -> DevDraw; also (re)initialises everything the pipeline writes, so a replay starts clean
Z2D_D void expand_draw(uint32_t i, const DrawIn* __restrict__ in, const StrokeIn* __restrict__ strokes, const DevSrc* __restrict__ srcs,
                        DevDraw* __restrict__ draws) {
  const DrawIn q = in[i];
  DevDraw d;
  memset(&d, 0, sizeof d);
  d.surface = q.surface;
  d.kind = q.opts & 1u;
  d.aa = (q.opts >> 1) & 3u;
  d.rule = (q.opts >> 3) & 1u;
  d.op = (q.opts >> 4) & 31u;
  d.precision = (q.opts >> 9) & 1u;
  d.reduces = (q.opts >> 10) & 1u;
  d.mode = (q.opts >> 11) & 3u;
  d.paint_raw = q.paint_raw;
  d.scale = d.aa == Z2D_AA_NONE ? 1.0 : 4.0;
  d.tolerance = q.tolerance;
  if (q.src_index == kNoIndex) {
    d.src.kind = Z2D_PARAM_PIXEL;
    d.src.px_rgba = q.px_rgba;
  } else {
    d.src = srcs[q.src_index];
  }
  if (q.stroke_index != kNoIndex) {
    const StrokeIn s = strokes[q.stroke_index];
    d.cap = s.cap; d.join = s.join;
    d.thickness = s.thickness; d.miter_limit = s.miter_limit; d.dash_offset = s.dash_offset;
    for (int k = 0; k < 6; k++) { d.ctm[k] = s.ctm[k]; d.inv[k] = s.inv[k]; }
    d.dash_begin = s.dash_begin; d.dash_count = s.dash_count;
    d.pen_begin = s.pen_begin; d.pen_count = s.pen_count;
    d.hair_aa = s.hair_aa; d.hair_tolerance = s.hair_tolerance;
  }
  d.ext[0] = f64_order(INFINITY);
  d.ext[1] = f64_order(-INFINITY);
  d.ext[2] = f64_order(INFINITY);
  d.ext[3] = f64_order(-INFINITY);
  draws[i] = d;
}
__global__ void k_expand_draws(const DrawIn* __restrict__ in, const StrokeIn* __restrict__ strokes, const DevSrc* __restrict__ srcs,
                               DevDraw* __restrict__ draws, uint32_t n_draws) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_draws) return;
  expand_draw(i, in, strokes, srcs, draws);
}

Z2D_D void setup_draw(uint32_t i, DevDraw* __restrict__ draws, const DevSurface* __restrict__ sfcs, uint32_t* __restrict__ draw_bands,
                       DrawBox* __restrict__ boxes, unsigned long long* __restrict__ counters) {
  DevDraw& d = draws[i];
  const DevSurface s = sfcs[d.surface];
  d.valid = 0;
  draw_bands[i] = 0;
  boxes[i] = DrawBox{-1, -1, -1, -1, 1, 0, 0u, 0u};
  if (d.n_edges == 0) return;
  const double top = f64_unorder(d.ext[0]), bottom = f64_unorder(d.ext[1]);
  const double left = f64_unorder(d.ext[2]), right = f64_unorder(d.ext[3]);
  const int W = s.w, H = s.vh;  // canvas height: a band surface clips afterwards
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
  } else if (d.mode == 2) {  // raster/direct.zig with an unbounded operator: every row of the surface, composited from row records
    rx0 = 0; rx1 = W; ry0 = 0; ry1 = H;
  } else {  // raster/direct.zig:44-49 (bounded operators)
    const int y0 = clampi((int)floor(top), 0, H - 1), y1 = clampi((int)ceil(bottom), y0, H - 1);
    ry0 = y0; ry1 = y1 + 1;
    rx0 = clampi((int)floor(left) - 1, 0, W);
    rx1 = clampi((int)ceil(right) + 1, rx0, W);
    if (rx1 <= rx0) return;
  }
  const int ry0c = ry0, ry1c = ry1;  // region rows of the canvas (a band surface clips below)
  // band surface: keep the part of the region that falls on the rows this surface holds
  ry0 = max(ry0, s.y0);
  ry1 = min(ry1, s.y0 + s.h);
  const bool rowrec = d.mode == 2;
  const bool whole = (unbounded && d.aa != Z2D_AA_NONE) || rowrec;  // the whole surface is touched (pre-clears / every pixel composited)
  if (ry1 <= ry0 && !whole) return;
  d.rx0 = rx0; d.rx1 = rx1; d.ry0 = ry0; d.ry1 = ry1;
  if ((rowrec || (d.flags & kDrawUnpaired)) && ry1c > ry0c) {
    // the outcome depends on the order the reference's sort leaves equal crossings in: k_edge_sim replays its scanline loop
    const int S = d.aa == Z2D_AA_NONE ? 1 : 4;
    d.sim_y0 = ry0c * S;
    d.sim_rows = (ry1c - ry0c) * S;
    d.sim_base = (uint32_t)atomicAdd(&counters[4], (unsigned long long)d.sim_rows);
    d.sim_scratch = (uint32_t)atomicAdd(&counters[5], (unsigned long long)d.n_edges);
  }
  if (rowrec) {  // no edges are binned: the row records carry the spans
    d.ey0 = 1;
    d.ey1 = 0;
  } else if (ry1 > ry0) {
    d.ey0 = ry0 >> kTileShift;  // canvas tile rows (edge binning)
    d.ey1 = (ry1 - 1) >> kTileShift;
  } else {
    d.ey0 = 1;  // empty
    d.ey1 = 0;
  }
  const int ty_org = s.y0 >> kTileShift;  // d.ty0 / d.ty1: tile rows of THIS surface (draw lists)
  if (whole) {
    d.tx0 = 0; d.tx1 = s.tiles_x - 1; d.ty0 = 0; d.ty1 = s.tiles_y - 1;
  } else {
    d.tx0 = rx0 >> kTileShift; d.tx1 = (rx1 - 1) >> kTileShift;
    d.ty0 = d.ey0 - ty_org; d.ty1 = d.ey1 - ty_org;
  }
  d.valid = 1;
  boxes[i] = DrawBox{d.tx0, d.tx1, d.ty0, d.ty1, d.ey0 - ty_org, d.ey1 - ty_org, 0u, 0u};
  draw_bands[i] = (uint32_t)(d.ey1 - d.ey0 + 1);
  if (counters && ry1 > ry0) atomicAdd(&counters[1], (unsigned long long)(rx1 - rx0) * (unsigned long long)(ry1 - ry0));
}
__global__ void k_setup_draws(DevDraw* __restrict__ draws, uint32_t n_draws, const DevSurface* __restrict__ sfcs,
                              uint32_t* __restrict__ draw_bands, DrawBox* __restrict__ boxes, unsigned long long* __restrict__ counters) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_draws) return;
  setup_draw(i, draws, sfcs, draw_bands, boxes, counters);
}

// second half of the setup: (draw, tile-row) slot base + the compact record the raster kernel reads
Z2D_D void assign_band_base(uint32_t i, DevDraw* __restrict__ draws, const uint32_t* __restrict__ band_off, DrawHot* __restrict__ hots,
                             DrawBox* __restrict__ boxes, const DevSurface* __restrict__ sfcs) {
  DevDraw& d = draws[i];
  d.band_base = band_off[i];
  {
    const bool pre = d.unbounded && d.aa == Z2D_AA_MULTISAMPLE_4X, rowrec = d.mode == 2 && d.sim_rows > 0;
    const bool unpaired = (d.flags & kDrawUnpaired) && d.sim_rows > 0;
    const bool special = pre || rowrec || unpaired || d.aa == Z2D_AA_SUPERSAMPLE_4X;
    const bool fast = !special && sfcs[d.surface].fmt <= Z2D_FMT_RGBA && d.src.kind == Z2D_PARAM_PIXEL && d.op == Z2D_OP_SRC_OVER &&
                      (d.reduces || d.precision == Z2D_PRECISION_INTEGER);
    boxes[i].band_base = d.band_base;
    boxes[i].item_flags = (d.aa & kItemAaMask) | (d.rule == Z2D_FILL_EVEN_ODD ? kItemEvenOdd : 0u) | (special ? kItemSpecial : 0u) |
                          (fast ? kItemFastBlend : 0u);
  }
  DrawHot h;
  h.aa = d.aa; h.rule = d.rule; h.op = d.op; h.precision = d.precision;
  h.reduces = d.reduces; h.paint_raw = d.paint_raw; h.px_rgba = d.src.px_rgba; h.src_kind = d.src.kind;
  h.rx0 = d.rx0; h.rx1 = d.rx1; h.ry0 = d.ry0; h.ry1 = d.ry1;
  h.ey0 = d.ey0; h.ey1 = d.ey1;
  h.band_base = d.band_base; h.unbounded = d.unbounded;
  h.pre_y0 = d.pre_y0; h.pre_y1 = d.pre_y1; h.pre_x = d.pre_x; h.pre_rows = d.pre_rows;
  h.flags = (d.flags & kDrawUnpaired) | (d.mode == 2 ? kDrawRowRecords : 0u);
  if (d.sim_rows <= 0) h.flags = 0;
  h.sim_base = d.sim_base; h.sim_y0 = d.sim_y0; h.sim_rows = d.sim_rows;
  hots[i] = h;
}
__global__ void k_assign_band_base(DevDraw* __restrict__ draws, uint32_t n_draws, const uint32_t* __restrict__ band_off,
                                   DrawHot* __restrict__ hots, DrawBox* __restrict__ boxes, const DevSurface* __restrict__ sfcs) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_draws) return;
  assign_band_base(i, draws, band_off, hots, boxes, sfcs);
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

Z2D_D void bin_count_edge(uint32_t i, const DevEdge* __restrict__ edges, const uint32_t* __restrict__ edge_draw, const DevDraw* __restrict__ draws,
                           uint32_t* __restrict__ band_count) {
  const uint32_t di = edge_draw[i];
  if (di == 0xffffffffu) return;  // dead pool slot (stroke_units.cuh)
  const DevDraw& d = draws[di];
  if (!d.valid || d.mode == 2) return;
  int t0, t1;
  if (!edge_band_range(edges[i], d, t0, t1)) return;
  for (int t = t0; t <= t1; t++) atomicAdd(&band_count[d.band_base + (uint32_t)(t - d.ey0)], 1u);
}
__global__ void k_bin_count(const DevEdge* __restrict__ edges, const uint32_t* __restrict__ edge_draw, uint32_t n_edges,
                            const DevDraw* __restrict__ draws, uint32_t* __restrict__ band_count) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_edges) return;
  bin_count_edge(i, edges, edge_draw, draws, band_count);
}

Z2D_D void bin_scatter_edge(uint32_t i, const DevEdge* __restrict__ edges, const uint32_t* __restrict__ edge_draw, const DevDraw* __restrict__ draws,
                             const uint32_t* __restrict__ band_off, uint32_t* __restrict__ band_cursor, DevEdge* __restrict__ band_edges,
                             int4* __restrict__ band_hdr, uint2* __restrict__ band_xr, uint32_t band_cap) {
  const uint32_t di = edge_draw[i];
  if (di == 0xffffffffu) return;
  const DevDraw& d = draws[di];
  if (!d.valid || d.mode == 2) return;
  const DevEdge e = edges[i];
  int t0, t1;
  if (!edge_band_range(e, d, t0, t1)) return;
  // integer header for the raster kernel's classification pass:
  //   x = conservative sample-column range [xlo, xhi] of every rounded crossing of this edge,
  //   z = first active sub-scanline (bit 31 set for an up edge, dir = +1), w = last active sub-scanline
  const bool down = e.y0 < e.y1;
  const double top = down ? e.y0 : e.y1, bottom = down ? e.y1 : e.y0;
  const double xe = e.x_start + e.x_inc * (bottom - top);
  const double xmin = e.x_start < xe ? e.x_start : xe, xmax = e.x_start < xe ? xe : e.x_start;
  const double big = 1.0e9;
  const double ya = fmax(floor(top - 0.5) + 1.0, 0.0), yb = fmin(floor(bottom - 0.5), big);
  int4 h;
  h.x = (int)fmax(fmin(floor(xmin) - 1.0, big), -big);
  h.y = (int)fmax(fmin(ceil(xmax) + 1.0, big), -big);
  h.z = (int)fmin(ya, big) | (down ? 0 : (int)0x80000000);
  h.w = (int)fmax(yb, -1.0);
  // tile columns this edge can touch (conservative, from the same bounds the raster kernel classifies with): the tile-row
  // lists visit only [min, max] over the row's edges.  Stored as {0x7fffffff - min, max + 1} so that zero means "none".
  const int S = (d.aa == Z2D_AA_NONE) ? 1 : 4;
  const int span = kTile * S;
  const uint32_t c_lo = 0x7fffffffu - (uint32_t)max(h.x - 1, 0) / (uint32_t)span;
  const uint32_t c_hi = (uint32_t)max(h.y + 1, 0) / (uint32_t)span + 1u;
  for (int t = t0; t <= t1; t++) {
    const uint32_t b = d.band_base + (uint32_t)(t - d.ey0);
    const uint32_t slot = band_off[b] + atomicAdd(&band_cursor[b], 1u);
    if (slot >= band_cap) continue;  // (small-batch path: fixed capacity, the batch is redone by the sized pipeline)
    band_edges[slot] = e;
    band_hdr[slot] = h;
    // (most edges of a row lie inside the column range the row's earlier edges reported: look before the atomic)
    const volatile uint32_t* seen = reinterpret_cast<const volatile uint32_t*>(&band_xr[b]);
    if (c_lo > seen[0]) atomicMax(&band_xr[b].x, c_lo);
    if (c_hi > seen[1]) atomicMax(&band_xr[b].y, c_hi);
  }
}
__global__ void k_bin_scatter(const DevEdge* __restrict__ edges, const uint32_t* __restrict__ edge_draw, uint32_t n_edges,
                              const DevDraw* __restrict__ draws, const uint32_t* __restrict__ band_off,
                              uint32_t* __restrict__ band_cursor, DevEdge* __restrict__ band_edges, int4* __restrict__ band_hdr,
                              uint2* __restrict__ band_xr) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_edges) return;
  bin_scatter_edge(i, edges, edge_draw, draws, band_off, band_cursor, band_edges, band_hdr, band_xr, 0xffffffffu);
}


// =====================================================================================
// Exact scanline replay for order-dependent draws.  The reference keeps ONE edge array per polygon, partitions the active
// edges to its front at every y-breakpoint (rescan, tess/Polygon.zig:275-299) and re-sorts that prefix by x on every
// scanline with edges and x values moving together (sort, 311-324).  Equal crossings therefore keep the order the previous
// scanlines left them in (pdq sorts slices of up to 12 elements by insertion, i.e. stably; beyond that its tie order is not
// pinned by anything in the reference and insertion order is used here too).  That order decides which crossing the
// non-zero filter (326-353) keeps, so it is visible wherever the crossings of a row do not pair up:
//   * a fill with a dangling edge (kDrawUnpaired): "for (0..filtered.len / 2)" drops the trailing unmatched crossing
//     (multisample.zig:156, supersample.zig, direct.zig) -- row record .x = the x of the dropped crossing (INT_MAX: none);
//   * direct.zig with an unbounded operator (mode 2): every pair clears the rest of its row (direct.zig:112-124), so only the
//     last processed pair survives -- row record = {start, end of that pair, pairs processed, filtered crossings}.
// One thread per such draw (rare, degenerate or special-purpose calls); rescanning on every row is equivalent to the
// reference's breakpoint schedule because a rescan that changes nothing swaps nothing.
// =====================================================================================
__global__ void k_edge_sim(const DevDraw* __restrict__ draws, uint32_t n_draws, const DevSurface* __restrict__ sfcs,
                           const DevEdge* __restrict__ edges, const uint32_t* __restrict__ sp_off, uint32_t* __restrict__ perm_all,
                           int32_t* __restrict__ xs_all, int4* __restrict__ rows_all) {
  const uint32_t di = blockIdx.x * blockDim.x + threadIdx.x;
  if (di >= n_draws) return;
  const DevDraw& d = draws[di];
  if (!d.valid || d.sim_rows <= 0) return;
  const DevEdge* E = edges + sp_off[d.sp_first];
  const uint32_t n = d.n_edges;
  uint32_t* perm = perm_all + d.sim_scratch;
  int32_t* xs = xs_all + d.sim_scratch;
  int4* rows = rows_all + d.sim_base;
  const bool even_odd = d.rule == Z2D_FILL_EVEN_ODD, pair_mode = d.mode == 2;
  const int W = sfcs[d.surface].w;
  for (uint32_t k = 0; k < n; k++) perm[k] = k;
  for (int r = 0; r < d.sim_rows; r++) {
    const double mid = (double)(d.sim_y0 + r) + 0.5;
    uint32_t na = 0;
    for (uint32_t from = 0; from < n; from++) {  // rescan: swap-partition, exactly as the reference does it
      const DevEdge e = E[perm[from]];
      const double top = e.y0 < e.y1 ? e.y0 : e.y1, bottom = e.y0 < e.y1 ? e.y1 : e.y0;
      if (top < mid && bottom >= mid) {
        if (from != na) {
          const uint32_t t = perm[na];
          perm[na] = perm[from];
          perm[from] = t;
        }
        na++;
      }
    }
    for (uint32_t k = 0; k < na; k++) {  // inc
      const DevEdge e = E[perm[k]];
      const double top = e.y0 < e.y1 ? e.y0 : e.y1;
      xs[k] = __double2int_rz(round_half_away(e.x_start + (e.x_inc * (mid - top))));
    }
    for (uint32_t a = 1; a < na; a++) {  // sort (insertion: stable)
      const int32_t key = xs[a];
      const uint32_t pk = perm[a];
      uint32_t b = a;
      while (b > 0 && xs[b - 1] > key) {
        xs[b] = xs[b - 1];
        perm[b] = perm[b - 1];
        b--;
      }
      xs[b] = key;
      perm[b] = pk;
    }
    uint32_t nf = na;
    if (!even_odd) {  // filter: keep 0 -> non-zero and non-zero -> 0 transitions
      int wind = 0;
      nf = 0;
      for (uint32_t from = 0; from < na; from++) {
        const DevEdge e = E[perm[from]];
        const int dir = e.y0 < e.y1 ? -1 : 1;
        xs[nf] = xs[from];
        const bool was_zero = wind == 0;
        wind += dir;
        if (was_zero || wind == 0) nf++;
      }
    }
    if (!pair_mode) {
      rows[r] = make_int4((nf & 1u) ? xs[nf - 1] : INT_MAX, 0, 0, 0);
    } else {  // direct.zig:94-124
      int n_pairs = 0, last_sx = 0, last_ex = 0;
      for (uint32_t p = 0; p < nf / 2; p++) {
        const int sx = max(0, xs[2 * p]);
        if (sx >= W) break;
        last_sx = sx;
        last_ex = clampi(xs[2 * p + 1], sx, W);
        n_pairs++;
      }
      rows[r] = make_int4(last_sx, last_ex, n_pairs, (int)min(nf, 0x7fffffffu));
    }
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

// One list item: draw i on tile row `band` (d, q: the two halves of its DrawBox).
Z2D_D uint4 band_list_item(uint32_t i, int band, const int4 d, const int4 q, const uint32_t* __restrict__ band_off,
                           const uint2* __restrict__ band_xr) {
  uint32_t fl = (uint32_t)q.w, eb = 0u, nbe = 0u;
  int tx0 = d.x, tx1 = d.y;
  if (band >= q.x && band <= q.y) {
    const uint32_t slot = (uint32_t)q.z + (uint32_t)(band - q.x);
    eb = band_off[slot];
    nbe = band_off[slot + 1] - eb;
    fl |= kItemInRows;
    if (!(fl & kItemSpecial)) {  // visit only the tile columns the edges of this row can touch
      const uint2 xr = band_xr[slot];
      tx0 = max(tx0, (int)(0x7fffffffu - xr.x));
      tx1 = min(tx1, (int)xr.y - 1);
    }
  } else if (!(fl & kItemSpecial)) {
    tx0 = 1;  // no edges on this tile row and nothing else to do there: never visited
    tx1 = 0;
  }
  if (tx1 < tx0) {
    tx0 = 0xffff;
    tx1 = 0;
  }
  return make_uint4(i, (uint32_t)tx0 | ((uint32_t)tx1 << 16), eb, min(nbe, 0xffffffu) | (fl << 24));
}

// The draws [b, e) of one surface against one of its tile rows: count them (WRITE = false) or write their list items at o.
// (`bx`: the boxes as int4 pairs, entry of draw i at bx[2 * (i - b0)]: global memory with b0 = 0, or a staged copy of [b, e))
template <bool WRITE>
Z2D_D uint32_t band_list_row(int band, uint32_t b, uint32_t e, const int4* bx, uint32_t b0, uint32_t o, uint4* __restrict__ items,
                             const uint32_t* __restrict__ band_off, const uint2* __restrict__ band_xr) {
    uint32_t n = 0;
    for (uint32_t i = b; i < e; i++) {
      const int4 d = bx[2 * (i - b0)];  // {tx0, tx1, ty0, ty1}
      if (d.x >= 0 && band >= d.z && band <= d.w) {
        if (WRITE) {
          items[o + n] = band_list_item(i, band, d, bx[2 * (i - b0) + 1], band_off, band_xr);
        }
        n++;
      }
    }
    return n;
}

constexpr int kBandListWarps = 8;
template <bool WRITE>
__global__ void __launch_bounds__(kBandListWarps * 32) k_band_lists(const DevSurface* __restrict__ sfcs, uint32_t n_sfc,
                             const uint32_t* __restrict__ work_base, const uint32_t* __restrict__ chunk_base,
                             const DrawBox* __restrict__ boxes, uint32_t* __restrict__ cnt, const uint32_t* __restrict__ off,
                             uint4* __restrict__ items, const uint32_t* __restrict__ band_off, const uint2* __restrict__ band_xr) {
  // blockIdx.x: (surface, chunk of kDrawChunk draws); blockIdx.y * 8 + warp: tile row.  A WARP per (chunk, tile row): its lanes
  // test 32 boxes at a time (staged in shared memory), a ballot keeps the hits in draw order, and in the write pass the hits of a
  // trip fetch their edge ranges and store their items side by side.  (A thread per tile row walking the 256 boxes one after the
  // other, each hit followed by two dependent loads, was 30-170 us per pass whatever the batch size.)
  const uint32_t si = find_surface_by(sfcs, n_sfc, chunk_base, blockIdx.x);
  const DevSurface s = sfcs[si];
  const uint32_t n_draws = s.draw_end - s.draw_begin;
  const uint32_t chunks = (n_draws + kDrawChunk - 1) / kDrawChunk;
  const uint32_t chunk = blockIdx.x - chunk_base[si];
  const uint32_t b = s.draw_begin + chunk * kDrawChunk;
  const uint32_t e = min(b + kDrawChunk, s.draw_end);
  const int band0 = (int)blockIdx.y * kBandListWarps;
  if (band0 >= s.tiles_y) return;
  __shared__ int4 s_box[2 * kDrawChunk];
  const int4* g_box = reinterpret_cast<const int4*>(boxes) + 2 * (size_t)b;
  for (uint32_t k = threadIdx.x; k < 2 * (e - b); k += blockDim.x) s_box[k] = __ldg(g_box + k);
  __syncthreads();
  const int band = band0 + (int)(threadIdx.x >> 5);
  if (band >= s.tiles_y) return;
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t w = work_base[si] + (uint32_t)band * chunks + chunk;
  const uint32_t o = WRITE ? off[w] : 0u;
  uint32_t n = 0;
  for (uint32_t base = 0; base < e - b; base += 32) {
    const uint32_t k = base + lane;
    int4 d = make_int4(-1, 0, 0, 0);
    if (k < e - b) d = s_box[2 * k];  // {tx0, tx1, ty0, ty1}
    const bool hit = d.x >= 0 && band >= d.z && band <= d.w;
    const uint32_t mask = __ballot_sync(0xffffffffu, hit);
    if (WRITE && hit) items[o + n + __popc(mask & ((1u << lane) - 1u))] = band_list_item(b + k, band, d, s_box[2 * k + 1], band_off, band_xr);
    n += __popc(mask);
  }
  if (!WRITE && lane == 0) cnt[w] = n;
}
template __global__ void k_band_lists<false>(const DevSurface*, uint32_t, const uint32_t*, const uint32_t*, const DrawBox*, uint32_t*,
                                             const uint32_t*, uint4*, const uint32_t*, const uint2*);
template __global__ void k_band_lists<true>(const DevSurface*, uint32_t, const uint32_t*, const uint32_t*, const DrawBox*, uint32_t*,
                                            const uint32_t*, uint4*, const uint32_t*, const uint2*);

}  // namespace z2d
#include "pattern.cuh"
#include "raster.cuh"  // blend helpers only: the K4 kernels are compiled in their own translation unit (raster.cu)
namespace z2d {

// =====================================================================================
// K5: surface-level compositor (SurfaceCompositor.run / StrideCompositor.run)
// =====================================================================================
Z2D_D uint32_t comp_pixel(const CompArgs& A, uint32_t raw, int dx, int dy, int sxp, int syp) {
  // (dx,dy): destination pixel; (sxp,syp): matching source-space pixel (compositor.zig:389-436).  Patterns are evaluated at
  // the CANVAS row (a band surface holds canvas rows y_origin ...).
  const int py = dy + A.y_origin;
  if (A.precision == Z2D_PRECISION_INTEGER) {
    RGBA16 d{0, 0, 0, 0}, s{0, 0, 0, 0};
    for (uint32_t k = 0; k < A.n_ops; k++) {
      const CompOp& o = A.ops[k];
      s = o.has_src ? src_int(o.src, A.T, dx, py, (size_t)o.src.sw * (size_t)syp + (size_t)sxp) : d;
      d = o.has_dst ? src_int(o.dst, A.T, dx, py, (size_t)o.dst.sw * (size_t)dy + (size_t)dx) : raw_to_rgba16(A.fmt, raw);
      d = int_op(o.op, d, s);
    }
    return rgba16_to_raw(A.fmt, d);
  }
  RGBAF d{0, 0, 0, 0}, s{0, 0, 0, 0};
  for (uint32_t k = 0; k < A.n_ops; k++) {
    const CompOp& o = A.ops[k];
    s = o.has_src ? src_float(o.src, A.T, dx, py, (size_t)o.src.sw * (size_t)syp + (size_t)sxp) : d;
    d = o.has_dst ? src_float(o.dst, A.T, dx, py, (size_t)o.dst.sw * (size_t)dy + (size_t)dx) : decode_raw(raw_to_rgba16(A.fmt, raw));
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

// ---- fast path: ONE operator, no dst override, 32-bit destination, full contiguous rows,
// source = single pixel or a same-width 32-bit surface.  The operator is a template constant
// (the 28-way switch folds away), channel positions are uniform registers, u8->f32 decode is
// a shared-memory table of the exact x/255.0f quotients, and each thread streams 2 x 128-bit
// vectors per iteration.  HBM-bound: 8 B per pixel (read + write), +4 B with a surface source.
template <int PREC, int OP, bool SURF>
Z2D_D uint32_t fast_px(const Fmt32& fd, const Fmt32& fs, const float* __restrict__ lut, uint32_t raw, RGBA16 s_const, RGBAF sf_const,
                       uint32_t sraw) {
  if (PREC == Z2D_PRECISION_INTEGER && OP == Z2D_OP_SRC_OVER) {
    // packed two-lane integer src_over (raster.cuh): half the ALU work of the per-channel form, identical results
    const uint32_t amask = fd.has_a ? 0xffffffffu : 0x00ffffffu;
    if (!SURF) return src_over_x4(raw, src_lanes(fd, s_const, 255, false), amask);
    if (fs.rs == fd.rs) {  // same channel order: the source word splits straight into lanes (no alpha channel: sa = 255)
      const uint32_t sw = fs.has_a ? sraw : (sraw | 0xff000000u);
      return src_over_x4(raw, make_uint2(sw & 0x00ff00ffu, (sw >> 8) & 0x00ff00ffu), amask);
    }
  }
  RGBA16 d = unpack32(fd, raw);
  if (PREC == Z2D_PRECISION_INTEGER) {
    RGBA16 s = SURF ? unpack32(fs, sraw) : s_const;
    return pack32(fd, int_op((uint32_t)OP, d, s));
  }
  RGBAF sf = sf_const;
  if (SURF) {
    RGBA16 s = unpack32(fs, sraw);
    sf = {lut[s.r], lut[s.g], lut[s.b], lut[s.a]};
  }
  RGBAF df{lut[d.r], lut[d.g], lut[d.b], lut[d.a]};
  return pack32(fd, encode_raw(float_op((uint32_t)OP, df, sf)));
}

template <int PREC, int OP, bool SURF>
Z2D_D void fast_loop(const CompArgs& A, const float* __restrict__ lut) {
  const Fmt32 fd = fmt32_of(A.fmt), fs = fmt32_of(A.ops[0].src.sfmt);
  const RGBA16 sc = unpack_rgba(A.ops[0].src.px_rgba);
  const RGBAF sfc{lut[sc.r], lut[sc.g], lut[sc.b], lut[sc.a]};
  const size_t n4 = ((size_t)A.scan_w * (size_t)A.rows) >> 2;
  uint4* __restrict__ dst = reinterpret_cast<uint4*>(A.data + ((size_t)A.dst_start_y * (size_t)A.w) * 4);
  const uint4* __restrict__ src = SURF ? reinterpret_cast<const uint4*>(A.ops[0].src.sdata + ((size_t)A.src_start_y * (size_t)A.ops[0].src.sw) * 4) : nullptr;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  constexpr bool kNoRead = OP == Z2D_OP_CLEAR || OP == Z2D_OP_SRC;
#ifndef Z2D_COMP_STREAMS
#define Z2D_COMP_STREAMS 2
#endif
#ifdef Z2D_COMP_CS
#define Z2D_LD(p) __ldcs(p)
#define Z2D_ST(p, v) __stcs(p, v)
#else
#define Z2D_LD(p) (*(p))
#define Z2D_ST(p, v) (*(p) = (v))
#endif
  for (; i + (Z2D_COMP_STREAMS - 1) * stride < n4; i += Z2D_COMP_STREAMS * stride) {  // independent 128-bit streams in flight per thread
    uint4 v[Z2D_COMP_STREAMS], sv[Z2D_COMP_STREAMS];
#pragma unroll
    for (int k = 0; k < Z2D_COMP_STREAMS; k++) {
      v[k] = kNoRead ? make_uint4(0, 0, 0, 0) : Z2D_LD(dst + i + k * stride);  // clear / src never look at dst: write only
      sv[k] = SURF ? __ldg(src + i + k * stride) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int k = 0; k < Z2D_COMP_STREAMS; k++) {
      v[k].x = fast_px<PREC, OP, SURF>(fd, fs, lut, v[k].x, sc, sfc, sv[k].x);
      v[k].y = fast_px<PREC, OP, SURF>(fd, fs, lut, v[k].y, sc, sfc, sv[k].y);
      v[k].z = fast_px<PREC, OP, SURF>(fd, fs, lut, v[k].z, sc, sfc, sv[k].z);
      v[k].w = fast_px<PREC, OP, SURF>(fd, fs, lut, v[k].w, sc, sfc, sv[k].w);
    }
#pragma unroll
    for (int k = 0; k < Z2D_COMP_STREAMS; k++) Z2D_ST(dst + i + k * stride, v[k]);
  }
  for (; i < n4; i += stride) {
    uint4 a = kNoRead ? make_uint4(0, 0, 0, 0) : Z2D_LD(dst + i);
    uint4 sa = make_uint4(0, 0, 0, 0);
    if (SURF) sa = __ldg(src + i);
    a.x = fast_px<PREC, OP, SURF>(fd, fs, lut, a.x, sc, sfc, sa.x);
    a.y = fast_px<PREC, OP, SURF>(fd, fs, lut, a.y, sc, sfc, sa.y);
    a.z = fast_px<PREC, OP, SURF>(fd, fs, lut, a.z, sc, sfc, sa.z);
    a.w = fast_px<PREC, OP, SURF>(fd, fs, lut, a.w, sc, sfc, sa.w);
    Z2D_ST(dst + i, a);
  }
}

template <int PREC, bool SURF>
__global__ void __launch_bounds__(256) k_composite_fast(const __grid_constant__ CompArgs A) {
  __shared__ float lut[256];
  lut[threadIdx.x] = (float)threadIdx.x / 255.0f;  // color.zig:243-250 decodeRGBARaw, exact quotients
  __syncthreads();
  switch (A.ops[0].op) {
#define Z2D_OPCASE(op) \
  case op: fast_loop<PREC, op, SURF>(A, lut); break;
    Z2D_OPCASE(Z2D_OP_CLEAR) Z2D_OPCASE(Z2D_OP_SRC) Z2D_OPCASE(Z2D_OP_DST) Z2D_OPCASE(Z2D_OP_SRC_OVER) Z2D_OPCASE(Z2D_OP_DST_OVER)
    Z2D_OPCASE(Z2D_OP_SRC_IN) Z2D_OPCASE(Z2D_OP_DST_IN) Z2D_OPCASE(Z2D_OP_SRC_OUT) Z2D_OPCASE(Z2D_OP_DST_OUT) Z2D_OPCASE(Z2D_OP_SRC_ATOP)
    Z2D_OPCASE(Z2D_OP_DST_ATOP) Z2D_OPCASE(Z2D_OP_XOR) Z2D_OPCASE(Z2D_OP_PLUS) Z2D_OPCASE(Z2D_OP_MULTIPLY) Z2D_OPCASE(Z2D_OP_SCREEN)
    Z2D_OPCASE(Z2D_OP_OVERLAY) Z2D_OPCASE(Z2D_OP_DARKEN) Z2D_OPCASE(Z2D_OP_LIGHTEN) Z2D_OPCASE(Z2D_OP_COLOR_DODGE)
    Z2D_OPCASE(Z2D_OP_COLOR_BURN) Z2D_OPCASE(Z2D_OP_HARD_LIGHT) Z2D_OPCASE(Z2D_OP_SOFT_LIGHT) Z2D_OPCASE(Z2D_OP_DIFFERENCE)
    Z2D_OPCASE(Z2D_OP_EXCLUSION) Z2D_OPCASE(Z2D_OP_HUE) Z2D_OPCASE(Z2D_OP_SATURATION) Z2D_OPCASE(Z2D_OP_COLOR) Z2D_OPCASE(Z2D_OP_LUMINOSITY)
#undef Z2D_OPCASE
    default: break;
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

}  // namespace z2d
#include "composite.cuh"
#include "smallbatch.cuh"
namespace z2d {

// ------------------------------------------------------------------------- launchers
static inline unsigned blocks_for(size_t n, unsigned threads) { return (unsigned)((n + threads - 1) / threads); }

}  // namespace z2d
#include "slowpath.cuh"
namespace z2d {

void launch_hairline(const DevSurface* sfcs, const DevDraw* draws, uint32_t draw_index, const z2d_node* nodes, uint32_t node_begin,
                     uint32_t node_end, const double* dashes, const GradTables& T, cudaStream_t st) {
  k_hairline<<<1, 32, 0, st>>>(sfcs, draws, draw_index, nodes, node_begin, node_end, dashes, T);
}
void launch_flatten_count(const DevSubPath* sps, uint32_t n_sp, const z2d_node* nodes, DevDraw* draws, uint32_t* sp_count, const void* pens,
                          const double* dashes, const uint32_t* order, cudaStream_t st) {
  if (n_sp) k_flatten_count<<<blocks_for(n_sp, Z2D_FLATTEN_THREADS), Z2D_FLATTEN_THREADS, 0, st>>>(sps, n_sp, nodes, draws, sp_count, (const PenV*)pens, dashes, order);
}
void launch_flatten_emit(const DevSubPath* sps, uint32_t n_sp, const z2d_node* nodes, const DevDraw* draws, const uint32_t* sp_off,
                         DevEdge* edges, uint32_t* edge_draw, const void* pens, const double* dashes, const uint32_t* order, cudaStream_t st) {
  if (n_sp) k_flatten_emit<<<blocks_for(n_sp, Z2D_FLATTEN_THREADS), Z2D_FLATTEN_THREADS, 0, st>>>(sps, n_sp, nodes, draws, sp_off, edges, edge_draw, (const PenV*)pens, dashes, order);
}
void launch_stroke_units(const DevSubPath* sps, uint32_t n_sp, const z2d_node* nodes, DevDraw* draws, const void* pens, const double* dashes,
                         const uint32_t* order, void* units, uint32_t unit_cap, void* links, uint32_t link_cap, void* ports, uint32_t* ctr,
                         DevEdge* edges, uint32_t* edge_draw, uint32_t edge_cap, cudaStream_t st) {
  if (!n_sp) return;
  k_stroke_walk<<<blocks_for((size_t)n_sp * Z2D_WALK_SPREAD, Z2D_WALK_THREADS), Z2D_WALK_THREADS, 0, st>>>(sps, n_sp, nodes, draws, (const PenV*)pens, dashes, order, (StrokeUnit*)units, unit_cap,
                                                     (StrokeLink*)links, link_cap, ctr);
  if (unit_cap)
    k_stroke_units<<<blocks_for(unit_cap, 128), 128, 0, st>>>((const StrokeUnit*)units, unit_cap, ctr, draws, (const PenV*)pens, dashes, (Pt*)ports,
                                                              edges, edge_draw, edge_cap);
  if (link_cap)
    k_stroke_links<<<blocks_for(link_cap, 256), 256, 0, st>>>((const StrokeLink*)links, link_cap, unit_cap, ctr, draws, (const Pt*)ports, edges, edge_draw,
                                                              edge_cap);
}
void launch_flatten_nodes(bool emit, const DevSubPath* sps, uint32_t n_sp, uint32_t* node_sp, uint32_t n_nodes, const z2d_node* nodes,
                          DevDraw* draws, uint32_t* counts, const uint32_t* offs, DevEdge* edges, uint32_t* edge_draw,
                          uint32_t* curve_list, cudaStream_t st) {
  // node_sp: n_nodes entries; curve_list: n_nodes + 1 entries, the last one is the number of curves
  if (!n_nodes || !n_sp) return;
  uint32_t* n_curves = curve_list + n_nodes;
  const uint32_t nb = blocks_for(n_nodes, 128);
  if (!emit) {
    cudaMemsetAsync(node_sp, 0xff, (size_t)n_nodes * 4, st);
    cudaMemsetAsync(n_curves, 0, 4, st);
    k_mark_nodes<<<blocks_for(n_sp, 128), 128, 0, st>>>(sps, n_sp, node_sp);
    k_flatten_nodes<false, false><<<nb, 128, 0, st>>>(sps, node_sp, n_nodes, nodes, draws, counts, nullptr, nullptr, nullptr, curve_list, n_curves);
    k_flatten_nodes<false, true><<<nb, 128, 0, st>>>(sps, node_sp, n_nodes, nodes, draws, counts, nullptr, nullptr, nullptr, curve_list, n_curves);
  } else {
    k_flatten_nodes<true, false><<<nb, 128, 0, st>>>(sps, node_sp, n_nodes, nodes, draws, nullptr, offs, edges, edge_draw, curve_list, n_curves);
    k_flatten_nodes<true, true><<<nb, 128, 0, st>>>(sps, node_sp, n_nodes, nodes, draws, nullptr, offs, edges, edge_draw, curve_list, n_curves);
  }
}
// single-pass node-parallel flattening: bounds (first call, before the scan of `counts`), then edges into the pool
void launch_node_bounds(const DevSubPath* sps, uint32_t n_sp, uint32_t* node_sp, uint32_t n_nodes, const z2d_node* nodes, const DevDraw* draws,
                        uint32_t* counts, uint32_t* curve_list, cudaStream_t st) {
  if (!n_nodes || !n_sp) return;
  uint32_t* n_curves = curve_list + n_nodes;
  cudaMemsetAsync(node_sp, 0xff, (size_t)n_nodes * 4, st);
  cudaMemsetAsync(n_curves, 0, 4, st);
  k_mark_nodes<<<blocks_for(n_sp, 128), 128, 0, st>>>(sps, n_sp, node_sp);
  k_node_bounds<<<blocks_for(n_nodes, 128), 128, 0, st>>>(sps, node_sp, n_nodes, nodes, draws, counts, curve_list, n_curves);
}
void launch_flatten_nodes_pool(const DevSubPath* sps, const uint32_t* node_sp, uint32_t n_nodes, const z2d_node* nodes, DevDraw* draws,
                               const uint32_t* node_offs, uint32_t* ctr, DevEdge* edges, uint32_t* edge_draw, uint32_t edge_cap,
                               const uint32_t* curve_list, cudaStream_t st) {
  if (!n_nodes) return;
  const uint32_t* n_curves = curve_list + n_nodes;
  const uint32_t nb = blocks_for(n_nodes, 128);
  k_take_node_range<<<1, 1, 0, st>>>(ctr, node_offs, n_nodes);
  k_flatten_nodes_pool<false><<<nb, 128, 0, st>>>(sps, node_sp, n_nodes, nodes, draws, node_offs, ctr, edges, edge_draw, edge_cap, curve_list, n_curves);
  k_flatten_nodes_pool<true><<<nb, 128, 0, st>>>(sps, node_sp, n_nodes, nodes, draws, node_offs, ctr, edges, edge_draw, edge_cap, curve_list, n_curves);
}
void launch_setup_draws(DevDraw* draws, uint32_t n, const DevSurface* sfcs, uint32_t* draw_bands, DrawBox* boxes, unsigned long long* counters, cudaStream_t st) {
  if (n) k_setup_draws<<<blocks_for(n, 128), 128, 0, st>>>(draws, n, sfcs, draw_bands, boxes, counters);
}
void launch_expand_draws(const DrawIn* in, const StrokeIn* strokes, const DevSrc* srcs, DevDraw* draws, uint32_t n, cudaStream_t st) {
  if (n) k_expand_draws<<<(n + 127) / 128, 128, 0, st>>>(in, strokes, srcs, draws, n);
}
void launch_assign_band_base(DevDraw* draws, uint32_t n, const uint32_t* band_off, DrawHot* hots, DrawBox* boxes, const DevSurface* sfcs,
                             cudaStream_t st) {
  if (n) k_assign_band_base<<<blocks_for(n, 256), 256, 0, st>>>(draws, n, band_off, hots, boxes, sfcs);
}
void launch_bin_count(const DevEdge* edges, const uint32_t* edge_draw, uint32_t n, const DevDraw* draws, uint32_t* band_count, cudaStream_t st) {
  if (n) k_bin_count<<<blocks_for(n, 256), 256, 0, st>>>(edges, edge_draw, n, draws, band_count);
}
void launch_bin_scatter(const DevEdge* edges, const uint32_t* edge_draw, uint32_t n, const DevDraw* draws, const uint32_t* band_off,
                        uint32_t* band_cursor, DevEdge* band_edges, int4* band_hdr, uint2* band_xr, cudaStream_t st) {
  if (n) k_bin_scatter<<<blocks_for(n, 256), 256, 0, st>>>(edges, edge_draw, n, draws, band_off, band_cursor, band_edges, band_hdr, band_xr);
}
void launch_band_lists(bool write, const DevSurface* sfcs, uint32_t n_sfc, const uint32_t* work_base, const uint32_t* chunk_base,
                       uint32_t n_chunks, const DrawBox* boxes, uint32_t* cnt, const uint32_t* off, uint4* items, const uint32_t* band_off,
                       const uint2* band_xr, uint32_t max_tiles_y, cudaStream_t st) {
  if (!n_chunks) return;
  const dim3 grid(n_chunks, (max_tiles_y + kBandListWarps - 1) / kBandListWarps);
  if (write)
    k_band_lists<true><<<grid, kBandListWarps * 32, 0, st>>>(sfcs, n_sfc, work_base, chunk_base, boxes, cnt, off, items, band_off, band_xr);
  else
    k_band_lists<false><<<grid, kBandListWarps * 32, 0, st>>>(sfcs, n_sfc, work_base, chunk_base, boxes, cnt, off, items, band_off, band_xr);
}
void launch_edge_sim(const DevDraw* draws, uint32_t n_draws, const DevSurface* sfcs, const DevEdge* edges, const uint32_t* sp_off,
                     uint32_t* perm, int32_t* xs, int4* rows, cudaStream_t st) {
  if (n_draws) k_edge_sim<<<blocks_for(n_draws, 64), 64, 0, st>>>(draws, n_draws, sfcs, edges, sp_off, perm, xs, rows);
}
// text.show on the device: one block per glyph instance, threads over the nodes of its cached outline; every point goes through
// Transformation.userToDevice (Transformation.zig:194-206: ax * x + by * y, then + tx; no contraction) into the batch's node array.
__global__ void __launch_bounds__(128) k_expand_glyphs(const GlyphInst* __restrict__ inst, const z2d_node* __restrict__ cache,
                                                       z2d_node* __restrict__ nodes) {
  const GlyphInst g = inst[blockIdx.x];
  for (uint32_t k = threadIdx.x; k < g.n_nodes; k += blockDim.x) {
    z2d_node nd = cache[g.src + k];
    const int np = nd.tag == Z2D_NODE_CURVE_TO ? 3 : (nd.tag == Z2D_NODE_CLOSE_PATH ? 0 : 1);
    for (int q = 0; q < np; q++) {
      const double x = nd.p[2 * q], y = nd.p[2 * q + 1];
      const double dx = g.m[0] * x + g.m[1] * y, dy = g.m[2] * x + g.m[3] * y;
      nd.p[2 * q] = dx + g.m[4];
      nd.p[2 * q + 1] = dy + g.m[5];
    }
    nodes[g.dst + k] = nd;
  }
}
void launch_expand_glyphs(const GlyphInst* inst, uint32_t n, const z2d_node* cache, z2d_node* nodes, cudaStream_t st) {
  if (n) k_expand_glyphs<<<n, 128, 0, st>>>(inst, cache, nodes);
}
void launch_small_batch(const SmallArgs& A, cudaStream_t st) { k_small_batch<<<1, kSmallThreads, 0, st>>>(A); }
void launch_composite(const CompArgs& A, int sm_count, cudaStream_t st) {
  const size_t n = (size_t)A.scan_w * (size_t)A.rows;
  if (n == 0) return;
  const bool vec = A.fmt <= Z2D_FMT_RGBA && A.dst_start_x == 0 && A.scan_w == A.w && (A.scan_w & 3) == 0;
  const size_t items = vec ? (n >> 2) : n;
  unsigned blocks = (unsigned)((items + 255) / 256);
  const unsigned cap = (unsigned)sm_count * 8u * 4u;  // grid-stride: a few waves of 8 resident CTAs per SM
  if (blocks > cap) blocks = cap;
  const CompOp& o0 = A.ops[0];
  const bool full_rows = A.dst_start_x == 0 && A.scan_w == A.w;
  const bool single = A.n_ops == 1 && !o0.has_dst && o0.has_src;
  if (single && full_rows && A.fmt > Z2D_FMT_RGBA && o0.src.kind == Z2D_PARAM_PIXEL) {  // packed / 8-bit alpha, single pixel: byte table
    const size_t chunks = (n * (size_t)fmt_bits(A.fmt) / 8 + 15) / 16 + 1;
    unsigned lb = (unsigned)std::min<size_t>((chunks + 255) / 256, (size_t)sm_count * 8u);
    k_composite_lut<<<lb ? lb : 1u, 256, 0, st>>>(A);
    return;
  }
  if (single && full_rows && (o0.src.kind == Z2D_PARAM_GRADIENT || o0.src.kind == Z2D_PARAM_DITHER) && A.max_stops <= (uint32_t)kGenMaxStops) {
    const int bits = fmt_bits(A.fmt);
    const size_t chunks = (n * (size_t)bits / 8 + 15) / 16 + 1;
    unsigned gb = (unsigned)std::min<size_t>((chunks + 255) / 256, (size_t)sm_count * 8u);
    if (!gb) gb = 1;
    const bool dith = o0.src.kind == Z2D_PARAM_DITHER, flt = A.precision != Z2D_PRECISION_INTEGER;
    const int fc = bits == 32 ? 0 : bits == 8 ? 1 : 2;
#define Z2D_GEN(FC, PR, DI) k_composite_gen<FC, PR, DI><<<gb, 256, 0, st>>>(A)
#define Z2D_GEN_FC(FC)                                                                    \
  if (!flt && !dith) Z2D_GEN(FC, Z2D_PRECISION_INTEGER, false);                            \
  else if (!flt) Z2D_GEN(FC, Z2D_PRECISION_INTEGER, true);                                 \
  else if (!dith) Z2D_GEN(FC, Z2D_PRECISION_FLOAT, false);                                 \
  else Z2D_GEN(FC, Z2D_PRECISION_FLOAT, true);
    if (fc == 0) { Z2D_GEN_FC(0) } else if (fc == 1) { Z2D_GEN_FC(1) } else { Z2D_GEN_FC(2) }
#undef Z2D_GEN_FC
#undef Z2D_GEN
    return;
  }
  const bool one = vec && single;
  const bool fast_px_src = one && o0.src.kind == Z2D_PARAM_PIXEL;
  const bool fast_sfc_src = one && o0.src.kind == Z2D_PARAM_SURFACE && o0.src.sfmt <= Z2D_FMT_RGBA && o0.src.sw == A.w && A.src_start_x == 0;
  if (fast_px_src || fast_sfc_src) {
    unsigned fb = (unsigned)(((items + 1) / 2 + 255) / 256);
#ifndef Z2D_COMP_CTAS
#define Z2D_COMP_CTAS 8
#endif
    const unsigned fcap = (unsigned)sm_count * Z2D_COMP_CTAS;  // persistent-style grid: resident CTAs per SM
    if (fb > fcap) fb = fcap;
    if (fb == 0) fb = 1;
    if (A.precision == Z2D_PRECISION_INTEGER) {
      if (fast_px_src) k_composite_fast<Z2D_PRECISION_INTEGER, false><<<fb, 256, 0, st>>>(A);
      else k_composite_fast<Z2D_PRECISION_INTEGER, true><<<fb, 256, 0, st>>>(A);
    } else {
      if (fast_px_src) k_composite_fast<Z2D_PRECISION_FLOAT, false><<<fb, 256, 0, st>>>(A);
      else k_composite_fast<Z2D_PRECISION_FLOAT, true><<<fb, 256, 0, st>>>(A);
    }
    return;
  }
  if (vec)
    k_composite_v4<<<blocks, 256, 0, st>>>(A);
  else
    k_composite_px<<<blocks, 256, 0, st>>>(A);
}
// ---------------------------------------------------------------- export (export_png.zig:150-373)
// The scanline bytes the reference hands to zlib: RGBA/ARGB de-multiplied in integer space (pixel_vector.zig:27-49) and written
// R,G,B,A; RGB/XRGB written R,G,B; alpha8 as 8-bit grey; alpha4/2/1 re-packed MSB-first with every row starting on a byte
// (the surface packs LSB-first with rows not byte-aligned).  With a colour profile each colour channel goes through
// round(255 * pow(c / 255, 1 / gamma)) -- 256 possible inputs, so a table built on the host.  One item = one pixel (>= 8 bits) or
// one output byte (packed greys).
__global__ void __launch_bounds__(256) k_export_rows(ExportArgs A) {
  const size_t total = (size_t)A.items_per_row * (size_t)A.h;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t y = i / A.items_per_row;
    const uint32_t x = (uint32_t)(i - y * A.items_per_row);
    uint8_t* row = A.out + y * A.row_bytes;
    if (x == 0 && A.filter_byte) row[0] = 0;  // scanline header: filter type none
    row += A.filter_byte;
    if (A.fmt <= Z2D_FMT_RGBA) {
      RGBA16 v = raw_to_rgba16(A.fmt, ((const uint32_t*)A.data)[y * (size_t)A.w + x]);
      const bool alpha = A.fmt == Z2D_FMT_ARGB || A.fmt == Z2D_FMT_RGBA;
      if (alpha) {
        const int d = v.a > 1 ? v.a : 1;
        v.r = v.a == 0 ? 0 : v.r * 255 / d;
        v.g = v.a == 0 ? 0 : v.g * 255 / d;
        v.b = v.a == 0 ? 0 : v.b * 255 / d;
      }
      if (A.gamma) {  // an unpremultiplied-looking pixel (c > a) can de-multiply past 255; the reference's u8 cast would trap
        v.r = A.gamma[v.r & 255];
        v.g = A.gamma[v.g & 255];
        v.b = A.gamma[v.b & 255];
      }
      if (alpha) {
        uint8_t* o = row + (size_t)x * 4;
        o[0] = (uint8_t)v.r; o[1] = (uint8_t)v.g; o[2] = (uint8_t)v.b; o[3] = (uint8_t)v.a;
      } else {
        uint8_t* o = row + (size_t)x * 3;
        o[0] = (uint8_t)v.r; o[1] = (uint8_t)v.g; o[2] = (uint8_t)v.b;
      }
    } else if (A.fmt == Z2D_FMT_ALPHA8) {
      row[x] = A.data[y * (size_t)A.w + x];
    } else {
      const int bits = fmt_bits(A.fmt), per = 8 / bits;
      uint32_t b = 0;
      for (int k = 0; k < per; ++k) {
        const uint32_t px = x * per + k;
        if (px >= (uint32_t)A.w) break;
        b |= load_raw(A.data, A.fmt, y * (size_t)A.w + px) << (8 - bits - k * bits);
      }
      row[x] = (uint8_t)b;
    }
  }
}

void launch_export(const ExportArgs& A, int sm_count, cudaStream_t st) {
  const size_t total = (size_t)A.items_per_row * (size_t)A.h;
  if (total == 0) return;
  size_t blocks = (total + 255) / 256;
  const size_t cap = (size_t)sm_count * 8;
  if (blocks > cap) blocks = cap;
  k_export_rows<<<(unsigned)blocks, 256, 0, st>>>(A);
}

// Surface.downsample (surface.zig:447-469, 687-709): 4x4 box average with truncation, every format (`T.average`: the sum of the
// 16 stored samples / 16 per channel; RGB / XRGB write padding 0).  Out of place (the reference compacts in place, sequentially:
// same result, every output is written before the inputs of later outputs are reached).  One thread per output pixel.
__global__ void k_downsample(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, uint32_t fmt, int w_in, int w_out, int h_out) {
  const size_t n = (size_t)w_out * (size_t)h_out;
  const int bits = fmt_bits(fmt);
  for (size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (size_t)gridDim.x * blockDim.x) {
    const size_t y = o / (size_t)w_out, x = o - y * (size_t)w_out;
    uint32_t acc[4] = {0u, 0u, 0u, 0u};
    for (int i = 0; i < 4; i++)
      for (int j = 0; j < 4; j++) {
        const uint32_t raw = load_raw(src, fmt, (y * 4 + (size_t)i) * (size_t)w_in + (x * 4 + (size_t)j));
        if (bits == 32) {
          acc[0] += raw & 255u; acc[1] += (raw >> 8) & 255u; acc[2] += (raw >> 16) & 255u; acc[3] += raw >> 24;
        } else {
          acc[0] += raw;
        }
      }
    uint32_t out;
    if (bits == 32) {
      out = (acc[0] / 16u) | ((acc[1] / 16u) << 8) | ((acc[2] / 16u) << 16) | ((acc[3] / 16u) << 24);
      if (fmt == Z2D_FMT_RGB || fmt == Z2D_FMT_XRGB) out &= 0x00ffffffu;
    } else {
      out = acc[0] / 16u;
    }
    store_raw(dst, fmt, o, out);
  }
}
void launch_downsample(const uint8_t* src, uint8_t* dst, uint32_t fmt, int w_in, int w_out, int h_out, cudaStream_t st) {
  const size_t n = (size_t)w_out * (size_t)h_out;
  if (!n) return;
  unsigned blocks = (unsigned)std::min<size_t>((n + 255) / 256, 148u * 16u);
  k_downsample<<<blocks, 256, 0, st>>>(src, dst, fmt, w_in, w_out, h_out);
}

void launch_paint(uint8_t* data, uint32_t fmt, size_t n_px, uint32_t raw, cudaStream_t st) {
  unsigned blocks = (unsigned)((n_px + 255) / 256);
  if (blocks > 148u * 16u) blocks = 148u * 16u;
  if (blocks == 0) blocks = 1;
  k_paint<<<blocks, 256, 0, st>>>(data, fmt, n_px, raw);
}
void launch_put_pixel(uint8_t* data, uint32_t fmt, size_t idx, uint32_t raw, cudaStream_t st) { k_put_pixel<<<1, 1, 0, st>>>(data, fmt, idx, raw); }

}  // namespace z2d
