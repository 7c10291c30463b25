// Small batches (a handful of draw calls: one SVG icon, one logo, one text run) in TWO launches instead of ~30.
//
// The sized pipeline (z2d_lib.cu run_pipeline) is a fixed sequence of ~30 launches with two host read-backs (the host
// allocates the edge / binned-edge buffers from totals the device computed): ~290 us per 5-fill scene, all of it fixed cost.
// For batches of <= kSmallMaxDraws fills k_small_batch runs every stage before the raster kernel -- expand, flatten (count,
// scan, emit), per-draw setup, edge binning (count, scan, scatter) and the tile-row lists -- as phases of ONE CTA separated by
// __syncthreads, with the stage functions of kernels.cu, into buffers of fixed capacity.  Nothing comes back to the host: the
// raster kernel is launched right behind it over the batch's tiles.  If a capacity would be exceeded, or the batch needs the
// scanline replay of k_edge_sim (a dangling edge), the kernel raises `out[0]`, the raster kernel sees it and does nothing, and
// the library redoes the batch with the sized pipeline the next time it touches the context (inputs are still resident).
#pragma once

namespace z2d {

// exclusive scan of v[0..n) in place, v[n] = total; the whole CTA cooperates (sh: blockDim.x words)
Z2D_D uint32_t block_scan_inplace(uint32_t* v, uint32_t n, uint32_t* sh) {
  const uint32_t T = blockDim.x, t = threadIdx.x;
  const uint32_t per = (n + T - 1) / T;
  const uint32_t b = min(t * per, n), e = min(b + per, n);
  uint32_t s = 0;
  for (uint32_t i = b; i < e; i++) s += v[i];
  sh[t] = s;
  __syncthreads();
  for (uint32_t off = 1; off < T; off <<= 1) {
    const uint32_t x = t >= off ? sh[t - off] : 0u;
    __syncthreads();
    sh[t] += x;
    __syncthreads();
  }
  uint32_t run = sh[t] - s;
  const uint32_t total = sh[T - 1];
  for (uint32_t i = b; i < e; i++) {
    const uint32_t x = v[i];
    v[i] = run;
    run += x;
  }
  if (t == 0) v[n] = total;
  __syncthreads();
  return total;
}

__global__ void __launch_bounds__(kSmallThreads, 1) k_small_batch(const __grid_constant__ SmallArgs A) {
  __shared__ uint32_t sh[kSmallThreads];
  __shared__ uint32_t s_abort;
  const uint32_t t = threadIdx.x, T = blockDim.x;
  if (t == 0) s_abort = 0u;
  if (t < 8) A.out[t] = 0u;
  for (uint32_t i = t; i < 8; i += T) A.counters[i] = 0ull;
  // ---- K0: draw records
  for (uint32_t i = t; i < A.n_draws; i += T) expand_draw(i, A.draws_in, A.strokes, A.srcs, A.draws);
  for (uint32_t i = t; i < A.n_nodes; i += T) A.node_sp[i] = 0xffffffffu;
  __syncthreads();
  // ---- K1, count: sub-path -> node map, then sequential sub-paths (thread each) and node-parallel nodes (thread each)
  for (uint32_t i = t; i < A.n_sp; i += T) {
    const DevSubPath sp = A.sps[i];
    if (i == 0 || A.sps[i - 1].draw != sp.draw) A.draws[sp.draw].sp_first = i;
    if (sp.flags & kSpNodeParallel)
      for (uint32_t k = sp.node_begin; k < sp.node_end; k++) A.node_sp[k] = i;
  }
  __syncthreads();
  uint32_t* cnt_sp = A.cnt;
  uint32_t* cnt_nd = A.cnt + A.n_sp;
  for (uint32_t i = t; i < A.n_sp; i += T) {
    const DevSubPath sp = A.sps[i];
    if (sp.flags & kSpNodeParallel) {
      cnt_sp[i] = 0;
      continue;
    }
    DevDraw& d = A.draws[sp.draw];
    EdgeSink<false> sink;
    sink.scale = d.scale;
    fill_subpath<false>(A.nodes, sp.node_begin, sp.node_end, d.tolerance, sink);
    cnt_sp[i] = sink.n;
    if (sink.unpaired && sink.n > 0) atomicOr(&d.flags, kDrawUnpaired);
    if (sink.n > 0) {
      ext_commit(d, sink.top, sink.bottom, sink.left, sink.right);
      atomicAdd(&d.n_edges, sink.n);
    }
  }
  for (uint32_t i = t; i < A.n_nodes; i += T) {
    const uint32_t spi = A.node_sp[i];
    const z2d_node nd = A.nodes[i];
    if (spi == 0xffffffffu || nd.tag == Z2D_NODE_MOVE_TO) {
      cnt_nd[i] = 0;
      continue;
    }
    flatten_node<false>(i, A.sps[spi], nd, A.nodes, A.draws, cnt_nd, nullptr, nullptr, nullptr);
  }
  __syncthreads();
  const uint32_t n_cnt = A.n_sp + A.n_nodes;
  const uint32_t n_edges = block_scan_inplace(A.cnt, n_cnt, sh);
  if (n_edges > kSmallEdgeCap) {
    if (t == 0) A.out[0] = 1u;
    return;
  }
  // ---- K1, emit
  for (uint32_t i = t; i < A.n_sp; i += T) {
    const DevSubPath sp = A.sps[i];
    if (sp.flags & kSpNodeParallel) continue;
    const DevDraw& d = A.draws[sp.draw];
    EdgeSink<true> sink;
    sink.scale = d.scale;
    sink.out = A.edges + cnt_sp[i];
    sink.out_draw = A.edge_draw + cnt_sp[i];
    sink.draw = sp.draw;
    sink.limit = A.cnt[i + 1] - A.cnt[i];
    fill_subpath<true>(A.nodes, sp.node_begin, sp.node_end, d.tolerance, sink);
  }
  for (uint32_t i = t; i < A.n_nodes; i += T) {
    const uint32_t spi = A.node_sp[i];
    const z2d_node nd = A.nodes[i];
    if (spi == 0xffffffffu || nd.tag == Z2D_NODE_MOVE_TO) continue;
    flatten_node<true>(i, A.sps[spi], nd, A.nodes, A.draws, nullptr, cnt_nd, A.edges, A.edge_draw);
  }
  __syncthreads();
  // ---- K2: regions, tile boxes, (draw, tile-row) slots
  for (uint32_t i = t; i < A.n_draws; i += T) {
    setup_draw(i, A.draws, A.sfcs, A.draw_bands, A.boxes, A.counters);
    if (A.draws[i].sim_rows > 0 || A.draws[i].mode != 0) atomicOr(&s_abort, 2u);  // needs k_edge_sim: sized pipeline
  }
  __syncthreads();
  if (s_abort) {
    if (t == 0) A.out[0] = s_abort;
    return;
  }
  const uint32_t n_slots = block_scan_inplace(A.draw_bands, A.n_draws, sh);
  if (n_slots > kSmallSlotCap) {
    if (t == 0) A.out[0] = 1u;
    return;
  }
  for (uint32_t i = t; i < A.n_draws; i += T) assign_band_base(i, A.draws, A.draw_bands, A.hots, A.boxes, A.sfcs);
  for (uint32_t i = t; i <= n_slots; i += T) {
    A.band_count[i] = 0u;
    A.band_cursor[i] = 0u;
    A.band_xr[i] = make_uint2(0u, 0u);
  }
  __syncthreads();
  // ---- K3a: bin edges
  for (uint32_t i = t; i < n_edges; i += T) bin_count_edge(i, A.edges, A.edge_draw, A.draws, A.band_count);
  __syncthreads();
  const uint32_t n_band = block_scan_inplace(A.band_count, n_slots, sh);
  if (n_band > kSmallBandCap) {
    if (t == 0) A.out[0] = 1u;
    return;
  }
  for (uint32_t i = t; i < n_edges; i += T)
    bin_scatter_edge(i, A.edges, A.edge_draw, A.draws, A.band_count, A.band_cursor, A.band_edges, A.band_hdr, A.band_xr, kSmallBandCap);
  __syncthreads();
  // ---- K3b: ordered draw list per surface tile row (work item = (surface, tile row); at most kDrawChunk draws: one chunk)
  auto work_item = [&](uint32_t w, uint32_t& si, int& band) Z2D_LAMBDA {
    uint32_t lo = 0, hi = A.n_sfc;
    while (hi - lo > 1) {
      const uint32_t mid = (lo + hi) >> 1;
      if (A.work_base[mid] <= w) lo = mid; else hi = mid;
    }
    si = lo;
    band = (int)(w - A.work_base[lo]);
  };
  for (uint32_t w = t; w < A.n_work; w += T) {
    uint32_t si;
    int band;
    work_item(w, si, band);
    const DevSurface& s = A.sfcs[si];
    A.list_cnt[w] = band_list_row<false>(band, s.draw_begin, s.draw_end, reinterpret_cast<const int4*>(A.boxes), 0u, 0u, nullptr, nullptr, nullptr);
  }
  __syncthreads();
  const uint32_t n_items = block_scan_inplace(A.list_cnt, A.n_work, sh);
  if (n_items > kSmallItemCap) {
    if (t == 0) A.out[0] = 1u;
    return;
  }
  for (uint32_t w = t; w < A.n_work; w += T) {
    uint32_t si;
    int band;
    work_item(w, si, band);
    const DevSurface& s = A.sfcs[si];
    band_list_row<true>(band, s.draw_begin, s.draw_end, reinterpret_cast<const int4*>(A.boxes), 0u, A.list_cnt[w], A.list_items, A.band_count, A.band_xr);
  }
  if (t == 0) {
    A.out[1] = n_edges;
    A.out[2] = n_slots;
    A.out[3] = n_items;
    A.out[4] = n_band;
  }
}

}  // namespace z2d
