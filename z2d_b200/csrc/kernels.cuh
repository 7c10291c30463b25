// Kernel argument blocks and launch wrappers (implemented in kernels.cu).
#pragma once
#include "z2d_batch.cuh"

namespace z2d {

#ifndef Z2D_RASTER_THREADS
#define Z2D_RASTER_THREADS 32
#endif
constexpr int kRasterThreads = Z2D_RASTER_THREADS;  // one warp per tile
constexpr uint32_t kDrawChunk = 256; // draws per band-list work item
constexpr uint32_t kMaxCompOps = 8;

struct RasterArgs {
  const DevSurface* sfcs;
  uint32_t n_sfc;
  uint32_t n_tiles;
  const uint32_t* work_base;   // per surface: first band-list work item
  const uint32_t* list_off;    // per work item (+1): offset into list_items
  const uint4* list_items;     // {draw, tx0 | tx1 << 16, first binned edge, edge count | kItem* << 24}, draw order within a tile-row
  const DevDraw* draws;
  const DrawHot* hots;
  const uint32_t* band_off;    // per (draw, tile-row) slot (+1): offset into band_edges
  const DevEdge* band_edges;
  const int4* band_hdr;        // per binned edge: {xlo, xhi, first row | dir<<31, last row}
  unsigned long long* counters;  // [0] covered pixels, [1] region pixels (may be null)
  const int4* sim_rows;        // row records of k_edge_sim (draws flagged kDrawUnpaired / kDrawRowRecords)
  const uint32_t* abort;       // small-batch path: non-zero => the prepare kernel gave up, composite nothing (may be null)
  GradTables T;
};

struct CompOp {
  uint32_t op, has_dst, has_src, _pad;
  DevSrc dst, src;
};

struct CompArgs {  // SurfaceCompositor.run after clipping (compositor.zig:347-388)
  uint8_t* data;
  uint32_t fmt;
  int32_t w, h;
  int32_t dst_start_x, dst_start_y, src_start_x, src_start_y;
  int32_t scan_w, rows;
  int32_t y_origin;        // band destination: canvas row of its first row (patterns are evaluated in canvas space)
  uint32_t max_stops;      // largest stop count among the gradients of this call (k_composite_gen keeps <= 16 in shared memory)
  uint32_t n_ops, precision;
  GradTables T;
  CompOp ops[kMaxCompOps];
};

size_t scan_tmp_len(uint32_t n);
// out[0..n) = exclusive scan of in[0..n), out[n] = total
void exclusive_scan(const uint32_t* in, uint32_t* out, uint32_t n, uint32_t* tmp, cudaStream_t st);

// pens: array of 6-double pen vertices {px,py,cw.dx,cw.dy,ccw.dx,ccw.dy} (tess/Pen.zig), dashes: concatenated dash arrays
void launch_flatten_count(const DevSubPath* sps, uint32_t n_sp, const z2d_node* nodes, DevDraw* draws, uint32_t* sp_count, const void* pens,
                          const double* dashes, const uint32_t* order, cudaStream_t st);
// order[t] = the sub-path thread t of the two kernels above plots (sub-paths bucketed by stroke style; null = identity);
// keys_scratch: 1024 uint32
void launch_sp_order(const DevSubPath* sps, uint32_t n_sp, const DevDraw* draws, uint32_t* keys_scratch, uint32_t* order, cudaStream_t st);
void launch_flatten_emit(const DevSubPath* sps, uint32_t n_sp, const z2d_node* nodes, const DevDraw* draws, const uint32_t* sp_off,
                         DevEdge* edges, uint32_t* edge_draw, const void* pens, const double* dashes, const uint32_t* order, cudaStream_t st);
// single-pass node-parallel flattening (kernels.cu, k_node_bounds / k_flatten_nodes_pool): per-node upper bounds into `counts`
// (then scanned by the caller), edges into the pool at ctr[3] + offset; ctr[4] != 0: a node exceeded its bound
void launch_node_bounds(const DevSubPath* sps, uint32_t n_sp, uint32_t* node_sp, uint32_t n_nodes, const z2d_node* nodes, const DevDraw* draws,
                        uint32_t* counts, uint32_t* curve_list, cudaStream_t st);
void launch_flatten_nodes_pool(const DevSubPath* sps, const uint32_t* node_sp, uint32_t n_nodes, const z2d_node* nodes, DevDraw* draws,
                               const uint32_t* node_offs, uint32_t* ctr, DevEdge* edges, uint32_t* edge_draw, uint32_t edge_cap,
                               const uint32_t* curve_list, cudaStream_t st);
void raster_preload();  // raster.cu: loads the K4 kernels (call once per process, before other kernels are used)
// unit stroker (stroke_units.cuh) for the sub-paths flagged kSpStrokeUnits: walker -> units -> links.  ctr[0..2] = units, links and
// edge slots taken (they run past the capacities when those are too small: the caller enlarges and repeats).
constexpr size_t kStrokeUnitBytes = 64, kStrokeLinkBytes = 16, kStrokePortBytes = 64;  // per unit / link / unit
void launch_stroke_units(const DevSubPath* sps, uint32_t n_sp, const z2d_node* nodes, DevDraw* draws, const void* pens, const double* dashes,
                         const uint32_t* order, void* units, uint32_t unit_cap, void* links, uint32_t link_cap, void* ports, uint32_t* ctr,
                         DevEdge* edges, uint32_t* edge_draw, uint32_t edge_cap, cudaStream_t st);
// node-parallel flattening of the sub-paths flagged kSpNodeParallel: count pass (emit = false: also builds node_sp) / emit pass
void launch_flatten_nodes(bool emit, const DevSubPath* sps, uint32_t n_sp, uint32_t* node_sp, uint32_t n_nodes, const z2d_node* nodes,
                          DevDraw* draws, uint32_t* counts, const uint32_t* offs, DevEdge* edges, uint32_t* edge_draw,
                          uint32_t* curve_list /* n_nodes + 1 */, cudaStream_t st);
void launch_setup_draws(DevDraw* draws, uint32_t n, const DevSurface* sfcs, uint32_t* draw_bands, DrawBox* boxes, unsigned long long* counters, cudaStream_t st);
void launch_expand_draws(const DrawIn* in, const StrokeIn* strokes, const DevSrc* srcs, DevDraw* draws, uint32_t n, cudaStream_t st);
void launch_assign_band_base(DevDraw* draws, uint32_t n, const uint32_t* band_off, DrawHot* hots, DrawBox* boxes, const DevSurface* sfcs,
                             cudaStream_t st);
void launch_bin_count(const DevEdge* edges, const uint32_t* edge_draw, uint32_t n, const DevDraw* draws, uint32_t* band_count, cudaStream_t st);
void launch_bin_scatter(const DevEdge* edges, const uint32_t* edge_draw, uint32_t n, const DevDraw* draws, const uint32_t* band_off,
                        uint32_t* band_cursor, DevEdge* band_edges, int4* band_hdr, uint2* band_xr /* zeroed, one per slot */, cudaStream_t st);
// chunk_base[s] = number of (surface, kDrawChunk-draw chunk) blocks before surface s; one block per chunk
void launch_band_lists(bool write, const DevSurface* sfcs, uint32_t n_sfc, const uint32_t* work_base, const uint32_t* chunk_base,
                       uint32_t n_chunks, const DrawBox* boxes, uint32_t* cnt, const uint32_t* off, uint4* items, const uint32_t* band_off,
                       const uint2* band_xr, uint32_t max_tiles_y, cudaStream_t st);
// exact scanline replay of the draws k_setup_draws gave row records (counters[4] rows, counters[5] scratch slots)
void launch_edge_sim(const DevDraw* draws, uint32_t n_draws, const DevSurface* sfcs, const DevEdge* edges, const uint32_t* sp_off,
                     uint32_t* perm, int32_t* xs, int4* rows, cudaStream_t st);
// ---- glyph instances of a text run (z2d_fill_glyphs): cached outline nodes [src, src + n_nodes) -> batch nodes [dst, ...)
struct GlyphInst {
  uint32_t src, n_nodes, dst, _pad;
  double m[6];
};
void launch_expand_glyphs(const GlyphInst* inst, uint32_t n, const z2d_node* cache, z2d_node* nodes, cudaStream_t st);

// ---- small batches in two launches (smallbatch.cuh)
constexpr uint32_t kSmallMaxDraws = 64, kSmallMaxNodes = 4096, kSmallMaxSubPaths = 1024, kSmallMaxWork = 4096;
constexpr uint32_t kSmallEdgeCap = 1u << 18, kSmallBandCap = 1u << 19, kSmallSlotCap = 1u << 14, kSmallItemCap = 1u << 14;
constexpr int kSmallThreads = 512;

struct SmallArgs {
  const DrawIn* draws_in;
  const StrokeIn* strokes;
  const DevSrc* srcs;
  const DevSubPath* sps;
  const z2d_node* nodes;
  const DevSurface* sfcs;
  const uint32_t* work_base;
  uint32_t n_draws, n_sp, n_nodes, n_sfc, n_work;
  DevDraw* draws;
  DrawHot* hots;
  DrawBox* boxes;
  uint32_t* draw_bands;     // n_draws + 1 (scanned in place)
  uint32_t* node_sp;        // n_nodes
  uint32_t* cnt;            // n_sp + n_nodes + 1 (scanned in place -> offsets)
  DevEdge* edges;
  uint32_t* edge_draw;
  uint32_t* band_count;     // kSmallSlotCap + 1 (scanned in place -> band_off)
  uint32_t* band_cursor;
  uint2* band_xr;
  DevEdge* band_edges;
  int4* band_hdr;
  uint32_t* list_cnt;       // n_work + 1 (scanned in place -> list_off)
  uint4* list_items;
  unsigned long long* counters;
  uint32_t* out;            // device copy of the result block: [0] abort reason, [1] edges, [2] slots, [3] items, [4] binned edges
};

void launch_small_batch(const SmallArgs& A, cudaStream_t st);
// rich: the batch holds strokes or gradient / dither sources (k_raster_tiles_rich, see raster.cuh)
void launch_raster(const RasterArgs& A, bool rich, cudaStream_t st);
// isolated single-draw modes (slowpath.cuh)
void launch_hairline(const DevSurface* sfcs, const DevDraw* draws, uint32_t draw_index, const z2d_node* nodes, uint32_t node_begin,
                     uint32_t node_end, const double* dashes, const GradTables& T, cudaStream_t st);
void launch_composite(const CompArgs& A, int sm_count, cudaStream_t st);
struct ExportArgs {  // z2d_surface_export
  const uint8_t* data;
  uint32_t fmt;
  int32_t w, h;
  uint32_t items_per_row;  // pixels (>= 8 bits per pixel) or output bytes (packed greys)
  uint32_t filter_byte;    // 1: every row starts with the PNG filter-type byte 0
  size_t row_bytes;        // including the filter byte
  const uint8_t* gamma;    // 256-entry channel table, or null (linear)
  uint8_t* out;
};
void launch_export(const ExportArgs& A, int sm_count, cudaStream_t st);
void launch_paint(uint8_t* data, uint32_t fmt, size_t n_px, uint32_t raw, cudaStream_t st);
void launch_downsample(const uint8_t* src, uint8_t* dst, uint32_t fmt, int w_in, int w_out, int h_out, cudaStream_t st);
void launch_put_pixel(uint8_t* data, uint32_t fmt, size_t idx, uint32_t raw, cudaStream_t st);

}  // namespace z2d
