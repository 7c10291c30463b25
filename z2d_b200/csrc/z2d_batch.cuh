// Batch data model shared by the host recorder (z2d_lib.cu) and the kernels.
//
// A *batch* is the ordered list of painter.fill / painter.stroke calls recorded
// since the last flush.  It is rendered by a fixed kernel pipeline:
//
//   K1 flatten_count  one thread per sub-path: nodes -> #edges, extents   (fill_plotter / Spline / Polygon.addEdge)
//   K1 flatten_emit   same walk, writes edges {y0,y1,x_start,x_inc}
//   K2 setup_draws    one thread per draw: extents -> pixel region, tile range   (multisample.zig:38-78 etc.)
//   K3 bin_edges      edges -> per (draw, tile-row) lists (count / scan / scatter; order-free)
//   K3 band_lists     per tile-row of each surface: ordered list of the draws touching it (count / scan / write)
//   K4 raster_tiles   one warp per 16x16-pixel tile: walks its draws IN SUBMISSION ORDER, evaluates the
//                     4x4 sample coverage of each from the binned edges and composites into the
//                     register-resident tile; the tile is read once and written once per batch.
#pragma once
#include <string.h>

#include "z2d_device.cuh"

namespace z2d {

constexpr int kTile = 16;          // tile edge in pixels
constexpr int kTileShift = 4;

struct DevSurface {
  uint8_t* data;
  uint32_t fmt;
  int32_t w, h;
  int32_t tiles_x, tiles_y;
  uint32_t tile_base;   // first global tile index of this surface in the batch
  uint32_t band_base;   // first global tile-row ("band") index
  uint32_t draw_begin, draw_end;  // draws of this surface (draws are grouped by surface, order preserved)
  // band surfaces (z2d_surface_create_band): the memory holds rows [y0, y0 + h) of a virtual canvas vh rows high; all
  // geometry, regions and pattern coordinates stay in canvas space.  Ordinary surfaces: y0 = 0, vh = h.
  int32_t y0, vh;
};

struct DevSubPath {  // one move_to ... run of nodes (state resets at every move_to)
  uint32_t draw;
  uint32_t node_begin, node_end;  // [begin,end) in the batch node array
  uint32_t flags;                 // kSpLastOfDraw | kSpNodeParallel | kSpStrokeUnits
};

constexpr uint32_t kSpLastOfDraw = 1u;     // node_end is the end of the draw's node list
constexpr uint32_t kSpNodeParallel = 2u;   // fill sub-path of the form move_to, segments..., close_path with >= 2 moving segments:
                                           // every node is flattened by its own thread (k_flatten_nodes)
constexpr uint32_t kSpStrokeUnits = 4u;    // sub-path of an ordinary stroke: tessellated by the unit stroker (stroke_units.cuh), its
                                           // edges go to the pool at the front of the edge array, not to a counted range

struct DevDraw {
  // --- recorded on the host
  uint32_t surface;
  uint32_t kind;       // 0 fill, 1 stroke
  uint32_t aa;         // effective AA mode: Z2D_AA_NONE / MULTISAMPLE_4X / SUPERSAMPLE_4X
  uint32_t rule, op, precision;  // precision already upgraded when the operator requires float
  uint32_t reduces;    // shared.zig:111-117 fillReducesToSource
  uint32_t paint_raw;  // T.fromPixel(source pixel) for the surface format (opaque fast path)
  double scale, tolerance;
  DevSrc src;
  // stroke parameters (painter.zig:287-304 already applied)
  uint32_t cap, join;
  double thickness, miter_limit, dash_offset;
  double ctm[6], inv[6];  // CTM and its inverse (Transformation.inverse, computed on the host)
  uint32_t dash_begin, dash_count;
  uint32_t pen_begin, pen_count;
  // isolated modes (slowpath.cuh): 0 tile pipeline, 1 hairline, 2 direct rasteriser with an unbounded operator
  uint32_t mode;
  uint32_t hair_aa;            // hairline: the caller's AA mode (none => Bresenham, else Wu)
  double hair_tolerance;       // hairline: untouched opts.tolerance (painter.zig:262-277)
  uint32_t node_begin, node_end;  // hairline: the whole node list of the draw
  // --- produced on the device
  uint32_t flags;      // kDrawUnpaired: some sub-path left a dangling edge (fill_plotter.zig:78-97 with 2 points)
  uint32_t _pad0;
  long long ext[4];    // order-encoded f64: top(min) bottom(max) left(min) right(max)
  uint32_t n_edges;
  uint32_t valid;
  int32_t rx0, rx1, ry0, ry1;   // pixel region whose coverage is evaluated: [rx0,rx1) x [ry0,ry1)
  int32_t tx0, tx1, ty0, ty1;   // tile range (inclusive) the draw touches
  int32_t ey0, ey1;             // tile-row range (inclusive) of the evaluated region (edge binning)
  int32_t pre_y0, pre_y1, pre_x, pre_rows;  // MSAA unbounded pre-clear (multisample.zig:96-110)
  uint32_t band_base;           // first (draw, tile-row) slot
  uint32_t unbounded;
  // exact scanline replay (k_edge_sim) for draws whose result depends on the ORDER of equal crossings: fills with a dangling
  // edge (kDrawUnpaired) and the direct rasteriser with an unbounded operator (mode 2)
  uint32_t sp_first;            // first sub-path of the draw (its edges start at sp_off[sp_first], in plotting order)
  uint32_t sim_base;            // first row record
  uint32_t sim_scratch;         // first slot of the per-edge scratch (permutation, x values)
  int32_t sim_y0, sim_rows;     // (sub-)scanlines [sim_y0, sim_y0 + sim_rows) of the canvas
  uint32_t _pad1;
};

constexpr uint32_t kDrawUnpaired = 1u;
constexpr uint32_t kDrawRowRecords = 2u;  // DrawHot.flags only: mode 2, composited from the row records of k_edge_sim

// What the host uploads per draw call (32 B); k_expand_draws turns it into the DevDraw the other kernels read.
// Sources other than a single pixel and all stroke parameters live in side tables.
constexpr uint32_t kNoIndex = 0xffffffffu;
struct DrawIn {
  uint32_t surface;
  uint32_t opts;          // kind | aa << 1 | rule << 3 | op << 4 | precision << 9 | reduces << 10 | mode << 11
  uint32_t paint_raw;
  uint32_t px_rgba;       // inline single-pixel source (src_index == kNoIndex)
  uint32_t src_index;     // index into the DevSrc side table
  uint32_t stroke_index;  // index into the StrokeIn side table (strokes only)
  double tolerance;
};
struct StrokeIn {
  uint32_t cap, join;
  uint32_t dash_begin, dash_count;
  uint32_t pen_begin, pen_count;
  uint32_t hair_aa, _pad;
  double thickness, miter_limit, dash_offset, hair_tolerance;
  double ctm[6], inv[6];
};
Z2D_HD uint32_t pack_draw_opts(uint32_t kind, uint32_t aa, uint32_t rule, uint32_t op, uint32_t precision, uint32_t reduces, uint32_t mode) {
  return kind | (aa << 1) | (rule << 3) | (op << 4) | (precision << 9) | (reduces << 10) | (mode << 11);
}

struct DevEdge {
  double y0, y1, x_start, x_inc;
};

// Compact per-draw records written by k_setup_draws: the tile-row lists only need the
// tile box, the raster kernel only the fields below (the 300-byte DevDraw is touched
// again only for gradient / dither sources).
struct DrawBox {  // tx0 < 0 => draw not valid
  int32_t tx0, tx1, ty0, ty1;   // tile box of this surface the draw touches (inclusive)
  int32_t es0, es1;             // tile rows of this surface that hold binned edges (inclusive; empty: es0 > es1)
  uint32_t band_base;           // first (draw, tile-row) slot: slot of surface tile row t = band_base + (t - es0)
  uint32_t item_flags;          // kItem* bits every list item of the draw carries (k_assign_band_base)
};

// One entry of a tile-row's ordered draw list (k_band_lists -> k_raster_tiles), everything the raster kernel needs to start
// on the pair without touching another table:
//   .x draw index   .y first | last << 16 tile column to visit (the x-range of the edges binned to THIS tile row, not the
//   draw's whole box: tiles left or right of every edge of a closed shape see winding 0)   .z first binned edge
//   .w number of binned edges (24 bits; 0xffffff: look it up) | flags << 24
constexpr uint32_t kItemAaMask = 3u;       // Z2D_AA_* of the draw
constexpr uint32_t kItemEvenOdd = 4u;
constexpr uint32_t kItemInRows = 8u;       // this tile row holds binned edges of the draw
constexpr uint32_t kItemSpecial = 16u;     // unbounded pre-clear / row records / supersample / dangling edge: consult DrawHot
constexpr uint32_t kItemFastBlend = 32u;   // 32-bit surface, single-pixel source, integer src_over: table-driven blend
struct alignas(16) DrawHot {
  uint32_t aa, rule, op, precision;
  uint32_t reduces, paint_raw, px_rgba, src_kind;
  int32_t rx0, rx1, ry0, ry1;
  int32_t ey0, ey1;
  uint32_t band_base, unbounded;
  int32_t pre_y0, pre_y1, pre_x, pre_rows;
  uint32_t flags;
  uint32_t sim_base;  // row records of k_edge_sim (kDrawUnpaired / kDrawRowRecords)
  int32_t sim_y0, sim_rows;
};

// order-preserving f64 <-> i64 (for atomicMin / atomicMax on extents)
Z2D_HD long long f64_order(double v) {
  long long b;
#ifdef __CUDA_ARCH__
  b = __double_as_longlong(v);
#else
  memcpy(&b, &v, 8);
#endif
  return b >= 0 ? b : (b ^ 0x7fffffffffffffffLL);
}
Z2D_HD double f64_unorder(long long k) {
  long long b = k >= 0 ? k : (k ^ 0x7fffffffffffffffLL);
#ifdef __CUDA_ARCH__
  return __longlong_as_double(b);
#else
  double v;
  memcpy(&v, &b, 8);
  return v;
#endif
}

// Zig @round on f64: half away from zero; exact (v - trunc(v) is exact)
Z2D_HD double round_half_away(double v) {
  double t = trunc(v);
  double f = v - t;
  if (fabs(f) >= 0.5) t += (v < 0.0 ? -1.0 : 1.0);
  return t;
}

}  // namespace z2d
