// ORACLE -- TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of the z2d (vancluever/z2d @ v0.12.0-unreleased) fill/stroke
// rasterise-and-composite path.  Nothing under oracle/ is linked, imported or
// executed by the product (z2d_b200/, libz2d_cuda.so); only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// use it, and only as the checker or the timed CPU baseline.
//
// Parity status: PINNED -- the reference cannot be built here (no Zig
// toolchain), so this restatement is pinned against the reference's own golden
// images (spec/files/*.png, copied to tests/golden/spec_files) and the
// operator known-answer tables of src/compositor.zig:3078-3860
// (tests/golden/compositor_kat.json).  See tests/test_oracle_*.py.
//
// All file:line citations are relative to the reference tree.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../include/z2d_cuda.h"

namespace zref {

// ---------------------------------------------------------------- geometry
struct Pt {
  double x, y;
};
static inline bool pt_eq(Pt a, Pt b) { return a.x == b.x && a.y == b.y; }  // Point.zig:16-18

// Zig @round: half away from zero (== C round()).
static inline double zround(double v) { return std::round(v); }

// Transformation.zig
struct Xf {
  double ax, by, cx, dy, tx, ty;
  bool is_identity() const { return ax == 1 && by == 0 && cx == 0 && dy == 1 && tx == 0 && ty == 0; }
  double det() const { return ax * dy - by * cx; }
  void dist(double& x, double& y) const {  // userToDeviceDistance
    double ix = x, iy = y;
    x = ax * ix + by * iy;
    y = cx * ix + dy * iy;
  }
  void point(double& x, double& y) const {  // userToDevice
    dist(x, y);
    x += tx;
    y += ty;
  }
  bool inverse(Xf& out) const;  // false -> InvalidMatrix
};

// Polygon.Edge (tess/Polygon.zig:18-47)
struct Edge {
  double y0, y1, x_start, x_inc;
  int dir() const { return y0 < y1 ? -1 : 1; }
  double top() const { return y0 < y1 ? y0 : y1; }
  double bottom() const { return y0 < y1 ? y1 : y0; }
};

struct Polygon {
  std::vector<Edge> edges;
  double scale = 1;
  double ext_top = 0, ext_bottom = 0, ext_left = 0, ext_right = 0;
  void add_edge(Pt p0, Pt p1);                                 // Polygon.zig:61-109
  void add_contour(const std::vector<Pt>& pts);               // Polygon.zig:115-137
  bool in_box(double scale, int box_w, int box_h) const;      // Polygon.zig:142-201
};

// ---------------------------------------------------------------- surfaces
struct Sfc {
  uint8_t* buf;
  uint32_t fmt;
  int32_t w, h;
};
size_t sfc_byte_len(uint32_t fmt, int64_t w, int64_t h);

struct RGBA16 {  // compositor.zig:665-703 (channels 0..255 held in ints)
  int r, g, b, a;
};
struct RGBAF {
  float r, g, b, a;
};

RGBA16 px_to_rgba16(const z2d_pixel& px);            // RGBA16.fromPixel
bool px_is_opaque(const z2d_pixel& px);              // pixel.zig Pixel.isOpaque
bool px_can_demultiply(const z2d_pixel& px);         // pixel.zig:504-514
RGBA16 sfc_load(const Sfc& s, size_t idx);           // fromStride / fromPixelT
void sfc_store(Sfc& s, size_t idx, RGBA16 v);        // toStride / toPixelT
void sfc_paint(Sfc& s, size_t idx, const z2d_pixel& px);  // buf[idx] = T.fromPixel(px)
void sfc_paint_stride(Sfc& s, int x, int y, size_t len, const z2d_pixel& px);   // surface.zig:537,787
void sfc_clear_stride(Sfc& s, int x, int y, size_t len);                        // surface.zig:323
void sfc_composite_stride(Sfc& s, int x, int y, size_t len, const z2d_pixel& px, uint32_t op,
                          uint8_t opacity);                                      // surface.zig:557,852
void sfc_downsample(Sfc& s);                                                    // surface.zig:447,687

// ---------------------------------------------------------------- sources
struct Grad {  // prepared gradient (stops converted to the interpolation space once)
  uint32_t type, method, polar;
  double geom[6];
  Xf inv;
  bool inv_identity;
  // radial precalcs (gradient.zig:262-300)
  double cdx, cdy, dr, min_dr, a, inv_a, inner_r, outer_r;
  std::vector<float> offsets;
  std::vector<RGBAF> colors;  // in method space: rgb(a) or h,s,l,a
};
void grad_prepare(const z2d_gradient& g, Grad& out);
float grad_offset(const Grad& g, int x, int y);
struct StopHit {
  RGBAF c0, c1;
  float t;
};
StopHit grad_search(const Grad& g, float offset);
RGBA16 grad_encode(const Grad& g, const StopHit& h);  // interpolateEncodeVec -> premultiplied RGBA8
RGBAF grad_linear(const Grad& g, const StopHit& h);   // interpolateVec -> demultiplied linear

struct Src {  // a compositor parameter evaluated per pixel
  uint32_t kind = Z2D_PARAM_NONE;
  z2d_pixel px{};
  Grad grad;  // GRADIENT, or DITHER over a gradient
  uint32_t dither_type = 0, dither_source = 0, dither_scale = 8;
  RGBAF dither_color{};  // DITHER over pixel/color: linear demultiplied
  const Sfc* sfc = nullptr;
};
void src_from_pattern(const z2d_pattern& p, Src& out);
void src_from_param(const z2d_comp_param& p, Src& out);

RGBA16 int_op(uint32_t op, RGBA16 d, RGBA16 s);   // IntegerOps (compositor.zig:1158-1568)
RGBAF float_op(uint32_t op, RGBAF d, RGBAF s);    // FloatOps   (compositor.zig:1571-2440)
bool op_requires_float(uint32_t op);              // compositor.zig:165-177
bool op_is_bounded(uint32_t op);                  // compositor.zig:187-196

struct StrideOp {
  uint32_t op;
  const Src* dst;  // nullptr == .none
  const Src* src;  // nullptr == .none
  // for SURFACE params: pixel index of the first pixel of the stride
  size_t dst_idx = 0, src_idx = 0;
};
// StrideCompositor.run (compositor.zig:540-621); (x,y) = device position of the first pixel.
void stride_run(Sfc& dst, size_t dst_idx, size_t len, int x, int y, const StrideOp* ops, size_t n_ops,
                uint32_t precision);
// SurfaceCompositor.run (compositor.zig:302-440)
struct SurfOp {
  uint32_t op;
  Src dst, src;
};
void surface_run(Sfc& dst, int dst_x, int dst_y, const SurfOp* ops, size_t n_ops, uint32_t precision);

// ---------------------------------------------------------------- tessellation / raster
int fill_plot(const z2d_node* nodes, size_t n, double scale, double tol, Polygon& out);  // fill_plotter.zig:21-97
struct StrokeParams {
  uint32_t cap, join;
  Xf ctm;
  const double* dashes;
  size_t n_dashes;
  double dash_offset, miter_limit, scale, thickness, tolerance;
};
int stroke_plot(const z2d_node* nodes, size_t n, const StrokeParams& sp, Polygon& out);  // stroke_plotter.zig:40-75

// Spline.decompose (tess/Spline.zig:37-71): calls emit(p) for every line_to.
template <class F>
void spline_decompose(Pt a, Pt b, Pt c, Pt d, double tolerance, F&& emit);

void raster_direct(Sfc& s, const Src& pat, Polygon& poly, uint32_t rule, uint32_t op, uint32_t prec);       // raster/direct.zig
extern uint64_t g_covered_px;
extern uint64_t g_assert_trips;
void raster_multisample(Sfc& s, const Src& pat, Polygon& poly, uint32_t rule, uint32_t op, uint32_t prec);  // raster/multisample.zig
void raster_supersample(Sfc& s, const Src& pat, Polygon& poly, uint32_t rule, uint32_t op, uint32_t prec);  // raster/supersample.zig

// polyline_plotter.zig + hairline.zig
int hairline_stroke(Sfc& s, const Src& pat, const z2d_node* nodes, size_t n, double tol, const double* dashes,
                    size_t n_dashes, double dash_offset, uint32_t op, uint32_t prec, uint32_t aa);

// shared.zig
void composite_opaque(uint32_t op, Sfc& s, const Src& pat, int x, int y, size_t len, uint32_t prec);
void composite_opacity(uint32_t op, Sfc& s, const Src& pat, int x, int y, size_t len, uint32_t prec, uint8_t opacity);

// ---- Spline implementation (header because of the functor) ----
struct Knots {
  Pt a, b, c, d;
  double error_sq() const;  // Spline.zig:83-123
  Knots de_casteljau();     // Spline.zig:128-151
};
template <class F>
static void spline_into(Knots& s1, Pt start, double tol_sq, F& emit) {  // Spline.zig:56-71
  if (s1.error_sq() < tol_sq) {
    if (!pt_eq(s1.a, start)) emit(s1.a);
    return;
  }
  Knots s2 = s1.de_casteljau();
  spline_into(s1, start, tol_sq, emit);
  spline_into(s2, start, tol_sq, emit);
}
template <class F>
void spline_decompose(Pt a, Pt b, Pt c, Pt d, double tolerance, F&& emit) {
  if (pt_eq(a, b) && pt_eq(c, d)) {  // Spline.zig:39-42
    emit(d);
    return;
  }
  Knots s1{a, b, c, d};
  spline_into(s1, a, tolerance * tolerance, emit);
  emit(d);
}

}  // namespace zref
