// ORACLE -- test infrastructure only (see ref_internal.h).
// C entry points (loaded with ctypes by tests/, __graft_entry__.smoke() and
// bench.py's CPU-baseline legs).  Same PODs as include/z2d_cuda.h, but
// stateless and operating directly on caller-owned HOST buffers.
#include <algorithm>

#include "ref_internal.h"

using namespace zref;

extern "C" {

typedef struct z2d_ref_surface {
  void* buf;
  uint32_t format;
  int32_t width, height;
} z2d_ref_surface;

size_t z2d_ref_surface_byte_len(uint32_t fmt, int32_t w, int32_t h) { return sfc_byte_len(fmt, w, h); }

// Surface.initPixel / paintPixel (surface.zig:128-157, 295, 525, 775)
int32_t z2d_ref_surface_paint_pixel(void* buf, uint32_t fmt, int32_t w, int32_t h, const z2d_pixel* px) {
  if (w < 1) return Z2D_E_INVALID_WIDTH;
  if (h < 1) return Z2D_E_INVALID_HEIGHT;
  Sfc s{(uint8_t*)buf, fmt, w, h};
  std::memset(buf, 0, sfc_byte_len(fmt, w, h));
  for (size_t i = 0, n = (size_t)w * (size_t)h; i < n; i++) sfc_paint(s, i, *px);
  return Z2D_OK;
}

// Surface.putPixel (surface.zig:288, 519, 769)
int32_t z2d_ref_surface_put_pixel(void* buf, uint32_t fmt, int32_t w, int32_t h, int32_t x, int32_t y, const z2d_pixel* px) {
  if (x < 0 || y < 0 || x >= w || y >= h) return Z2D_OK;
  Sfc s{(uint8_t*)buf, fmt, w, h};
  sfc_paint(s, (size_t)w * (size_t)y + (size_t)x, *px);
  return Z2D_OK;
}

// Surface.downsample (surface.zig:447-490): in place; returns the new dimensions
int32_t z2d_ref_surface_downsample(void* buf, uint32_t fmt, int32_t w, int32_t h, int32_t* w_out, int32_t* h_out) {
  Sfc s{(uint8_t*)buf, fmt, w, h};
  sfc_downsample(s);
  *w_out = s.w;
  *h_out = s.h;
  return Z2D_OK;
}

// Polygon.inBox (Polygon.zig:142-201) on bare extents (the reference's own KAT table sets nothing else)
int32_t z2d_ref_in_box(double left, double top, double right, double bottom, double scale, int32_t box_width, int32_t box_height) {
  Polygon p;
  p.ext_left = left; p.ext_top = top; p.ext_right = right; p.ext_bottom = bottom;
  return p.in_box(scale, box_width, box_height) ? 1 : 0;
}

// benchmark statistic: pixels with coverage > 0 composited so far by the MSAA rasteriser
uint64_t z2d_ref_covered_px(int32_t reset) {
  uint64_t v = g_covered_px;
  if (reset) g_covered_px = 0;
  return v;
}

// Number of times a call reached a `debug.assert` of the reference that does not hold (the reference panics there in safe
// builds; see ref_stroke.cpp).  Tests use it to keep such calls out of parity comparisons: there is no reference result.
uint64_t z2d_ref_assert_trips(int32_t reset) {
  uint64_t v = g_assert_trips;
  if (reset) g_assert_trips = 0;
  return v;
}

static int check_pattern(const z2d_pattern* p) {  // painter.zig:73-79
  if (p->kind == Z2D_PATTERN_OPAQUE && !px_can_demultiply(p->pixel)) return Z2D_E_PIXEL_SOURCE_NOT_PREMULTIPLIED;
  return Z2D_OK;
}

static bool is_closed_node_set(const z2d_node* nodes, size_t n) {  // path_nodes.zig:23-37
  if (n == 0) return false;
  bool closed = false;
  for (size_t i = 0; i < n; i++) {
    if (nodes[i].tag == Z2D_NODE_MOVE_TO) {
      if (!closed && i != 0) break;
    } else if (nodes[i].tag == Z2D_NODE_CLOSE_PATH) {
      closed = true;
    } else {
      closed = false;
    }
  }
  return closed;
}

static void run_raster(Sfc& s, const Src& pat, Polygon& poly, uint32_t aa, uint32_t rule, uint32_t op, uint32_t prec) {
  switch (aa) {
    case Z2D_AA_NONE: raster_direct(s, pat, poly, rule, op, prec); break;
    case Z2D_AA_SUPERSAMPLE_4X: raster_supersample(s, pat, poly, rule, op, prec); break;
    default: raster_multisample(s, pat, poly, rule, op, prec);
  }
}

// painter.fill (painter.zig:66-143)
int32_t z2d_ref_fill(void* buf, uint32_t fmt, int32_t w, int32_t h, const z2d_pattern* pattern, const z2d_node* nodes,
                     size_t n, const z2d_fill_opts* o) {
  int rc = check_pattern(pattern);
  if (rc) return rc;
  if (n == 0) return Z2D_OK;
  if (!is_closed_node_set(nodes, n)) return Z2D_E_PATH_NOT_CLOSED;
  uint32_t aa = (fmt == Z2D_FMT_ALPHA1) ? (uint32_t)Z2D_AA_NONE : o->anti_aliasing_mode;
  double scale = aa == Z2D_AA_NONE ? 1 : 4;
  Polygon poly;
  rc = fill_plot(nodes, n, scale, std::max(o->tolerance, 0.001), poly);
  if (rc) return rc;
  Sfc s{(uint8_t*)buf, fmt, w, h};
  Src pat;
  src_from_pattern(*pattern, pat);
  run_raster(s, pat, poly, aa, o->fill_rule, o->op, o->precision);
  return Z2D_OK;
}

static void stroke_params(const z2d_stroke_opts* o, double scale, StrokeParams& sp) {  // painter.zig:287-304
  const double min_w = 0.00390625;
  sp.cap = o->line_width >= 2 ? o->line_cap_mode : (uint32_t)Z2D_CAP_BUTT;
  sp.join = o->line_width >= 2 ? o->line_join_mode : (uint32_t)Z2D_JOIN_MITER;
  sp.miter_limit = o->line_width >= 2 ? o->miter_limit : 10.0;
  sp.ctm = {o->ctm[0], o->ctm[1], o->ctm[2], o->ctm[3], o->ctm[4], o->ctm[5]};
  sp.dashes = o->dashes;
  sp.n_dashes = o->n_dashes;
  sp.dash_offset = o->dash_offset;
  sp.scale = scale;
  sp.thickness = o->line_width >= min_w ? o->line_width : min_w;
  sp.tolerance = std::max(o->tolerance, 0.001);
}

// painter.stroke (painter.zig:214-344)
int32_t z2d_ref_stroke(void* buf, uint32_t fmt, int32_t w, int32_t h, const z2d_pattern* pattern, const z2d_node* nodes,
                       size_t n, const z2d_stroke_opts* o) {
  int rc = check_pattern(pattern);
  if (rc) return rc;
  Xf ctm{o->ctm[0], o->ctm[1], o->ctm[2], o->ctm[3], o->ctm[4], o->ctm[5]}, inv;
  if (!ctm.inverse(inv)) return Z2D_E_INVALID_MATRIX;
  if (n == 0) return Z2D_OK;
  uint32_t aa = (fmt == Z2D_FMT_ALPHA1) ? (uint32_t)Z2D_AA_NONE : o->anti_aliasing_mode;
  Sfc s{(uint8_t*)buf, fmt, w, h};
  Src pat;
  src_from_pattern(*pattern, pat);
  if (o->hairline) {
    // note: the un-forced AA mode is passed on (painter.zig:261)
    return hairline_stroke(s, pat, nodes, n, o->tolerance, o->dashes, o->n_dashes, o->dash_offset, o->op, o->precision,
                           o->anti_aliasing_mode);
  }
  double scale = aa == Z2D_AA_NONE ? 1 : 4;
  StrokeParams sp;
  stroke_params(o, scale, sp);
  Polygon poly;
  rc = stroke_plot(nodes, n, sp, poly);
  if (rc) return rc;
  run_raster(s, pat, poly, aa, Z2D_FILL_NON_ZERO, o->op, o->precision);
  return Z2D_OK;
}

// compositor.SurfaceCompositor.run (compositor.zig:302-440); SURFACE params
// point at z2d_ref_surface descriptors.
int32_t z2d_ref_composite(void* buf, uint32_t fmt, int32_t w, int32_t h, int32_t dst_x, int32_t dst_y,
                          const z2d_comp_op* ops, size_t n_ops, uint32_t precision) {
  Sfc s{(uint8_t*)buf, fmt, w, h};
  std::vector<SurfOp> sops(n_ops);
  std::vector<Sfc> sfcs;
  sfcs.reserve(n_ops * 2);
  auto conv = [&](const z2d_comp_param& p, Src& out) {
    if (p.kind == Z2D_PARAM_SURFACE) {
      const z2d_ref_surface* rs = (const z2d_ref_surface*)p.surface;
      sfcs.push_back(Sfc{(uint8_t*)rs->buf, rs->format, rs->width, rs->height});
      out.kind = Z2D_PARAM_SURFACE;
      out.sfc = &sfcs.back();
    } else {
      src_from_param(p, out);
    }
  };
  for (size_t i = 0; i < n_ops; i++) {
    sops[i].op = ops[i].op;
    conv(ops[i].dst, sops[i].dst);
    conv(ops[i].src, sops[i].src);
  }
  surface_run(s, dst_x, dst_y, sops.data(), n_ops, precision);
  return Z2D_OK;
}

// compositor.runPixel (compositor.zig:626-645) on RGBA pixels -- used by the
// operator known-answer tests.  in/out: {r,g,b,a}.
void z2d_ref_run_pixel(uint32_t precision, const uint8_t* dst, const uint8_t* src, uint32_t op, uint8_t* out) {
  RGBA16 d{dst[0], dst[1], dst[2], dst[3]}, s{src[0], src[1], src[2], src[3]}, r;
  if (precision == Z2D_PRECISION_INTEGER) {
    r = int_op(op, d, s);
  } else {
    RGBAF df{d.r / 255.0f, d.g / 255.0f, d.b / 255.0f, d.a / 255.0f}, sf{s.r / 255.0f, s.g / 255.0f, s.b / 255.0f, s.a / 255.0f};
    RGBAF rf = float_op(op, df, sf);
    r = {(int)std::round(255.0f * rf.r), (int)std::round(255.0f * rf.g), (int)std::round(255.0f * rf.b), (int)std::round(255.0f * rf.a)};
  }
  out[0] = (uint8_t)r.r; out[1] = (uint8_t)r.g; out[2] = (uint8_t)r.b; out[3] = (uint8_t)r.a;
}

// Edge-list extraction for tessellation parity tests.  Writes up to `cap`
// edges as {y0,y1,x_start,x_inc}; returns the edge count (or a negative status).
// extents = {top,bottom,left,right}.
int64_t z2d_ref_flatten_fill(const z2d_node* nodes, size_t n, double scale, double tolerance, double* edges, size_t cap,
                             double* extents) {
  Polygon poly;
  int rc = fill_plot(nodes, n, scale, std::max(tolerance, 0.001), poly);
  if (rc) return rc;
  for (size_t i = 0; i < poly.edges.size() && i < cap; i++) std::memcpy(edges + 4 * i, &poly.edges[i], 32);
  if (extents) { extents[0] = poly.ext_top; extents[1] = poly.ext_bottom; extents[2] = poly.ext_left; extents[3] = poly.ext_right; }
  return (int64_t)poly.edges.size();
}

int64_t z2d_ref_flatten_stroke(const z2d_node* nodes, size_t n, const z2d_stroke_opts* o, double scale, double* edges,
                               size_t cap, double* extents) {
  StrokeParams sp;
  stroke_params(o, scale, sp);
  Polygon poly;
  int rc = stroke_plot(nodes, n, sp, poly);
  if (rc) return rc;
  for (size_t i = 0; i < poly.edges.size() && i < cap; i++) std::memcpy(edges + 4 * i, &poly.edges[i], 32);
  if (extents) { extents[0] = poly.ext_top; extents[1] = poly.ext_bottom; extents[2] = poly.ext_left; extents[3] = poly.ext_right; }
  return (int64_t)poly.edges.size();
}

// Per-pixel source evaluation (gradient / dither parity tests): premultiplied RGBA8.
void z2d_ref_pattern_pixel(const z2d_pattern* p, int32_t x, int32_t y, uint8_t* out) {
  Src s;
  src_from_pattern(*p, s);
  uint8_t px[4] = {0, 0, 0, 0};
  Sfc tmp{px, Z2D_FMT_RGBA, 1, 1};
  StrideOp o{Z2D_OP_SRC, nullptr, &s};
  stride_run(tmp, 0, 1, x, y, &o, 1, Z2D_PRECISION_INTEGER);
  std::memcpy(out, px, 4);
}

}  // extern "C"
