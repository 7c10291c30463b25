// C ABI of libz2d_cuda.so (include/z2d_cuda.h): contexts, device-resident
// surfaces, the draw-call recorder and the batch pipeline driver.
//
// Host work on the boundary is limited to what painter.fill / painter.stroke do
// before tessellation (argument validation, AA-mode selection, option clamping:
// painter.zig:66-104, 214-304), splitting the node list into sub-paths and
// converting gradient stops to their interpolation space once per call
// (the reference redoes that per pixel, color_vector.zig:197-207).
#include <algorithm>
#include <sched.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <nvtx3/nvToolsExt.h>  // header-only NVTX 3: ranges cost nothing unless a profiler injects its library

#include "blue_noise_table.h"
#include "kernels.cuh"

using namespace z2d;

namespace {

struct DevBuf {  // growable device buffer
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    size_t ncap = cap ? cap : 4096;
    while (ncap < bytes) ncap = ncap + ncap / 2 + 4096;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMalloc(&p, ncap);
    if (e == cudaSuccess) cap = ncap;
    return e;
  }
  // like ensure, but the first `keep` bytes survive a reallocation (device-to-device copy on `st`)
  cudaError_t ensure_keep(size_t bytes, size_t keep, cudaStream_t st) {
    if (bytes <= cap) return cudaSuccess;
    size_t ncap = cap ? cap : 4096;
    while (ncap < bytes) ncap = ncap + ncap / 2 + 4096;
    void* np = nullptr;
    cudaError_t e = cudaMalloc(&np, ncap);
    if (e != cudaSuccess) return e;
    if (p && keep) e = cudaMemcpyAsync(np, p, keep, cudaMemcpyDeviceToDevice, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (p) cudaFree(p);
    p = np;
    cap = ncap;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <class T>
  T* as() const { return (T*)p; }
};

template <class T>
struct PinnedVec {  // growable pinned host array (async H2D source)
  T* p = nullptr;
  size_t n = 0, cap = 0;
  bool reserve(size_t want) {
    if (want <= cap) return true;
    size_t ncap = cap ? cap : 1024;
    while (ncap < want) ncap *= 2;
    T* np = nullptr;
    if (cudaHostAlloc((void**)&np, ncap * sizeof(T), cudaHostAllocDefault) != cudaSuccess) return false;
    if (n) memcpy(np, p, n * sizeof(T));
    if (p) cudaFreeHost(p);
    p = np;
    cap = ncap;
    return true;
  }
  bool append(const T* src, size_t k) {
    if (!reserve(n + k)) return false;
    memcpy(p + n, src, k * sizeof(T));
    n += k;
    return true;
  }
  bool push(const T& v) { return append(&v, 1); }
  void clear() { n = 0; }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    n = cap = 0;
  }
};

}  // namespace

struct z2d_sfc {
  z2d_ctx* ctx;
  uint8_t* data;
  uint32_t fmt;
  int32_t w, h;
  size_t bytes;
  int32_t slot[2];  // index in the surface table of recording batch 0 / 1, or -1
  int32_t y0 = 0, vh = 0;  // band surface: rows [y0, y0 + h) of a canvas vh rows high (ordinary surface: 0, h)
  // band VIEW: `data` points into another surface's memory -- a canvas of this process (parent) or of a peer process (IPC)
  bool external = false;
  void* ipc_base = nullptr;  // cudaIpcOpenMemHandle mapping to close
};

struct BatchMeta {  // shape of the most recently uploaded batch (kept for z2d_replay)
  bool valid = false;
  uint32_t n_draws = 0, n_sp = 0, n_sfc = 0, n_tiles = 0, n_work = 0, n_par_sp = 0, n_chunks = 0, n_strokes = 0, n_srcs = 0, n_unit_sp = 0, max_tiles_y = 0;
  int set = 0;  // which InputSet holds the batch
  size_t n_nodes = 0, h2d_bytes = 0;
};

// one call of a run handled by the parallel recorder (z2d_submit)
struct FillPlan {
  int32_t status;
  uint32_t eligible, skip;  // skip: valid call that records nothing (empty node list)
  uint32_t first, m, n_sp, n_par, all_simple;
  uint32_t node_base, sp_base, draw_idx, slot;
};

struct Batch {  // one recorded command batch (host side)
  int index = 0;
  PinnedVec<z2d_node> nodes;
  PinnedVec<DevSubPath> subpaths;
  PinnedVec<DrawIn> draws;
  PinnedVec<uint8_t> side;        // pinned staging of the side tables (built at flush)
  std::vector<StrokeIn> strokes;  // side tables of the batch
  std::vector<DevSrc> srcs;
  uint32_t iso_mode = 0, iso_node_begin = 0, iso_node_end = 0;  // the isolated draw of a 1-draw batch
  uint32_t n_par_sp = 0;  // sub-paths flagged kSpNodeParallel
  uint32_t n_unit_sp = 0; // sub-paths flagged kSpStrokeUnits
  std::vector<z2d_sfc*> batch_sfcs;
  std::vector<DevGrad> grads;
  std::vector<float> stop_offsets;
  std::vector<float4> stop_colors;
  std::vector<GlyphInst> ginst;  // glyph instances of the batch's text runs (z2d_fill_glyphs): expanded into `nodes` on the device
  std::vector<double> dashes;   // concatenated dash arrays of the batch's dashed strokes
  std::vector<double> pens;     // pen vertices, 6 doubles each {px,py,cw.dx,cw.dy,ccw.dx,ccw.dy}
  // pens already built for this batch, by (thickness, tolerance, ctm): a scene strokes with a handful of widths, and a pen is
  // up to ~60 sin / cos pairs on the host and 48 bytes per vertex over PCIe (config 3: 60 MB per 50 000 strokes before this)
  struct PenEntry {
    double key[8];
    uint32_t begin, count;
  };
  static constexpr int kPenCache = 16;
  PenEntry pen_cache[kPenCache];
  int pen_cache_n = 0, pen_cache_next = 0;
};

struct InputSet {  // device copies of one batch's uploaded inputs; two sets, so that the upload of batch k+1 (copy stream)
                   // overlaps the kernels of batch k (which read the other set)
  DevBuf d_nodes, d_subpaths, d_draws_in;
  DevBuf d_side;  // all small tables of the batch, packed into one pinned blob on the host and uploaded with one copy
  const StrokeIn* strokes = nullptr;
  const DevSrc* srcs = nullptr;
  const DevSurface* sfcs = nullptr;
  const uint32_t* work_base = nullptr;
  const uint32_t* chunk_base = nullptr;
  const DevGrad* grads = nullptr;
  const float* stop_off = nullptr;
  const float4* stop_col = nullptr;
  const void* pens = nullptr;
  const double* dashes = nullptr;
  const GlyphInst* ginst = nullptr;
  cudaEvent_t done = nullptr;  // recorded on the main stream after the last kernel that reads this set
  bool used = false;
  void release() {
    DevBuf* b[] = {&d_nodes, &d_subpaths, &d_draws_in, &d_side};
    for (DevBuf* x : b) x->release();
  }
};

struct z2d_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int sm_count = 148;
  std::string last_error;
  z2d_stats stats{};

  // recorded batches: the application thread records into `rec` while the worker may be executing the other one
  Batch bat[2];
  Batch* rec = &bat[0];
  uint32_t chunk_draws = 32768;
  unsigned record_threads = 4;          // host threads z2d_submit may use for long runs of plain fills
  std::vector<FillPlan> fill_plans;     // scratch of the parallel recorder  // hand the recording batch to the worker every this many draws (0: never)
  std::thread worker;
  std::mutex mu;
  std::condition_variable cv;
  Batch* pending = nullptr;  // batch handed to the worker
  std::atomic<bool> busy{false};  // written under `mu`; read lock-free by the recorder
  bool stop = false;
  int async_rc = 0;          // first error of the batches the worker executed since the last wait

  // device state
  InputSet in[2];
  cudaStream_t copy_stream = nullptr;  // H2D of batch inputs
  cudaStream_t d2h_stream = nullptr;   // z2d_surface_download_async: read-backs that overlap the kernels of later batches
  cudaEvent_t ev_d2h = nullptr;
  bool d2h_pending = false;
  cudaEvent_t ev_up = nullptr;
  DevBuf d_node_sp, d_curve_list, d_sp_order, d_sp_keys;
  DevBuf d_blue, d_draws;
  DevBuf d_sp_count, d_sp_off, d_edges, d_edge_draw, d_draw_bands, d_draw_band_off, d_band_count, d_band_off, d_band_cursor;
  DevBuf d_band_edges, d_list_cnt, d_list_off, d_list_items, d_scan_tmp;
  DevBuf d_comp_grads, d_comp_stop_off, d_comp_stop_col;
  uint32_t* h_total = nullptr;  // pinned readback slot
  DevBuf d_counters, d_boxes, d_hots, d_band_hdr, d_band_xr;
  DevBuf d_export, d_gamma;  // z2d_surface_export: scanline staging, sRGB channel table
  DevBuf d_sim_rows, d_sim_perm, d_sim_x;  // k_edge_sim: row records, per-edge scratch
  // unit stroker (stroke_units.cuh): unit / link records, port points, cursors; capacities in records, kept from batch to batch
  DevBuf d_su_units, d_su_links, d_su_ports, d_su_ctr;
  uint32_t su_unit_cap = 0, su_link_cap = 0, su_edge_cap = 0;
  uint32_t last_scan_edges = 0;  // counted (non-pool) edges of the previous batch: sizes the edge array before the total is known
  bool stroke_units = true;      // Z2D_NO_STROKE_UNITS=1: every stroke through the sub-path stroker
  bool fill_single_pass = true;  // Z2D_NO_FILL_SINGLE_PASS=1: node-parallel fills counted, scanned, then emitted (two passes)
  // glyph cache (z2d_glyph_cache_add): outlines in Path space, host mirror + device copy; per glyph its node range and sub-paths
  struct GlyphEntry {
    uint32_t node_off, n_nodes, sp_off, n_sp;
  };
  std::vector<z2d_node> glyph_nodes;
  std::vector<DevSubPath> glyph_sps;  // node_begin / node_end relative to the glyph
  std::vector<GlyphEntry> glyphs;
  DevBuf d_glyph_nodes;
  size_t glyph_nodes_on_device = 0;
  DevBuf d_small_out;                      // small-batch path: the prepare kernel's result block (abort flag, totals)
  uint32_t* h_small = nullptr;             // its pinned host copy
  bool small_pending = false;              // a small batch is on the stream and its abort flag has not been looked at yet
  bool small_enabled = true;
  cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  BatchMeta last;
  bool stats_pending = false;
};

namespace {

struct NvtxRange {  // host-side range around the launches of one pipeline stage (Nsight Systems timeline)
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

int fail(z2d_ctx* c, const char* what, cudaError_t e) {
  if (c) {
    char buf[256];
    snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString(e));
    c->last_error = buf;
  }
  return Z2D_E_DEVICE;
}
#define CK(c, call)                              \
  do {                                           \
    cudaError_t _e = (call);                     \
    if (_e != cudaSuccess) return fail(c, #call, _e); \
  } while (0)

// ---------------------------------------------------------------- colour (host)
// color.zig:166-195,408-478 -- conversions of gradient stops into the interpolation space.
struct F4 {
  float r, g, b, a;
};
const float kGammaH = 2.2f;
float zmodf_h(float a, float b) {
  float r = std::fmod(a, b);
  if (r < 0) r += b;
  return r;
}
float hsl_channel_h(float n, float hue, float sat, float light) {
  float k = std::fmod(n + hue / 30.0f, 12.0f);
  float a = sat * std::min(light, 1.0f - light);
  return light - a * std::max(-1.0f, std::min(std::min(k - 3.0f, 9.0f - k), 1.0f));
}
F4 hsl_to_rgb_h(F4 h) {
  float hue = std::fmod(h.r, 360.0f);
  if (hue < 0) hue += 360.0f;
  return {hsl_channel_h(0, hue, h.g, h.b), hsl_channel_h(8, hue, h.g, h.b), hsl_channel_h(4, hue, h.g, h.b), h.a};
}
F4 hsl_from_rgb_h(F4 s) {
  float mx = std::max(s.r, std::max(s.g, s.b)), mn = std::min(s.r, std::min(s.g, s.b));
  float range = mx - mn, light = (mn + mx) / 2;
  float sat = (light == 0 || light == 1) ? 0 : (mx - light) / std::min(light, 1 - light);
  float hue = 0;
  if (range != 0) {
    if (mx == s.r) hue = 60 * zmodf_h((s.g - s.b) / range, 6);
    else if (mx == s.g) hue = 60 * ((s.b - s.r) / range + 2);
    else if (mx == s.b) hue = 60 * ((s.r - s.g) / range + 4);
  }
  if (sat < 0) {
    hue += 180;
    if (hue >= 360) hue -= 360;
    sat = std::fabs(sat);
  }
  return {hue, sat, light, s.a};
}
F4 color_to_linear_h(const z2d_color& c) {
  F4 v{c.c[0], c.c[1], c.c[2], c.c[3]};
  if (c.space == Z2D_COLOR_LINEAR_RGB) return v;
  if (c.space == Z2D_COLOR_SRGB) return {std::pow(v.r, kGammaH), std::pow(v.g, kGammaH), std::pow(v.b, kGammaH), v.a};
  return hsl_to_rgb_h(v);
}
F4 color_to_srgb_h(const z2d_color& c) {
  F4 v{c.c[0], c.c[1], c.c[2], c.c[3]};
  if (c.space == Z2D_COLOR_SRGB) return v;
  F4 lin = (c.space == Z2D_COLOR_LINEAR_RGB) ? v : hsl_to_rgb_h(v);
  const float inv = 1 / kGammaH;
  return {std::pow(lin.r, inv), std::pow(lin.g, inv), std::pow(lin.b, inv), lin.a};
}
F4 color_to_hsl_h(const z2d_color& c) {
  if (c.space == Z2D_COLOR_HSL) return {c.c[0], c.c[1], c.c[2], c.c[3]};
  return hsl_from_rgb_h(color_to_linear_h(c));
}

// Gradient.init + Radial/Conic pre-calculation (gradient.zig:262-300, 689)
uint32_t add_gradient(const z2d_gradient& g, std::vector<DevGrad>& grads, std::vector<float>& offs, std::vector<float4>& cols) {
  DevGrad o{};
  o.type = g.type;
  o.method = g.method;
  o.polar = g.polar;
  o.n_stops = g.n_stops;
  o.stop_base = (uint32_t)offs.size();
  for (int i = 0; i < 6; i++) {
    o.geom[i] = g.geom[i];
    o.inv[i] = g.inv_ctm[i];
  }
  o.inv_identity = (g.inv_ctm[0] == 1 && g.inv_ctm[1] == 0 && g.inv_ctm[2] == 0 && g.inv_ctm[3] == 1 && g.inv_ctm[4] == 0 && g.inv_ctm[5] == 0);
  if (g.type == Z2D_GRADIENT_RADIAL) {
    o.inner_r = std::max(0.0, g.geom[2]);
    o.outer_r = std::max(0.0, g.geom[5]);
    o.cdx = g.geom[3] - g.geom[0];
    o.cdy = g.geom[4] - g.geom[1];
    o.dr = o.outer_r - o.inner_r;
    o.min_dr = -o.inner_r;
    double a = 0;
    a += o.cdx * o.cdx;
    a += o.cdy * o.cdy;
    a += o.dr * -o.dr;
    o.a = a;
    o.inv_a = (a != 0) ? 1 / a : 0;
  } else if (g.type == Z2D_GRADIENT_CONIC) {
    const double two_pi = M_PI * 2;
    double r = std::fmod(g.geom[2], two_pi);
    if (r < 0) r += two_pi;
    o.geom[2] = r;
  }
  for (uint32_t i = 0; i < g.n_stops; i++) {
    const z2d_color& c = g.stops[i].color;
    F4 v = g.method == Z2D_INTERP_LINEAR_RGB ? color_to_linear_h(c) : g.method == Z2D_INTERP_SRGB ? color_to_srgb_h(c) : color_to_hsl_h(c);
    offs.push_back(g.stops[i].offset);
    cols.push_back(make_float4(v.r, v.g, v.b, v.a));
  }
  grads.push_back(o);
  return (uint32_t)grads.size() - 1;
}

bool px_can_demultiply(const z2d_pixel& px) {  // pixel.zig:504-514
  if (px.format != Z2D_FMT_ARGB && px.format != Z2D_FMT_RGBA) return true;
  if (px.a == 0) return true;
  return px.r * 255 / px.a <= 255 && px.g * 255 / px.a <= 255 && px.b * 255 / px.a <= 255;
}
bool px_is_opaque(const z2d_pixel& px) {  // pixel.zig:128-137
  switch (px.format) {
    case Z2D_FMT_XRGB: case Z2D_FMT_RGB: return true;
    case Z2D_FMT_ARGB: case Z2D_FMT_RGBA: case Z2D_FMT_ALPHA8: return px.a == 255;
    case Z2D_FMT_ALPHA4: return px.a == 15;
    case Z2D_FMT_ALPHA2: return px.a == 3;
    default: return px.a == 1;
  }
}

int pattern_to_src(const z2d_pattern& p, DevSrc& s, std::vector<DevGrad>& grads, std::vector<float>& offs, std::vector<float4>& cols) {
  memset(&s, 0, sizeof s);
  auto set_pixel = [&](const z2d_pixel& px) {
    RGBA16 v = pixel_to_rgba16(px.format, px.r, px.g, px.b, px.a);
    s.px_rgba = (uint32_t)v.r | ((uint32_t)v.g << 8) | ((uint32_t)v.b << 16) | ((uint32_t)v.a << 24);
    s.px_format = px.format;
    s.px_raw = (uint32_t)px.r | ((uint32_t)px.g << 8) | ((uint32_t)px.b << 16) | ((uint32_t)px.a << 24);
    return v;
  };
  switch (p.kind) {
    case Z2D_PATTERN_OPAQUE:
      if (p.pixel.format > Z2D_FMT_ALPHA1) return Z2D_E_INVALID_ARG;
      s.kind = Z2D_PARAM_PIXEL;
      set_pixel(p.pixel);
      return Z2D_OK;
    case Z2D_PATTERN_GRADIENT:
      if (!p.gradient) return Z2D_E_INVALID_ARG;
      s.kind = Z2D_PARAM_GRADIENT;
      s.grad = add_gradient(*p.gradient, grads, offs, cols);
      return Z2D_OK;
    case Z2D_PATTERN_DITHER:
      s.kind = Z2D_PARAM_DITHER;
      s.dither_type = p.dither_type;
      s.dither_source = p.dither_source;
      s.dither_scale = p.dither_scale;
      if (p.dither_source == Z2D_DITHER_SRC_GRADIENT) {
        if (!p.gradient) return Z2D_E_INVALID_ARG;
        s.grad = add_gradient(*p.gradient, grads, offs, cols);
      } else if (p.dither_source == Z2D_DITHER_SRC_COLOR) {
        F4 c = color_to_linear_h(p.dither_color);
        s.dcol[0] = c.r; s.dcol[1] = c.g; s.dcol[2] = c.b; s.dcol[3] = c.a;
      } else {  // LinearRGB.decodeRGBA (color.zig:204-212): integer de-multiply, then /255
        RGBA16 v = set_pixel(p.pixel);
        if (v.a == 0) v = {0, 0, 0, 0};
        else v = {v.r * 255 / v.a, v.g * 255 / v.a, v.b * 255 / v.a, v.a};
        s.dcol[0] = (float)v.r / 255.0f; s.dcol[1] = (float)v.g / 255.0f; s.dcol[2] = (float)v.b / 255.0f; s.dcol[3] = (float)v.a / 255.0f;
      }
      return Z2D_OK;
    default: return Z2D_E_INVALID_ARG;
  }
}


// Transformation.inverse (Transformation.zig:103-162); the caller has already checked invertibility
void invert_ctm(const double* m, double* o) {
  const double ax = m[0], by = m[1], cx = m[2], dy = m[3], tx = m[4], ty = m[5];
  if (by == 0 && cx == 0) {
    if (ax != 1 || dy != 1) {
      o[0] = 1 / ax; o[1] = 0; o[2] = 0; o[3] = 1 / dy; o[4] = -tx / ax; o[5] = -ty / dy;
    } else {
      o[0] = 1; o[1] = 0; o[2] = 0; o[3] = 1; o[4] = -tx; o[5] = -ty;
    }
    return;
  }
  const double det = ax * dy - by * cx;
  const double k = 1 / det;
  o[0] = dy * k; o[1] = -by * k; o[2] = -cx * k; o[3] = ax * k;
  o[4] = (by * ty - dy * tx) * k;
  o[5] = (cx * tx - ax * ty) * k;
}

// arc.transformed_circle_major_axis (internal/arc.zig:92-267)
double major_axis(const double* m, double radius) {
  const double eps = 0.00390625;
  const double det = m[0] * m[3] - m[1] * m[2];
  if (std::fabs(det * det - 1.0) < eps) {
    if (std::fabs(m[1]) < eps && std::fabs(m[2]) < eps) return radius;
    if (std::fabs(m[0]) < eps && std::fabs(m[3]) < eps) return radius;
  }
  const double i = m[0] * m[0] + m[1] * m[1], j = m[2] * m[2] + m[3] * m[3];
  const double f = 0.5 * (i + j), g = 0.5 * (i - j), h = m[0] * m[2] + m[1] * m[3];
  return radius * std::sqrt(f + std::hypot(g, h));
}

// Pen.init (tess/Pen.zig:36-129).  Built on the host with the C library's acos/cos/sin (the values the CPU
// path uses) once per distinct (thickness, tolerance, CTM); the device only looks vertices up.
void add_pen(z2d_ctx* c, DevDraw& d) {
  const double key[8] = {d.thickness, d.tolerance, d.ctm[0], d.ctm[1], d.ctm[2], d.ctm[3], d.ctm[4], d.ctm[5]};
  for (int k = 0; k < c->rec->pen_cache_n; k++) {
    const Batch::PenEntry& e = c->rec->pen_cache[k];
    if (memcmp(key, e.key, sizeof key) == 0) {
      d.pen_begin = e.begin;
      d.pen_count = e.count;
      return;
    }
  }
  const double radius = d.thickness / 2, tol = d.tolerance;
  int n;
  const double major = major_axis(d.ctm, radius);
  if (tol >= major * 4) {
    n = 1;
  } else if (tol >= major) {
    n = 4;
  } else {
    const double delta = std::acos(1 - tol / major);
    if (delta == 0) {
      n = 4;
    } else {
      n = (int)std::ceil(2 * M_PI / delta);
      if (n < 4) n = 4;
      else if (n % 2 != 0) n = n + 1;
    }
  }
  const bool reflect = (d.ctm[0] * d.ctm[3] - d.ctm[1] * d.ctm[2]) < 0;
  const size_t base = c->rec->pens.size();
  c->rec->pens.resize(base + (size_t)n * 6);
  double* v = c->rec->pens.data() + base;
  for (int i = 0; i < n; i++) {
    double t = 2 * M_PI * (double)i / (double)n;
    if (reflect) t = -t;
    const double dx = radius * std::cos(t), dy = radius * std::sin(t);
    v[i * 6 + 0] = d.ctm[0] * dx + d.ctm[1] * dy;
    v[i * 6 + 1] = d.ctm[2] * dx + d.ctm[3] * dy;
  }
  for (int i = 0; i < n; i++) {
    const int next = (i >= n - 1) ? 0 : i + 1;
    const int prev = std::max(0, i == 0 ? n - 1 : i - 1);
    v[i * 6 + 2] = v[i * 6 + 0] - v[prev * 6 + 0];  // slope_cw = Slope.init(prev, this)
    v[i * 6 + 3] = v[i * 6 + 1] - v[prev * 6 + 1];
    v[i * 6 + 4] = v[next * 6 + 0] - v[i * 6 + 0];  // slope_ccw = Slope.init(this, next)
    v[i * 6 + 5] = v[next * 6 + 1] - v[i * 6 + 1];
  }
  d.pen_begin = (uint32_t)(base / 6);
  d.pen_count = (uint32_t)n;
  Batch::PenEntry& e = c->rec->pen_cache[c->rec->pen_cache_next];  // round robin once the cache is full
  memcpy(e.key, key, sizeof key);
  e.begin = d.pen_begin;
  e.count = d.pen_count;
  c->rec->pen_cache_next = (c->rec->pen_cache_next + 1) % Batch::kPenCache;
  if (c->rec->pen_cache_n < Batch::kPenCache) c->rec->pen_cache_n++;
}

cudaError_t upload(z2d_ctx* c, DevBuf& b, const void* src, size_t bytes) {
  cudaError_t e = b.ensure(bytes ? bytes : 16);
  if (e != cudaSuccess) return e;
  if (bytes) return cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, c->stream);
  return cudaSuccess;
}

int read_total(z2d_ctx* c, const uint32_t* dev, uint32_t& out) {
  CK(c, cudaMemcpyAsync(c->h_total, dev, 4, cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  out = *c->h_total;
  return Z2D_OK;
}

void clear_batch(z2d_ctx* c, Batch& B) {
  B.nodes.clear();
  B.subpaths.clear();
  B.draws.clear();
  B.n_par_sp = 0;
  B.n_unit_sp = 0;
  B.strokes.clear();
  B.srcs.clear();
  (void)c;
  for (z2d_sfc* s : B.batch_sfcs) s->slot[B.index] = -1;
  B.batch_sfcs.clear();
  B.grads.clear();
  B.stop_offsets.clear();
  B.stop_colors.clear();
  B.dashes.clear();
  B.ginst.clear();
  B.pens.clear();
  B.pen_cache_n = 0;
  B.pen_cache_next = 0;
}

GradTables tables(z2d_ctx* c, const DevGrad* g, const float* so, const float4* sc) {
  GradTables T;
  T.grads = g;
  T.stop_offsets = so;
  T.stop_colors = sc;
  T.blue_noise = c->d_blue.as<uint16_t>();
  return T;
}

GradTables tables(z2d_ctx* c, const DevBuf& g, const DevBuf& so, const DevBuf& sc) {
  GradTables T;
  T.grads = g.as<DevGrad>();
  T.stop_offsets = so.as<float>();
  T.stop_colors = sc.as<float4>();
  T.blue_noise = c->d_blue.as<uint16_t>();
  return T;
}

// ------------------------------------------------------------------ the pipeline
// The device stage of a batch: everything it reads is device resident (replayable).
int run_pipeline(z2d_ctx* c, bool replay) {
  const BatchMeta& m = c->last;
  if (!m.valid || m.n_draws == 0) return Z2D_OK;
  InputSet& S = c->in[m.set];
  cudaStream_t st = c->stream;
  const uint32_t n_draws = m.n_draws, n_sp = m.n_sp, n_sfc = m.n_sfc, n_work = m.n_work;
  uint32_t launches = 0;
  auto scan = [&](DevBuf& in, DevBuf& out, uint32_t n) -> cudaError_t {
    cudaError_t e = out.ensure(((size_t)n + 1) * 4);
    if (e != cudaSuccess) return e;
    e = c->d_scan_tmp.ensure(scan_tmp_len(n) * 4);
    if (e != cudaSuccess) return e;
    exclusive_scan(in.as<uint32_t>(), out.as<uint32_t>(), n, c->d_scan_tmp.as<uint32_t>(), st);
    launches += 3;
    return cudaGetLastError();
  };
  NvtxRange r_batch("z2d batch");
  int su_tries = 0;
  bool par_pool = m.n_par_sp != 0 && c->fill_single_pass;  // node-parallel fills flattened in one pass into the edge pool
restart:
  nvtxRangePushA("z2d K0-K1 expand + flatten (count)");
  CK(c, c->d_counters.ensure(64));
  CK(c, cudaMemsetAsync(c->d_counters.p, 0, 64, st));
  CK(c, cudaEventRecord(c->ev[0], st));
  launch_expand_draws(S.d_draws_in.as<DrawIn>(), S.strokes, S.srcs, c->d_draws.as<DevDraw>(), n_draws, st);

  // K1: flatten (count, scan, emit).  Count slots: one per sub-path (sequential plotters: strokes, irregular fills) followed,
  // when the batch has node-parallel sub-paths, by one per node.
  const uint32_t n_nodes = (uint32_t)m.n_nodes;
  const bool par = m.n_par_sp != 0;
  const uint32_t n_cnt = n_sp + (par ? n_nodes : 0u);
  CK(c, c->d_sp_count.ensure((size_t)n_cnt * 4 + 16));
  // batches with strokes: threads take sub-paths in style order so that the lanes of a warp follow the same branches
  const uint32_t* order = nullptr;
  if (m.n_strokes >= 64) {
    CK(c, c->d_sp_order.ensure((size_t)n_sp * 4 + 16));
    CK(c, c->d_sp_keys.ensure(1024 * 4));
    launch_sp_order(S.d_subpaths.as<DevSubPath>(), n_sp, c->d_draws.as<DevDraw>(), c->d_sp_keys.as<uint32_t>(), c->d_sp_order.as<uint32_t>(), st);
    order = c->d_sp_order.as<uint32_t>();
    launches += 3;
  }
  launch_flatten_count(S.d_subpaths.as<DevSubPath>(), n_sp, S.d_nodes.as<z2d_node>(), c->d_draws.as<DevDraw>(), c->d_sp_count.as<uint32_t>(),
                       S.pens, S.dashes, order, st);
  if (par) {
    CK(c, c->d_node_sp.ensure((size_t)n_nodes * 4 + 16));
    CK(c, c->d_curve_list.ensure(((size_t)n_nodes + 1) * 4 + 16));
    if (par_pool)  // one pass: upper bounds now, edges (and extents) into the pool below
      launch_node_bounds(S.d_subpaths.as<DevSubPath>(), n_sp, c->d_node_sp.as<uint32_t>(), n_nodes, S.d_nodes.as<z2d_node>(),
                         c->d_draws.as<DevDraw>(), c->d_sp_count.as<uint32_t>() + n_sp, c->d_curve_list.as<uint32_t>(), st);
    else
      launch_flatten_nodes(false, S.d_subpaths.as<DevSubPath>(), n_sp, c->d_node_sp.as<uint32_t>(), n_nodes, S.d_nodes.as<z2d_node>(),
                           c->d_draws.as<DevDraw>(), c->d_sp_count.as<uint32_t>() + n_sp, nullptr, nullptr, nullptr,
                           c->d_curve_list.as<uint32_t>(), st);
    launches += 5;
  }
  CK(c, scan(c->d_sp_count, c->d_sp_off, n_cnt));
  // K1 for ordinary strokes: the unit stroker writes its edges straight into the pool at the FRONT of the edge array, so the
  // array is sized before the counted total is known (the previous batch's, corrected after the read-back below)
  const bool units = m.n_unit_sp != 0;
  const bool pool = units || par_pool;
  if (pool) {
    if (c->su_edge_cap == 0) c->su_edge_cap = 32u * n_nodes + 65536u;
    CK(c, c->d_su_ctr.ensure(64));
    CK(c, cudaMemsetAsync(c->d_su_ctr.p, 0, 32, st));
    CK(c, c->d_edges.ensure(((size_t)c->su_edge_cap + c->last_scan_edges) * sizeof(DevEdge) + 32));
    CK(c, c->d_edge_draw.ensure(((size_t)c->su_edge_cap + c->last_scan_edges) * 4 + 16));
  }
  if (par_pool) {
    launch_flatten_nodes_pool(S.d_subpaths.as<DevSubPath>(), c->d_node_sp.as<uint32_t>(), n_nodes, S.d_nodes.as<z2d_node>(),
                              c->d_draws.as<DevDraw>(), c->d_sp_off.as<uint32_t>() + n_sp, c->d_su_ctr.as<uint32_t>(), c->d_edges.as<DevEdge>(),
                              c->d_edge_draw.as<uint32_t>(), c->su_edge_cap, c->d_curve_list.as<uint32_t>(), st);
    launches += 3;
  }
  if (units) {
    if (c->su_unit_cap == 0) {
      c->su_unit_cap = 4u * n_nodes + 4096u;
      c->su_link_cap = 2u * c->su_unit_cap;
    }
    CK(c, c->d_su_units.ensure((size_t)c->su_unit_cap * kStrokeUnitBytes));
    CK(c, c->d_su_links.ensure((size_t)c->su_link_cap * kStrokeLinkBytes));
    CK(c, c->d_su_ports.ensure((size_t)c->su_unit_cap * kStrokePortBytes));
    launch_stroke_units(S.d_subpaths.as<DevSubPath>(), n_sp, S.d_nodes.as<z2d_node>(), c->d_draws.as<DevDraw>(), S.pens, S.dashes, order,
                        c->d_su_units.p, c->su_unit_cap, c->d_su_links.p, c->su_link_cap, c->d_su_ports.p, c->d_su_ctr.as<uint32_t>(),
                        c->d_edges.as<DevEdge>(), c->d_edge_draw.as<uint32_t>(), c->su_edge_cap, st);
    launches += 4;
  }

  nvtxRangePop();
  nvtxRangePushA("z2d K2 setup + K3b list sizes + read-back");
  // K2: per-draw regions and (draw, tile-row) slots; K3b (count half): sizes of the per-tile-row draw lists.  Both only need
  // the extents gathered by the count pass, so they run before the edges exist and the three totals the host needs for
  // allocation come back in ONE round trip.
  CK(c, c->d_draw_bands.ensure((size_t)n_draws * 4 + 16));
  CK(c, c->d_boxes.ensure((size_t)n_draws * sizeof(DrawBox) + 16));
  CK(c, c->d_hots.ensure((size_t)n_draws * sizeof(DrawHot) + 16));
  launch_setup_draws(c->d_draws.as<DevDraw>(), n_draws, S.sfcs, c->d_draw_bands.as<uint32_t>(), c->d_boxes.as<DrawBox>(),
                     c->d_counters.as<unsigned long long>(), st);
  CK(c, scan(c->d_draw_bands, c->d_draw_band_off, n_draws));
  CK(c, c->d_list_cnt.ensure((size_t)n_work * 4 + 16));
  launch_band_lists(false, S.sfcs, n_sfc, S.work_base, S.chunk_base, m.n_chunks,
                    c->d_boxes.as<DrawBox>(), c->d_list_cnt.as<uint32_t>(), nullptr, nullptr, nullptr, nullptr, m.max_tiles_y, st);
  CK(c, scan(c->d_list_cnt, c->d_list_off, n_work));
  // counted edges: sequential sub-paths (+ nodes, when they are not in the pool)
  CK(c, cudaMemcpyAsync(c->h_total + 0, c->d_sp_off.as<uint32_t>() + (par_pool ? n_sp : n_cnt), 4, cudaMemcpyDeviceToHost, st));
  CK(c, cudaMemcpyAsync(c->h_total + 1, c->d_draw_band_off.as<uint32_t>() + n_draws, 4, cudaMemcpyDeviceToHost, st));
  CK(c, cudaMemcpyAsync(c->h_total + 2, c->d_list_off.as<uint32_t>() + n_work, 4, cudaMemcpyDeviceToHost, st));
  CK(c, cudaMemcpyAsync(c->h_total + 4, c->d_counters.as<unsigned long long>() + 4, 16, cudaMemcpyDeviceToHost, st));  // k_edge_sim sizes
  if (pool) CK(c, cudaMemcpyAsync(c->h_total + 8, c->d_su_ctr.p, 20, cudaMemcpyDeviceToHost, st));
  CK(c, cudaStreamSynchronize(st));
  uint32_t pool_edges = 0;
  if (pool) {
    const uint32_t need_u = c->h_total[8], need_l = c->h_total[9], need_e = c->h_total[10];
    if (par_pool && (c->h_total[12] != 0u || need_e > (1u << 30))) {
      // a node produced more edges than its bound (NaN / infinite coordinates), or the bounds add up to an absurd array
      // (tolerance ~ 0): this batch goes through the two-pass kernels
      par_pool = false;
      nvtxRangePop();
      goto restart;
    }
    if ((units && (need_u > c->su_unit_cap || need_l > c->su_link_cap)) || need_e > c->su_edge_cap) {
      // a capacity was too small (first batch with strokes, or denser strokes than before): enlarge and redo the batch from its
      // resident inputs.  When the unit array overflowed the edge need is only a lower bound, so this can take a second round.
      if (++su_tries > 4) return fail(c, "stroke unit capacities", cudaErrorUnknown);
      auto grow = [](uint32_t cap, uint32_t need) { return need > cap ? need + need / 8u + 1024u : cap; };
      const bool units_lost = units && need_u > c->su_unit_cap;
      if (units) {
        c->su_unit_cap = grow(c->su_unit_cap, need_u);
        c->su_link_cap = grow(c->su_link_cap, need_l);
      }
      c->su_edge_cap = grow(c->su_edge_cap, units_lost ? std::max(need_e, 12u * need_u) : need_e);
      nvtxRangePop();
      goto restart;
    }
    pool_edges = need_e;
  }
  const uint32_t n_scan_edges = c->h_total[0], n_slots = c->h_total[1], n_items = c->h_total[2];
  const uint32_t n_edges = pool_edges + n_scan_edges;
  c->last_scan_edges = n_scan_edges;
  const unsigned long long sim_rows = *reinterpret_cast<unsigned long long*>(c->h_total + 4),
                           sim_slots = *reinterpret_cast<unsigned long long*>(c->h_total + 6);

  nvtxRangePop();
  nvtxRangePushA("z2d K1 flatten (emit) + edge replay");
  // K1 (emit half)
  // (the pool part [0, pool_edges) is already written: a reallocation keeps it)
  CK(c, c->d_edges.ensure_keep((size_t)n_edges * sizeof(DevEdge) + 32, (size_t)pool_edges * sizeof(DevEdge), st));
  CK(c, c->d_edge_draw.ensure_keep((size_t)n_edges * 4 + 16, (size_t)pool_edges * 4, st));
  DevEdge* const scan_edges = c->d_edges.as<DevEdge>() + pool_edges;  // counted ranges follow the pool
  uint32_t* const scan_edge_draw = c->d_edge_draw.as<uint32_t>() + pool_edges;
  CK(c, c->d_list_items.ensure((size_t)n_items * sizeof(uint4) + 16));
  CK(c, c->d_band_xr.ensure((size_t)n_slots * sizeof(uint2) + 16));
  CK(c, c->d_band_count.ensure((size_t)n_slots * 4 + 16));
  CK(c, c->d_band_cursor.ensure((size_t)n_slots * 4 + 16));
  launch_flatten_emit(S.d_subpaths.as<DevSubPath>(), n_sp, S.d_nodes.as<z2d_node>(), c->d_draws.as<DevDraw>(), c->d_sp_off.as<uint32_t>(),
                      scan_edges, scan_edge_draw, S.pens, S.dashes, order, st);
  if (par && !par_pool) {
    launch_flatten_nodes(true, S.d_subpaths.as<DevSubPath>(), n_sp, c->d_node_sp.as<uint32_t>(), n_nodes, S.d_nodes.as<z2d_node>(),
                         c->d_draws.as<DevDraw>(), nullptr, c->d_sp_off.as<uint32_t>() + n_sp, scan_edges,
                         scan_edge_draw, c->d_curve_list.as<uint32_t>(), st);
    launches += 2;
  }
  if (sim_rows) {  // order-dependent draws (dangling edges, direct rasteriser with an unbounded operator): exact scanline replay
    CK(c, c->d_sim_rows.ensure((size_t)sim_rows * sizeof(int4)));
    CK(c, c->d_sim_perm.ensure((size_t)sim_slots * 4 + 16));
    CK(c, c->d_sim_x.ensure((size_t)sim_slots * 4 + 16));
    launch_edge_sim(c->d_draws.as<DevDraw>(), n_draws, S.sfcs, scan_edges, c->d_sp_off.as<uint32_t>(),
                    c->d_sim_perm.as<uint32_t>(), c->d_sim_x.as<int32_t>(), c->d_sim_rows.as<int4>(), st);
    launches += 1;
  }
  CK(c, cudaEventRecord(c->ev[1], st));

  nvtxRangePop();
  nvtxRangePushA("z2d K3 bin edges + tile-row lists");
  // K3a: edges -> (draw, tile-row) lists
  launch_assign_band_base(c->d_draws.as<DevDraw>(), n_draws, c->d_draw_band_off.as<uint32_t>(), c->d_hots.as<DrawHot>(), c->d_boxes.as<DrawBox>(),
                          S.sfcs, st);
  CK(c, cudaMemsetAsync(c->d_band_count.p, 0, (size_t)n_slots * 4 + 16, st));
  CK(c, cudaMemsetAsync(c->d_band_cursor.p, 0, (size_t)n_slots * 4 + 16, st));
  CK(c, cudaMemsetAsync(c->d_band_xr.p, 0, (size_t)n_slots * sizeof(uint2) + 16, st));
  launch_bin_count(c->d_edges.as<DevEdge>(), c->d_edge_draw.as<uint32_t>(), n_edges, c->d_draws.as<DevDraw>(), c->d_band_count.as<uint32_t>(), st);
  CK(c, scan(c->d_band_count, c->d_band_off, n_slots));
  uint32_t n_band_edges = 0;
  {
    int rc = read_total(c, c->d_band_off.as<uint32_t>() + n_slots, n_band_edges);
    if (rc) return rc;
  }
  CK(c, c->d_band_edges.ensure((size_t)n_band_edges * sizeof(DevEdge) + 32));
  CK(c, c->d_band_hdr.ensure((size_t)n_band_edges * sizeof(int4) + 32));
  launch_bin_scatter(c->d_edges.as<DevEdge>(), c->d_edge_draw.as<uint32_t>(), n_edges, c->d_draws.as<DevDraw>(), c->d_band_off.as<uint32_t>(),
                     c->d_band_cursor.as<uint32_t>(), c->d_band_edges.as<DevEdge>(), c->d_band_hdr.as<int4>(), c->d_band_xr.as<uint2>(), st);
  CK(c, cudaEventRecord(c->ev[2], st));

  // K3b (write half): ordered draw list per surface tile-row
  launch_band_lists(true, S.sfcs, n_sfc, S.work_base, S.chunk_base, m.n_chunks,
                    c->d_boxes.as<DrawBox>(), nullptr, c->d_list_off.as<uint32_t>(), c->d_list_items.as<uint4>(), c->d_band_off.as<uint32_t>(),
                    c->d_band_xr.as<uint2>(), m.max_tiles_y, st);
  CK(c, cudaEventRecord(c->ev[3], st));

  nvtxRangePop();
  nvtxRangePushA("z2d K4 raster tiles");
  // K4: fused coverage + compositing
  RasterArgs A;
  A.sfcs = S.sfcs;
  A.n_sfc = n_sfc;
  A.n_tiles = m.n_tiles;
  A.work_base = S.work_base;
  A.list_off = c->d_list_off.as<uint32_t>();
  A.list_items = c->d_list_items.as<uint4>();
  A.draws = c->d_draws.as<DevDraw>();
  A.hots = c->d_hots.as<DrawHot>();
  A.band_off = c->d_band_off.as<uint32_t>();
  A.band_edges = c->d_band_edges.as<DevEdge>();
  A.band_hdr = c->d_band_hdr.as<int4>();
  A.counters = c->d_counters.as<unsigned long long>();
  A.sim_rows = c->d_sim_rows.as<int4>();
  A.abort = nullptr;
  A.T = tables(c, S.grads, S.stop_off, S.stop_col);
  launch_raster(A, m.n_strokes != 0 || m.n_srcs != 0, st);
  CK(c, cudaGetLastError());
  CK(c, cudaEventRecord(c->ev[4], st));
  nvtxRangePop();
  launches += 9;

  z2d_stats& s = c->stats;
  memset(&s, 0, sizeof s);
  s.draws = n_draws;
  s.nodes = m.n_nodes;
  s.edges = n_edges;
  s.band_edges = n_band_edges;
  s.tile_items = n_items;
  s.tiles = m.n_tiles;
  s.kernel_launches = launches;
  s.h2d_bytes = replay ? 0 : m.h2d_bytes;
  c->stats_pending = true;
  return Z2D_OK;
}

// Small batches: everything before the raster kernel in ONE single-CTA launch into fixed-capacity buffers (smallbatch.cuh),
// no host read-back.  The prepare kernel may give up (capacity, or a draw that needs the scanline replay): check_small, run
// before the context is used again, then redoes the batch with the sized pipeline -- its inputs are still resident and the
// raster kernel of the abandoned attempt composited nothing.
int run_small(z2d_ctx* c) {
  const BatchMeta& m = c->last;
  InputSet& S = c->in[m.set];
  cudaStream_t st = c->stream;
  const uint32_t n_nodes = (uint32_t)m.n_nodes;
  CK(c, c->d_counters.ensure(64));
  CK(c, c->d_small_out.ensure(32));
  CK(c, c->d_hots.ensure((size_t)kSmallMaxDraws * sizeof(DrawHot)));
  CK(c, c->d_boxes.ensure((size_t)kSmallMaxDraws * sizeof(DrawBox)));
  CK(c, c->d_draw_bands.ensure(((size_t)kSmallMaxDraws + 1) * 4));
  CK(c, c->d_node_sp.ensure((size_t)kSmallMaxNodes * 4));
  CK(c, c->d_sp_count.ensure(((size_t)kSmallMaxSubPaths + kSmallMaxNodes + 1) * 4));
  CK(c, c->d_edges.ensure((size_t)kSmallEdgeCap * sizeof(DevEdge)));
  CK(c, c->d_edge_draw.ensure((size_t)kSmallEdgeCap * 4));
  CK(c, c->d_band_count.ensure(((size_t)kSmallSlotCap + 1) * 4));
  CK(c, c->d_band_cursor.ensure(((size_t)kSmallSlotCap + 1) * 4));
  CK(c, c->d_band_xr.ensure(((size_t)kSmallSlotCap + 1) * sizeof(uint2)));
  CK(c, c->d_band_edges.ensure((size_t)kSmallBandCap * sizeof(DevEdge)));
  CK(c, c->d_band_hdr.ensure((size_t)kSmallBandCap * sizeof(int4)));
  CK(c, c->d_list_cnt.ensure(((size_t)kSmallMaxWork + 1) * 4));
  CK(c, c->d_list_items.ensure((size_t)kSmallItemCap * sizeof(uint4)));
  SmallArgs P;
  P.draws_in = S.d_draws_in.as<DrawIn>();
  P.strokes = S.strokes;
  P.srcs = S.srcs;
  P.sps = S.d_subpaths.as<DevSubPath>();
  P.nodes = S.d_nodes.as<z2d_node>();
  P.sfcs = S.sfcs;
  P.work_base = S.work_base;
  P.n_draws = m.n_draws; P.n_sp = m.n_sp; P.n_nodes = n_nodes; P.n_sfc = m.n_sfc; P.n_work = m.n_work;
  P.draws = c->d_draws.as<DevDraw>();
  P.hots = c->d_hots.as<DrawHot>();
  P.boxes = c->d_boxes.as<DrawBox>();
  P.draw_bands = c->d_draw_bands.as<uint32_t>();
  P.node_sp = c->d_node_sp.as<uint32_t>();
  P.cnt = c->d_sp_count.as<uint32_t>();
  P.edges = c->d_edges.as<DevEdge>();
  P.edge_draw = c->d_edge_draw.as<uint32_t>();
  P.band_count = c->d_band_count.as<uint32_t>();
  P.band_cursor = c->d_band_cursor.as<uint32_t>();
  P.band_xr = c->d_band_xr.as<uint2>();
  P.band_edges = c->d_band_edges.as<DevEdge>();
  P.band_hdr = c->d_band_hdr.as<int4>();
  P.list_cnt = c->d_list_cnt.as<uint32_t>();
  P.list_items = c->d_list_items.as<uint4>();
  P.counters = c->d_counters.as<unsigned long long>();
  P.out = c->d_small_out.as<uint32_t>();
  NvtxRange r("z2d small batch");
  CK(c, cudaEventRecord(c->ev[0], st));
  launch_small_batch(P, st);
  for (int i = 1; i <= 3; i++) CK(c, cudaEventRecord(c->ev[i], st));
  RasterArgs A;
  A.sfcs = S.sfcs;
  A.n_sfc = m.n_sfc;
  A.n_tiles = m.n_tiles;
  A.work_base = S.work_base;
  A.list_off = P.list_cnt;
  A.list_items = P.list_items;
  A.draws = P.draws;
  A.hots = P.hots;
  A.band_off = P.band_count;
  A.band_edges = P.band_edges;
  A.band_hdr = P.band_hdr;
  A.counters = P.counters;
  A.sim_rows = nullptr;
  A.abort = P.out;
  A.T = tables(c, S.grads, S.stop_off, S.stop_col);
  launch_raster(A, m.n_srcs != 0 || getenv("Z2D_SMALL_RICH") != nullptr, st);
  CK(c, cudaGetLastError());
  CK(c, cudaEventRecord(c->ev[4], st));
  CK(c, cudaMemcpyAsync(c->h_small, P.out, 32, cudaMemcpyDeviceToHost, st));
  c->small_pending = true;
  z2d_stats& s = c->stats;
  memset(&s, 0, sizeof s);
  s.draws = m.n_draws;
  s.nodes = m.n_nodes;
  s.tiles = m.n_tiles;
  s.kernel_launches = 2;
  s.h2d_bytes = m.h2d_bytes;
  c->stats_pending = true;
  return Z2D_OK;
}

// Looks at the abort flag of the small batch on the stream, if any (blocks until that batch is done), and redoes an
// abandoned batch with the sized pipeline.  Called before anything else touches the context.
int check_small(z2d_ctx* c) {
  if (!c->small_pending) return Z2D_OK;
  c->small_pending = false;
  CK(c, cudaStreamSynchronize(c->stream));
  if (c->h_small[0] != 0u) {
    const int rc = run_pipeline(c, false);
    if (c->last.valid) cudaEventRecord(c->in[c->last.set].done, c->stream);  // the next upload into this input set waits for the redo
    return rc;
  }
  c->stats.edges = c->h_small[1];
  c->stats.tile_items = c->h_small[3];
  c->stats.band_edges = c->h_small[4];
  return Z2D_OK;
}

// Hairline strokes are order-coupled per pixel (slowpath.cuh); they run as a batch of exactly one draw.
int run_isolated(z2d_ctx* c, Batch& B) {
  const BatchMeta& m = c->last;
  InputSet& S = c->in[m.set];
  cudaStream_t st = c->stream;
  const GradTables T = tables(c, S.grads, S.stop_off, S.stop_col);
  uint32_t launches = 0, n_edges = 0;
  CK(c, cudaEventRecord(c->ev[0], st));
  launch_expand_draws(S.d_draws_in.as<DrawIn>(), S.strokes, S.srcs, c->d_draws.as<DevDraw>(), 1, st);
  launch_hairline(S.sfcs, c->d_draws.as<DevDraw>(), 0, S.d_nodes.as<z2d_node>(), B.iso_node_begin, B.iso_node_end, S.dashes, T, st);
  launches = 2;
  CK(c, cudaGetLastError());
  for (int i = 1; i <= 4; i++) CK(c, cudaEventRecord(c->ev[i], st));
  // the pinned batch arrays are reused by the next draw: their async uploads must have landed (the tile pipeline gets this
  // from its read-backs).  Only the copies are waited for -- the hairline kernel itself stays asynchronous, so a hairline in
  // the middle of a scene no longer drains the device.
  CK(c, cudaEventSynchronize(c->ev_up));
  z2d_stats& s = c->stats;
  memset(&s, 0, sizeof s);
  s.draws = 1;
  s.nodes = m.n_nodes;
  s.edges = n_edges;
  s.kernel_launches = launches;
  s.h2d_bytes = m.h2d_bytes;
  c->stats_pending = true;
  c->last.valid = false;  // not replayable
  return Z2D_OK;
}

int flush_impl(z2d_ctx* c, Batch& B) {
  {
    const int rc0 = check_small(c);  // (before c->last is overwritten: an abandoned small batch is redone first)
    if (rc0 != Z2D_OK) return rc0;
  }
  const uint32_t n_draws = (uint32_t)B.draws.n;
  if (n_draws == 0) {
    clear_batch(c, B);
    return Z2D_OK;
  }
  cudaStream_t st = c->stream;
  const uint32_t n_sfc = (uint32_t)B.batch_sfcs.size();

  // 1. group draws by surface, keeping submission order inside each surface (draws on
  //    different surfaces are independent), and build the surface / work tables.
  std::vector<uint32_t> per_sfc(n_sfc + 1, 0);
  bool grouped = true;
  for (uint32_t i = 0; i < n_draws; i++) {
    per_sfc[B.draws.p[i].surface + 1]++;
    grouped &= i == 0 || B.draws.p[i].surface >= B.draws.p[i - 1].surface;
  }
  for (uint32_t s = 0; s < n_sfc; s++) per_sfc[s + 1] += per_sfc[s];
  if (!grouped) {
    std::vector<uint32_t> remap(n_draws), cur(per_sfc.begin(), per_sfc.end() - 1);
    std::vector<DrawIn> sorted(n_draws);
    for (uint32_t i = 0; i < n_draws; i++) {
      uint32_t k = cur[B.draws.p[i].surface]++;
      remap[i] = k;
      sorted[k] = B.draws.p[i];
    }
    for (size_t i = 0; i < B.subpaths.n; i++) B.subpaths.p[i].draw = remap[B.subpaths.p[i].draw];
    memcpy(B.draws.p, sorted.data(), sizeof(DrawIn) * n_draws);
  }

  std::vector<DevSurface> sfcs(n_sfc);
  std::vector<uint32_t> work_base(n_sfc + 1, 0), chunk_base(n_sfc + 1, 0);
  uint32_t n_tiles = 0;
  for (uint32_t s = 0; s < n_sfc; s++) {
    z2d_sfc* hs = B.batch_sfcs[s];
    DevSurface& d = sfcs[s];
    d.data = hs->data;
    d.fmt = hs->fmt;
    d.w = hs->w;
    d.h = hs->h;
    d.tiles_x = (hs->w + kTile - 1) / kTile;
    d.tiles_y = (hs->h + kTile - 1) / kTile;
    d.y0 = hs->y0;
    d.vh = hs->vh;
    d.tile_base = n_tiles;
    d.band_base = 0;
    d.draw_begin = per_sfc[s];
    d.draw_end = per_sfc[s + 1];
    n_tiles += (uint32_t)d.tiles_x * (uint32_t)d.tiles_y;
    const uint32_t chunks = (d.draw_end - d.draw_begin + kDrawChunk - 1) / kDrawChunk;
    work_base[s + 1] = work_base[s] + (uint32_t)d.tiles_y * chunks;
    chunk_base[s + 1] = chunk_base[s] + chunks;
  }
  const uint32_t n_work = work_base[n_sfc];
  const uint32_t n_sp = (uint32_t)B.subpaths.n;

  // 2. upload the batch on the copy stream into this batch's input set, so that the transfer overlaps the kernels of the
  //    previous batch (which read the other set); the set is free once the last batch that used it has passed its raster kernel.
  //    The small tables travel as one pinned blob (sections 256-byte aligned) with typed device pointers into it.
  InputSet& S = c->in[B.index];
  struct Sec { const void* src; size_t bytes; size_t off; };
  Sec sec[11] = {{B.strokes.data(), B.strokes.size() * sizeof(StrokeIn), 0}, {B.srcs.data(), B.srcs.size() * sizeof(DevSrc), 0},
                 {sfcs.data(), sfcs.size() * sizeof(DevSurface), 0},         {work_base.data(), work_base.size() * 4, 0},
                 {chunk_base.data(), chunk_base.size() * 4, 0},             {B.grads.data(), B.grads.size() * sizeof(DevGrad), 0},
                 {B.stop_offsets.data(), B.stop_offsets.size() * 4, 0},     {B.stop_colors.data(), B.stop_colors.size() * sizeof(float4), 0},
                 {B.pens.data(), B.pens.size() * 8, 0},                     {B.dashes.data(), B.dashes.size() * 8, 0},
                 {B.ginst.data(), B.ginst.size() * sizeof(GlyphInst), 0}};
  size_t side_bytes = 0;
  for (Sec& x : sec) {
    x.off = side_bytes;
    side_bytes += (x.bytes + 255) & ~(size_t)255;
  }
  side_bytes += 256;
  if (!B.side.reserve(side_bytes)) return Z2D_E_OUT_OF_MEMORY;
  for (const Sec& x : sec)
    if (x.bytes) memcpy(B.side.p + x.off, x.src, x.bytes);
  CK(c, S.d_nodes.ensure(B.nodes.n * sizeof(z2d_node) + 16));
  CK(c, S.d_subpaths.ensure(B.subpaths.n * sizeof(DevSubPath) + 16));
  CK(c, S.d_draws_in.ensure(B.draws.n * sizeof(DrawIn) + 16));
  CK(c, S.d_side.ensure(side_bytes));
  CK(c, c->d_draws.ensure(B.draws.n * sizeof(DevDraw)));
  if (S.used) CK(c, cudaStreamWaitEvent(c->copy_stream, S.done, 0));
  auto up = [&](DevBuf& b, const void* src, size_t bytes) -> cudaError_t {
    return bytes ? cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, c->copy_stream) : cudaSuccess;
  };
  CK(c, up(S.d_nodes, B.nodes.p, B.nodes.n * sizeof(z2d_node)));
  CK(c, up(S.d_subpaths, B.subpaths.p, B.subpaths.n * sizeof(DevSubPath)));
  CK(c, up(S.d_draws_in, B.draws.p, B.draws.n * sizeof(DrawIn)));
  CK(c, up(S.d_side, B.side.p, side_bytes));
  {
    const uint8_t* base = S.d_side.as<uint8_t>();
    S.strokes = reinterpret_cast<const StrokeIn*>(base + sec[0].off);
    S.srcs = reinterpret_cast<const DevSrc*>(base + sec[1].off);
    S.sfcs = reinterpret_cast<const DevSurface*>(base + sec[2].off);
    S.work_base = reinterpret_cast<const uint32_t*>(base + sec[3].off);
    S.chunk_base = reinterpret_cast<const uint32_t*>(base + sec[4].off);
    S.grads = reinterpret_cast<const DevGrad*>(base + sec[5].off);
    S.stop_off = reinterpret_cast<const float*>(base + sec[6].off);
    S.stop_col = reinterpret_cast<const float4*>(base + sec[7].off);
    S.pens = base + sec[8].off;
    S.dashes = reinterpret_cast<const double*>(base + sec[9].off);
    S.ginst = reinterpret_cast<const GlyphInst*>(base + sec[10].off);
  }
  CK(c, cudaEventRecord(c->ev_up, c->copy_stream));
  CK(c, cudaStreamWaitEvent(c->stream, c->ev_up, 0));
  if (!B.ginst.empty()) {  // text runs: the device writes the transformed outline nodes into the uploaded node array
    if (c->glyph_nodes_on_device != c->glyph_nodes.size()) {
      CK(c, c->d_glyph_nodes.ensure(c->glyph_nodes.size() * sizeof(z2d_node) + 64));
      CK(c, cudaMemcpyAsync(c->d_glyph_nodes.p, c->glyph_nodes.data(), c->glyph_nodes.size() * sizeof(z2d_node), cudaMemcpyHostToDevice, st));
      CK(c, cudaStreamSynchronize(st));  // (pageable source: the copy is staged; the sync only closes the race with a later add)
      c->glyph_nodes_on_device = c->glyph_nodes.size();
    }
    launch_expand_glyphs(S.ginst, (uint32_t)B.ginst.size(), c->d_glyph_nodes.as<z2d_node>(), S.d_nodes.as<z2d_node>(), st);
    CK(c, cudaGetLastError());
  }
  BatchMeta& m = c->last;
  m.valid = true;
  m.set = B.index;
  m.n_draws = n_draws;
  m.n_sp = n_sp;
  m.n_par_sp = B.n_par_sp;
  m.n_unit_sp = B.n_unit_sp;
  m.n_strokes = (uint32_t)B.strokes.size();
  m.n_srcs = (uint32_t)B.srcs.size();
  m.n_sfc = n_sfc;
  m.n_tiles = n_tiles;
  m.max_tiles_y = 0;
  for (const DevSurface& d : sfcs) m.max_tiles_y = std::max(m.max_tiles_y, (uint32_t)d.tiles_y);
  m.n_work = n_work;
  m.n_chunks = chunk_base[n_sfc];
  m.n_nodes = B.nodes.n;
  m.h2d_bytes = B.nodes.n * sizeof(z2d_node) + B.subpaths.n * sizeof(DevSubPath) + B.draws.n * sizeof(DrawIn) + B.strokes.size() * sizeof(StrokeIn) + B.srcs.size() * sizeof(DevSrc) +
                B.pens.size() * 8 + B.dashes.size() * 8 +
                sfcs.size() * sizeof(DevSurface) + work_base.size() * 4 + B.grads.size() * sizeof(DevGrad) +
                B.stop_offsets.size() * 4 + B.stop_colors.size() * sizeof(float4);
  bool small = c->small_enabled && n_draws <= kSmallMaxDraws && B.nodes.n <= kSmallMaxNodes && n_sp <= kSmallMaxSubPaths && n_work <= kSmallMaxWork &&
               B.strokes.empty();
  for (uint32_t i = 0; small && i < n_draws; i++) small = ((B.draws.p[i].opts >> 11) & 3u) == 0u;  // no hairline / row-record draws
  int rc = (n_draws == 1 && ((B.draws.p[0].opts >> 11) & 3u) == 1u) ? run_isolated(c, B) : (small ? run_small(c) : run_pipeline(c, false));
  cudaEventRecord(S.done, c->stream);  // (also on failure: the set is reusable after whatever was enqueued)
  S.used = true;
  clear_batch(c, B);
  return rc;
}

int execute_batch(z2d_ctx* c, Batch& B) {
  int rc = flush_impl(c, B);
  if (rc != Z2D_OK) clear_batch(c, B);
  return rc;
}

// ---- worker: executes handed-over batches so that recording the next chunk overlaps upload + device work.
// Only one of {worker, application thread} touches the device state at a time: every operation on the
// application thread that uses the stream first waits for the worker to go idle.
void worker_main(z2d_ctx* c) {
  cudaSetDevice(c->device);
  std::unique_lock<std::mutex> lk(c->mu);
  for (;;) {
    c->cv.wait(lk, [&] { return c->pending != nullptr || c->stop; });
    if (c->pending == nullptr) return;  // stop requested and nothing queued
    Batch* b = c->pending;
    lk.unlock();
    const int rc = execute_batch(c, *b);
    lk.lock();
    if (rc != Z2D_OK && c->async_rc == Z2D_OK) c->async_rc = rc;
    c->pending = nullptr;
    c->busy = false;
    c->cv.notify_all();
  }
}

int wait_idle(z2d_ctx* c) {  // returns (and clears) the first error of the batches executed by the worker
  std::unique_lock<std::mutex> lk(c->mu);
  c->cv.wait(lk, [&] { return !c->busy; });
  const int rc = c->async_rc;
  c->async_rc = Z2D_OK;
  return rc;
}

// hand the recording batch to the worker and continue recording into the other one
int kick(z2d_ctx* c) {
  if (c->rec->draws.n == 0) return Z2D_OK;
  int rc = wait_idle(c);
  if (!c->worker.joinable()) c->worker = std::thread(worker_main, c);
  {
    std::lock_guard<std::mutex> lk(c->mu);
    c->pending = c->rec;
    c->busy = true;
  }
  c->cv.notify_all();
  c->rec = &c->bat[c->rec->index ^ 1];
  return rc;
}

// everything recorded so far is enqueued on the stream when this returns
int flush(z2d_ctx* c) {
  cudaSetDevice(c->device);
  int rc = wait_idle(c);
  int rc2 = execute_batch(c, *c->rec);
  // A small batch may have been abandoned by its prepare kernel; whatever the caller enqueues or reads next must come after
  // its redo, so the flag is looked at here (one stream sync per small batch instead of the sized pipeline's two read-backs).
  const int rc3 = check_small(c);
  return rc != Z2D_OK ? rc : (rc2 != Z2D_OK ? rc2 : rc3);
}

uint32_t batch_slot(z2d_ctx* c, z2d_sfc* s) {
  int32_t& slot = s->slot[c->rec->index];
  if (slot < 0) {
    slot = (int32_t)c->rec->batch_sfcs.size();
    c->rec->batch_sfcs.push_back(s);
  }
  return (uint32_t)slot;
}

bool is_closed_node_set(const z2d_node* nodes, size_t n) {  // path_nodes.zig:23-37
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

// Split a node list that starts at a move_to into sub-paths (one per move_to; a lone move_to draws nothing).  out == nullptr:
// count only.  Flags sub-paths eligible for node-parallel flattening (kSpNodeParallel): move_to, segments..., one close_path
// at the very end, and at least two segments whose end differs from their start (then the plotter holds >= 3 points at the
// close, fill_plotter.zig:82).
//
// keep_lone (dashed strokes with round or square caps): a sub-path without a single segment that moves -- a lone move_to, or
// move_to + close_path -- is not dropped.  The dashed plotter finishes it as a dot when the dash pattern starts "on"
// (dashed_plotter.zig:334-367 and 266-271, plotDotted 369-465), and a square dot is oriented by the slope the plotter last saw
// -- state that survives move_to (current_slope, dashed_plotter.zig:98).  Such a sub-path is appended to the sub-path before it
// (one thread plots both, in order) or, at the start of the list, becomes a sub-path of its own.
//
// all_simple (optional): every sub-path that draws something qualifies.  A fill call with a sub-path that does not may leave
// a dangling edge (a two-point "polygon"), whose result depends on the order of the call's edges (k_edge_sim): such a call
// is recorded with allow_parallel = false, so that all its edges are emitted by the sequential plotter, in plotting order.
int split_subpaths(const z2d_node* q, size_t m, uint32_t base, uint32_t draw_index, bool allow_parallel, DevSubPath* out, uint32_t& n_sp,
                   uint32_t& n_par, bool keep_lone = false, bool* all_simple = nullptr) {
  n_sp = 0;
  n_par = 0;
  if (all_simple) *all_simple = true;
  size_t i = 0;
  while (i < m) {
    size_t j = i + 1;
    double cx = q[i].p[0], cy = q[i].p[1];
    int moving = 0;
    bool simple = allow_parallel;
    bool any_seg = false, slope_dep = false;  // slope_dep: reaches a dot (close / end) before any segment set the plotter's slope
    while (j < m && q[j].tag != Z2D_NODE_MOVE_TO) {
      const uint32_t tag = q[j].tag;
      if (tag > Z2D_NODE_CLOSE_PATH) return Z2D_E_INVALID_ARG;
      if (tag == Z2D_NODE_CLOSE_PATH) {
        if (j + 1 < m && q[j + 1].tag != Z2D_NODE_MOVE_TO) simple = false;  // something follows the close
        if (!any_seg) slope_dep = true;
      } else {
        const double ex = tag == Z2D_NODE_CURVE_TO ? q[j].p[4] : q[j].p[0], ey = tag == Z2D_NODE_CURVE_TO ? q[j].p[5] : q[j].p[1];
        moving += (ex != cx || ey != cy);
        any_seg = any_seg || ex != cx || ey != cy ||
                  (tag == Z2D_NODE_CURVE_TO && (q[j].p[0] != cx || q[j].p[1] != cy || q[j].p[2] != cx || q[j].p[3] != cy));
        cx = ex;
        cy = ey;
      }
      j++;
    }
    if (!any_seg) slope_dep = true;
    simple = simple && q[j - 1].tag == Z2D_NODE_CLOSE_PATH && moving >= 2;
    if (keep_lone && slope_dep) {
      if (n_sp > 0) {  // sub-paths of one draw are contiguous: extend the previous one over this sub-path
        if (out) {
          out[n_sp - 1].node_end = base + (uint32_t)j;
          out[n_sp - 1].flags = (out[n_sp - 1].flags & ~kSpLastOfDraw) | ((j == m) ? kSpLastOfDraw : 0u);
        }
      } else {
        if (out) out[n_sp] = DevSubPath{draw_index, base + (uint32_t)i, base + (uint32_t)j, (j == m) ? kSpLastOfDraw : 0u};
        n_sp++;
      }
    } else if (j - i > 1) {
      if (out) {
        DevSubPath sp;
        sp.draw = draw_index;
        sp.node_begin = base + (uint32_t)i;
        sp.node_end = base + (uint32_t)j;
        sp.flags = ((j == m) ? kSpLastOfDraw : 0u) | (simple ? kSpNodeParallel : 0u);
        out[n_sp] = sp;
      }
      n_sp++;
      n_par += simple ? 1u : 0u;
      if (all_simple && !simple) *all_simple = false;
    }
    i = j;
  }
  return Z2D_OK;
}

// Split the node list at every move_to and append nodes + sub-path records to the batch.
int record_nodes(z2d_ctx* c, uint32_t draw_index, const z2d_node* nodes, size_t n, bool allow_parallel, bool keep_lone) {
  // leading nodes before the first move_to: line_to / curve_to have no current point
  // (fill_plotter.zig:50,53 -> InternalError.InvalidState); a leading close_path is a no-op.
  size_t first = 0;
  while (first < n && nodes[first].tag != Z2D_NODE_MOVE_TO) {
    if (nodes[first].tag == Z2D_NODE_LINE_TO || nodes[first].tag == Z2D_NODE_CURVE_TO) return Z2D_E_INVALID_STATE;
    if (nodes[first].tag > Z2D_NODE_CLOSE_PATH) return Z2D_E_INVALID_ARG;
    first++;
  }
  const uint32_t base = (uint32_t)c->rec->nodes.n;
  if (!c->rec->nodes.append(nodes + first, n - first)) return Z2D_E_OUT_OF_MEMORY;
  uint32_t n_sp = 0, n_par = 0;
  bool all_simple = true;
  int rc = split_subpaths(nodes + first, n - first, base, draw_index, allow_parallel, nullptr, n_sp, n_par, keep_lone, &all_simple);
  if (rc) return rc;
  allow_parallel = allow_parallel && all_simple;
  if (!c->rec->subpaths.reserve(c->rec->subpaths.n + n_sp)) return Z2D_E_OUT_OF_MEMORY;
  split_subpaths(nodes + first, n - first, base, draw_index, allow_parallel, c->rec->subpaths.p + c->rec->subpaths.n, n_sp, n_par, keep_lone);
  c->rec->subpaths.n += n_sp;
  c->rec->n_par_sp += n_par;
  return Z2D_OK;
}

// Band surfaces (one band per GPU of a very large canvas): every rank is handed the SAME calls, so a fill whose control-point
// hull lies entirely above or below the rows this surface holds is dropped before it costs an upload and a flattening pass.
// Exactly neutral: a curve stays inside the hull of its control points, so the fill has no edge -- hence no coverage region --
// on these rows (only for bounded operators: the others also touch pixels outside the shape).
bool fill_misses_band(const z2d_sfc* s, const z2d_node* nodes, size_t n, uint32_t op) {
  if ((s->y0 == 0 && s->vh == s->h) || !op_is_bounded(op) || n == 0 || nodes[0].tag != Z2D_NODE_MOVE_TO) return false;
  double lo = INFINITY, hi = -INFINITY;
  for (size_t i = 0; i < n; i++) {
    const z2d_node& nd = nodes[i];
    const int np = nd.tag == Z2D_NODE_CURVE_TO ? 3 : (nd.tag == Z2D_NODE_CLOSE_PATH ? 0 : 1);
    for (int k = 0; k < np; k++) {
      const double y = nd.p[2 * k + 1];
      if (!(y == y)) return false;  // NaN: leave it to the pipeline
      lo = y < lo ? y : lo;
      hi = y > hi ? y : hi;
    }
  }
  return hi < (double)s->y0 - 2.0 || lo > (double)(s->y0 + s->h) + 2.0;
}

constexpr size_t kMaxBatchNodes = 8u << 20;   // flush thresholds (bounds pinned + device scratch)
constexpr size_t kMaxBatchDraws = 1u << 20;
constexpr size_t kMinChunkDraws = 8192;        // smallest batch worth handing to the worker (fixed cost of a batch ~0.3 ms)

}  // namespace

// =====================================================================================
extern "C" {

int32_t z2d_version(void) { return 1; }

const char* z2d_last_error(const z2d_ctx* ctx) { return ctx ? ctx->last_error.c_str() : "null context"; }

int32_t z2d_ctx_create(int32_t device, void* stream, z2d_ctx** out) {
  if (!out) return Z2D_E_INVALID_ARG;
  *out = nullptr;
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return Z2D_E_DEVICE;
  z2d_ctx* c = new z2d_ctx();
  c->device = device;
  c->bat[1].index = 1;
  {
    // Recording threads for z2d_submit: from the cores this process may actually run on (its affinity mask -- a launcher that
    // pins every rank to its own slice of the host is honoured), shared with the other ranks of a torchrun launch when nobody
    // pinned anything (LOCAL_WORLD_SIZE processes on the same mask).  Two cores stay free for the caller and the batch worker.
    unsigned avail = std::thread::hardware_concurrency();
    cpu_set_t set;
    bool pinned = false;
    if (sched_getaffinity(0, sizeof set, &set) == 0) {
      const unsigned cnt = (unsigned)CPU_COUNT(&set);
      pinned = cnt && cnt < avail;
      if (cnt) avail = cnt;
    }
    const char* env = getenv("Z2D_RECORD_THREADS");
    const char* lws = getenv("LOCAL_WORLD_SIZE");
    const unsigned ranks = (!pinned && lws && atoi(lws) > 0) ? (unsigned)atoi(lws) : 1u;
    const unsigned mine = avail / ranks;
    c->record_threads = env ? (unsigned)atoi(env) : std::min(4u, mine > 2u ? mine - 2u : 1u);
    if (c->record_threads < 1) c->record_threads = 1;
  }
  if (stream) {
    c->stream = (cudaStream_t)stream;
  } else {
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
      delete c;
      return Z2D_E_DEVICE;
    }
    c->own_stream = true;
  }
  cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
  for (auto& e : c->ev) cudaEventCreate(&e);
  cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&c->d2h_stream, cudaStreamNonBlocking);
  cudaEventCreateWithFlags(&c->ev_d2h, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&c->ev_up, cudaEventDisableTiming);
  for (InputSet& is : c->in) cudaEventCreateWithFlags(&is.done, cudaEventDisableTiming);
  if (getenv("Z2D_NO_RASTER_PRELOAD") == nullptr) raster_preload();
  c->small_enabled = getenv("Z2D_NO_SMALL_BATCH") == nullptr;
  c->stroke_units = getenv("Z2D_NO_STROKE_UNITS") == nullptr;
  c->fill_single_pass = getenv("Z2D_NO_FILL_SINGLE_PASS") == nullptr;
  if (cudaHostAlloc((void**)&c->h_total, 64, cudaHostAllocDefault) != cudaSuccess || cudaHostAlloc((void**)&c->h_small, 32, cudaHostAllocDefault) != cudaSuccess || c->d_blue.ensure(sizeof(z2d_blue_noise_64x64)) != cudaSuccess ||
      cudaMemcpyAsync(c->d_blue.p, z2d_blue_noise_64x64, sizeof(z2d_blue_noise_64x64), cudaMemcpyHostToDevice, c->stream) != cudaSuccess) {
    delete c;
    return Z2D_E_DEVICE;
  }
  *out = c;
  return Z2D_OK;
}

void z2d_ctx_destroy(z2d_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  wait_idle(c);
  if (c->worker.joinable()) {
    {
      std::lock_guard<std::mutex> lk(c->mu);
      c->stop = true;
    }
    c->cv.notify_all();
    c->worker.join();
  }
  cudaStreamSynchronize(c->stream);
  clear_batch(c, c->bat[0]);
  clear_batch(c, c->bat[1]);
  DevBuf* bufs[] = {&c->d_blue, &c->d_draws, &c->d_sp_order, &c->d_sp_keys, &c->d_sp_count, &c->d_sp_off, &c->d_edges, &c->d_edge_draw, &c->d_draw_bands,
                    &c->d_draw_band_off, &c->d_band_count, &c->d_band_off, &c->d_band_cursor, &c->d_band_edges, &c->d_list_cnt,
                    &c->d_list_off, &c->d_list_items, &c->d_scan_tmp, &c->d_comp_grads, &c->d_comp_stop_off, &c->d_comp_stop_col,
                    &c->d_node_sp, &c->d_curve_list};
  for (DevBuf* b : bufs) b->release();
  c->d_export.release();
  c->d_gamma.release();
  c->d_glyph_nodes.release();
  c->d_small_out.release();
  if (c->h_small) cudaFreeHost(c->h_small);
  c->d_sim_rows.release();
  c->d_sim_perm.release();
  c->d_sim_x.release();
  for (InputSet& is : c->in) {
    is.release();
    if (is.done) cudaEventDestroy(is.done);
  }
  if (c->ev_up) cudaEventDestroy(c->ev_up);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  if (c->d2h_stream) {
    cudaStreamSynchronize(c->d2h_stream);
    cudaStreamDestroy(c->d2h_stream);
  }
  if (c->ev_d2h) cudaEventDestroy(c->ev_d2h);
  for (Batch& b : c->bat) {
    b.nodes.release();
    b.subpaths.release();
    b.draws.release();
    b.side.release();
  }
  if (c->h_total) cudaFreeHost(c->h_total);
  c->d_counters.release();
  c->d_boxes.release();
  c->d_band_hdr.release();
  c->d_band_xr.release();
  c->d_hots.release();
  for (auto& e : c->ev) if (e) cudaEventDestroy(e);
  if (c->own_stream) cudaStreamDestroy(c->stream);
  delete c;
}

int32_t z2d_ctx_set_chunk(z2d_ctx* c, uint32_t max_draws) {
  if (!c) return Z2D_E_INVALID_ARG;
  cudaSetDevice(c->device);
  int rc = flush(c);
  c->chunk_draws = max_draws;
  return rc;
}

int32_t z2d_flush(z2d_ctx* c) {
  if (!c) return Z2D_E_INVALID_ARG;
  cudaSetDevice(c->device);
  return flush(c);
}

int32_t z2d_sync(z2d_ctx* c) {
  if (!c) return Z2D_E_INVALID_ARG;
  cudaSetDevice(c->device);
  int rc = flush(c);
  if (rc) return rc;
  CK(c, cudaStreamSynchronize(c->stream));
  if (c->d2h_pending) {
    CK(c, cudaStreamSynchronize(c->d2h_stream));
    c->d2h_pending = false;
  }
  return Z2D_OK;
}

int32_t z2d_get_stats(const z2d_ctx* cc, z2d_stats* out) {
  if (!cc || !out) return Z2D_E_INVALID_ARG;
  z2d_ctx* c = const_cast<z2d_ctx*>(cc);
  wait_idle(c);
  {
    cudaSetDevice(c->device);
    const int rc0 = check_small(c);
    if (rc0 != Z2D_OK) return rc0;
  }
  if (c->stats_pending) {  // device counters and stage timings of the last batch
    cudaSetDevice(c->device);
    CK(c, cudaStreamSynchronize(c->stream));
    unsigned long long h[4] = {0, 0, 0, 0};
    CK(c, cudaMemcpy(h, c->d_counters.p, 32, cudaMemcpyDeviceToHost));
    c->stats.covered_px = h[0];
    c->stats.region_px = h[1];
    c->stats.tile_pairs = h[2];
    c->stats.crossings = h[3];
    cudaEventElapsedTime(&c->stats.ms_flatten, c->ev[0], c->ev[1]);
    cudaEventElapsedTime(&c->stats.ms_bin, c->ev[1], c->ev[2]);
    cudaEventElapsedTime(&c->stats.ms_lists, c->ev[2], c->ev[3]);
    cudaEventElapsedTime(&c->stats.ms_raster, c->ev[3], c->ev[4]);
    cudaEventElapsedTime(&c->stats.ms_total, c->ev[0], c->ev[4]);
    c->stats_pending = false;
  }
  *out = c->stats;
  return Z2D_OK;
}

int32_t z2d_replay(z2d_ctx* c) {
  if (!c) return Z2D_E_INVALID_ARG;
  cudaSetDevice(c->device);
  int rc = flush(c);
  if (rc) return rc;
  rc = run_pipeline(c, true);
  if (c->last.valid) cudaEventRecord(c->in[c->last.set].done, c->stream);  // the next upload into this input set waits for the replay
  return rc;
}

// ------------------------------------------------------------------------- surfaces
int32_t z2d_surface_create(z2d_ctx* c, uint32_t format, int32_t width, int32_t height, const z2d_pixel* initial_px, z2d_sfc** out) {
  if (!c || !out || format > Z2D_FMT_ALPHA1) return Z2D_E_INVALID_ARG;
  *out = nullptr;
  if (width < 1) return Z2D_E_INVALID_WIDTH;    // surface.zig:389,628
  if (height < 1) return Z2D_E_INVALID_HEIGHT;
  cudaSetDevice(c->device);
  z2d_sfc* s = new z2d_sfc();
  s->ctx = c;
  s->fmt = format;
  s->w = width;
  s->h = height;
  s->vh = height;
  s->bytes = ((size_t)width * (size_t)height * (size_t)fmt_bits(format) + 7) / 8;
  s->slot[0] = s->slot[1] = -1;
  const size_t alloc = (s->bytes + 31) & ~(size_t)15;  // word-granular atomics on packed formats may touch the padding
  cudaError_t e = cudaMalloc((void**)&s->data, alloc);
  if (e != cudaSuccess) {
    delete s;
    c->last_error = cudaGetErrorString(e);
    return e == cudaErrorMemoryAllocation ? Z2D_E_OUT_OF_MEMORY : Z2D_E_DEVICE;
  }
  e = cudaMemsetAsync(s->data, 0, alloc, c->stream);
  if (e == cudaSuccess && initial_px) {
    const uint32_t raw = pixel_to_raw(format, initial_px->format, initial_px->r, initial_px->g, initial_px->b, initial_px->a);
    if (raw) launch_paint(s->data, format, (size_t)width * (size_t)height, raw, c->stream);
    e = cudaGetLastError();
  }
  if (e != cudaSuccess) {
    cudaFree(s->data);
    delete s;
    return fail(c, "surface init", e);
  }
  *out = s;
  return Z2D_OK;
}

int32_t z2d_surface_create_band(z2d_ctx* c, uint32_t format, int32_t width, int32_t canvas_height, int32_t band_y0, int32_t band_rows,
                                const z2d_pixel* initial_px, z2d_sfc** out) {
  if (!out) return Z2D_E_INVALID_ARG;
  *out = nullptr;
  if (canvas_height < 1) return Z2D_E_INVALID_HEIGHT;
  if (band_y0 < 0 || band_rows < 1 || (band_y0 & (kTile - 1)) != 0 || band_y0 + band_rows > canvas_height) return Z2D_E_INVALID_ARG;
  int rc = z2d_surface_create(c, format, width, band_rows, initial_px, out);
  if (rc) return rc;
  (*out)->y0 = band_y0;
  (*out)->vh = canvas_height;
  return Z2D_OK;
}

int32_t z2d_surface_band(const z2d_sfc* s, int32_t* band_y0, int32_t* canvas_height) {
  if (!s) return Z2D_E_INVALID_ARG;
  if (band_y0) *band_y0 = s->y0;
  if (canvas_height) *canvas_height = s->vh;
  return Z2D_OK;
}

void z2d_surface_destroy(z2d_sfc* s) {
  if (!s) return;
  z2d_ctx* c = s->ctx;
  cudaSetDevice(c->device);
  flush(c);
  cudaStreamSynchronize(c->stream);
  if (c->d2h_pending) {
    cudaStreamSynchronize(c->d2h_stream);
    c->d2h_pending = false;
  }
  if (s->ipc_base) cudaIpcCloseMemHandle(s->ipc_base);
  else if (!s->external) cudaFree(s->data);
  delete s;
}

// ---- band views: the band's rows live INSIDE a full canvas, so finishing a band is finishing its part of the canvas.  With
// the canvas on another GPU of the node (IPC handle, NVLink peer mapping) the raster kernel's tile loads and its one
// write-back per touched tile go straight over NVLink into rank 0's canvas: the gather of SURVEY 8e is fused into K4's
// write-back, there is no separate copy or collective to wait for.
static int32_t make_band_view(z2d_ctx* c, uint8_t* base, void* ipc_base, uint32_t format, int32_t width, int32_t canvas_height, int32_t band_y0,
                              int32_t band_rows, z2d_sfc** out) {
  if (!c || !out || format > Z2D_FMT_ALPHA1) return Z2D_E_INVALID_ARG;
  *out = nullptr;
  if (width < 1) return Z2D_E_INVALID_WIDTH;
  if (canvas_height < 1) return Z2D_E_INVALID_HEIGHT;
  if (band_y0 < 0 || band_rows < 1 || (band_y0 & (kTile - 1)) != 0 || band_y0 + band_rows > canvas_height) return Z2D_E_INVALID_ARG;
  const size_t bit0 = (size_t)band_y0 * (size_t)width * (size_t)fmt_bits(format);
  const size_t bits = (size_t)band_rows * (size_t)width * (size_t)fmt_bits(format);
  // packed formats: a band must start and end on a byte (16 bytes for the vector paths) so that no byte is shared with a neighbour band
  if ((bit0 & 127) != 0 || ((bits & 7) != 0 && band_y0 + band_rows != canvas_height)) return Z2D_E_INVALID_ARG;
  z2d_sfc* s = new z2d_sfc();
  s->ctx = c;
  s->fmt = format;
  s->w = width;
  s->h = band_rows;
  s->y0 = band_y0;
  s->vh = canvas_height;
  s->bytes = (bits + 7) / 8;
  s->slot[0] = s->slot[1] = -1;
  s->data = base + bit0 / 8;
  s->external = true;
  s->ipc_base = ipc_base;
  *out = s;
  return Z2D_OK;
}

int32_t z2d_surface_band_view(z2d_sfc* canvas, int32_t band_y0, int32_t band_rows, z2d_sfc** out) {
  if (!canvas || canvas->external || canvas->y0 != 0 || canvas->vh != canvas->h) return Z2D_E_INVALID_ARG;
  return make_band_view(canvas->ctx, canvas->data, nullptr, canvas->fmt, canvas->w, canvas->h, band_y0, band_rows, out);
}

int32_t z2d_surface_ipc_export(z2d_sfc* s, void* handle64) {
  if (!s || !handle64 || s->external) return Z2D_E_INVALID_ARG;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  z2d_ctx* c = s->ctx;
  cudaSetDevice(c->device);
  int rc = flush(c);
  if (rc) return rc;
  CK(c, cudaStreamSynchronize(c->stream));  // the memset / paint of the creation must have landed before a peer writes
  CK(c, cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle64), s->data));
  return Z2D_OK;
}

int32_t z2d_surface_open_peer_band(z2d_ctx* c, const void* handle64, uint32_t format, int32_t width, int32_t canvas_height, int32_t band_y0,
                                   int32_t band_rows, z2d_sfc** out) {
  if (!c || !handle64 || !out) return Z2D_E_INVALID_ARG;
  *out = nullptr;
  cudaSetDevice(c->device);
  void* base = nullptr;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof h);
  CK(c, cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
  int32_t rc = make_band_view(c, (uint8_t*)base, base, format, width, canvas_height, band_y0, band_rows, out);
  if (rc != Z2D_OK) cudaIpcCloseMemHandle(base);
  return rc;
}

size_t z2d_surface_byte_len(const z2d_sfc* s) { return s ? s->bytes : 0; }
int32_t z2d_surface_width(const z2d_sfc* s) { return s ? s->w : 0; }
int32_t z2d_surface_height(const z2d_sfc* s) { return s ? s->h : 0; }
uint32_t z2d_surface_format(const z2d_sfc* s) { return s ? s->fmt : 0; }
void* z2d_surface_device_ptr(z2d_sfc* s) { return s ? s->data : nullptr; }

int32_t z2d_surface_upload(z2d_sfc* s, const void* host, size_t n) {
  if (!s || !host || n != s->bytes) return Z2D_E_INVALID_ARG;
  z2d_ctx* c = s->ctx;
  cudaSetDevice(c->device);
  int rc = flush(c);
  if (rc) return rc;
  CK(c, cudaMemcpyAsync(s->data, host, n, cudaMemcpyHostToDevice, c->stream));
  return Z2D_OK;
}

int32_t z2d_surface_download(z2d_sfc* s, void* host, size_t n) {
  if (!s || !host || n != s->bytes) return Z2D_E_INVALID_ARG;
  z2d_ctx* c = s->ctx;
  cudaSetDevice(c->device);
  int rc = flush(c);
  if (rc) return rc;
  CK(c, cudaMemcpyAsync(host, s->data, n, cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  return Z2D_OK;
}

int32_t z2d_surface_download_async(z2d_sfc* s, void* host, size_t n) {
  if (!s || !host || n != s->bytes) return Z2D_E_INVALID_ARG;
  z2d_ctx* c = s->ctx;
  cudaSetDevice(c->device);
  int rc = flush(c);  // everything recorded so far is on the stream; the copy waits for it and nothing later waits for the copy
  if (rc) return rc;
  CK(c, cudaEventRecord(c->ev_d2h, c->stream));
  CK(c, cudaStreamWaitEvent(c->d2h_stream, c->ev_d2h, 0));
  CK(c, cudaMemcpyAsync(host, s->data, n, cudaMemcpyDeviceToHost, c->d2h_stream));
  c->d2h_pending = true;
  return Z2D_OK;
}

// export_png.zig:150-373 without the zlib/chunk framing: the scanline bytes, produced on the device
static size_t export_row_bytes(const z2d_sfc* s, uint32_t flags) {
  size_t rb;
  switch (s->fmt) {
    case Z2D_FMT_ARGB:
    case Z2D_FMT_RGBA: rb = (size_t)s->w * 4; break;
    case Z2D_FMT_XRGB:
    case Z2D_FMT_RGB: rb = (size_t)s->w * 3; break;
    default: rb = ((size_t)s->w * (size_t)fmt_bits(s->fmt) + 7) / 8;
  }
  return rb + ((flags & Z2D_EXPORT_FILTER_BYTE) ? 1 : 0);
}

size_t z2d_surface_export_size(const z2d_sfc* s, uint32_t flags) { return s ? export_row_bytes(s, flags) * (size_t)s->h : 0; }

int32_t z2d_surface_export(z2d_sfc* s, uint32_t flags, void* host, size_t n) {
  if (!s || !host || (flags & ~(Z2D_EXPORT_SRGB | Z2D_EXPORT_FILTER_BYTE))) return Z2D_E_INVALID_ARG;
  const size_t rb = export_row_bytes(s, flags);
  if (n != rb * (size_t)s->h) return Z2D_E_INVALID_ARG;
  z2d_ctx* c = s->ctx;
  cudaSetDevice(c->device);
  int rc = flush(c);
  if (rc) return rc;
  if (n == 0) return Z2D_OK;
  CK(c, c->d_export.ensure(n));
  const bool colour = s->fmt <= Z2D_FMT_RGBA;
  if ((flags & Z2D_EXPORT_SRGB) && colour && !c->d_gamma.p) {
    // color_vector.zig:210-224, 266-284: round(255 * pow(c / 255, 1 / 2.2)) in f32, one entry per 8-bit channel value
    uint8_t lut[256];
    const float inv_gamma = 1.0f / 2.2f;
    for (int i = 0; i < 256; ++i) lut[i] = (uint8_t)roundf(255.0f * powf((float)i / 255.0f, inv_gamma));
    CK(c, c->d_gamma.ensure(256));
    CK(c, cudaMemcpyAsync(c->d_gamma.p, lut, 256, cudaMemcpyHostToDevice, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));  // lut is on this stack frame
  }
  ExportArgs A{};
  A.data = s->data;
  A.fmt = s->fmt;
  A.w = s->w;
  A.h = s->h;
  A.filter_byte = (flags & Z2D_EXPORT_FILTER_BYTE) ? 1u : 0u;
  A.row_bytes = rb;
  A.items_per_row = s->fmt <= Z2D_FMT_ALPHA8 ? (uint32_t)s->w : (uint32_t)(rb - A.filter_byte);
  A.gamma = ((flags & Z2D_EXPORT_SRGB) && colour) ? c->d_gamma.as<uint8_t>() : nullptr;
  A.out = c->d_export.as<uint8_t>();
  launch_export(A, c->sm_count, c->stream);
  CK(c, cudaGetLastError());
  CK(c, cudaMemcpyAsync(host, A.out, n, cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  return Z2D_OK;
}

int32_t z2d_surface_paint_pixel(z2d_sfc* s, const z2d_pixel* px) {
  if (!s || !px || px->format > Z2D_FMT_ALPHA1) return Z2D_E_INVALID_ARG;
  z2d_ctx* c = s->ctx;
  cudaSetDevice(c->device);
  int rc = flush(c);
  if (rc) return rc;
  const uint32_t raw = pixel_to_raw(s->fmt, px->format, px->r, px->g, px->b, px->a);
  launch_paint(s->data, s->fmt, (size_t)s->w * (size_t)s->h, raw, c->stream);
  CK(c, cudaGetLastError());
  return Z2D_OK;
}

int32_t z2d_surface_downsample(z2d_sfc* s) {
  if (!s || s->external || s->y0 != 0 || s->vh != s->h) return Z2D_E_INVALID_ARG;  // whole surfaces that own their pixels
  if (s->w < 4 || s->h < 4) return Z2D_OK;  // surface.zig:448: nothing happens
  z2d_ctx* c = s->ctx;
  cudaSetDevice(c->device);
  int rc = flush(c);
  if (rc) return rc;
  const int32_t w = s->w / 4, h = s->h / 4;
  const size_t bytes = ((size_t)w * (size_t)h * (size_t)fmt_bits(s->fmt) + 7) / 8, alloc = (bytes + 31) & ~(size_t)15;
  uint8_t* nd = nullptr;
  cudaError_t e = cudaMalloc((void**)&nd, alloc);
  if (e != cudaSuccess) return e == cudaErrorMemoryAllocation ? Z2D_E_OUT_OF_MEMORY : fail(c, "downsample", e);
  CK(c, cudaMemsetAsync(nd, 0, alloc, c->stream));
  launch_downsample(s->data, nd, s->fmt, s->w, w, h, c->stream);
  CK(c, cudaGetLastError());
  CK(c, cudaStreamSynchronize(c->stream));  // the old buffer is released: the reference resizes in place (surface.zig:471-490)
  if (c->d2h_pending) {
    CK(c, cudaStreamSynchronize(c->d2h_stream));
    c->d2h_pending = false;
  }
  cudaFree(s->data);
  s->data = nd;
  s->w = w;
  s->h = h;
  s->vh = h;
  s->bytes = bytes;
  return Z2D_OK;
}

int32_t z2d_surface_put_pixel(z2d_sfc* s, int32_t x, int32_t y, const z2d_pixel* px) {
  if (!s || !px || px->format > Z2D_FMT_ALPHA1) return Z2D_E_INVALID_ARG;
  if (x < 0 || y < 0 || x >= s->w || y >= s->vh) return Z2D_OK;  // surface.zig:520,770
  y -= s->y0;  // band surface: canvas row -> row held here
  if (y < 0 || y >= s->h) return Z2D_OK;
  z2d_ctx* c = s->ctx;
  cudaSetDevice(c->device);
  int rc = flush(c);
  if (rc) return rc;
  const uint32_t raw = pixel_to_raw(s->fmt, px->format, px->r, px->g, px->b, px->a);
  launch_put_pixel(s->data, s->fmt, (size_t)s->w * (size_t)y + (size_t)x, raw, c->stream);
  CK(c, cudaGetLastError());
  return Z2D_OK;
}

int32_t z2d_surface_get_pixel(z2d_sfc* s, int32_t x, int32_t y, z2d_pixel* out) {
  if (!s || !out) return Z2D_E_INVALID_ARG;
  if (x < 0 || y < 0 || x >= s->w || y >= s->vh) return 1;  // surface.zig:280-286: null
  y -= s->y0;
  if (y < 0 || y >= s->h) return 1;
  z2d_ctx* c = s->ctx;
  cudaSetDevice(c->device);
  int rc = flush(c);
  if (rc) return rc;
  const size_t idx = (size_t)s->w * (size_t)y + (size_t)x;
  const int bits = fmt_bits(s->fmt);
  uint32_t word = 0;
  const size_t byte = bits == 32 ? idx * 4 : idx * (size_t)bits / 8;
  CK(c, cudaMemcpyAsync(&word, s->data + byte, bits == 32 ? 4 : 1, cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  z2d_pixel px{};
  px.format = s->fmt;
  if (bits == 32) {
    const RGBA16 v = raw_to_rgba16(s->fmt, word);
    px.r = (uint8_t)v.r;
    px.g = (uint8_t)v.g;
    px.b = (uint8_t)v.b;
    px.a = (uint8_t)v.a;
  } else {  // alpha formats: the stored sample, LSB-first within the byte (surface.zig:880-887)
    px.a = (uint8_t)((word >> ((idx * (size_t)bits) & 7)) & ((1u << bits) - 1u));
  }
  *out = px;
  return Z2D_OK;
}

// ------------------------------------------------------------------------- painter
static int32_t record_draw(z2d_ctx* c, z2d_sfc* s, const z2d_pattern* pattern, const z2d_node* nodes, size_t n, DevDraw& d) {
  const size_t save_g = c->rec->grads.size(), save_o = c->rec->stop_offsets.size(), save_c = c->rec->stop_colors.size();
  int rc = pattern_to_src(*pattern, d.src, c->rec->grads, c->rec->stop_offsets, c->rec->stop_colors);
  if (rc) return rc;
  d.surface = batch_slot(c, s);
  d.reduces = (d.src.kind == Z2D_PARAM_PIXEL && (d.op == Z2D_OP_SRC || (d.op == Z2D_OP_SRC_OVER && px_is_opaque(pattern->pixel)))) ? 1u : 0u;
  d.paint_raw = d.src.kind == Z2D_PARAM_PIXEL
                    ? pixel_to_raw(s->fmt, pattern->pixel.format, pattern->pixel.r, pattern->pixel.g, pattern->pixel.b, pattern->pixel.a)
                    : 0u;
  const size_t save_nodes = c->rec->nodes.n, save_sp = c->rec->subpaths.n, save_st = c->rec->strokes.size(), save_src = c->rec->srcs.size();
  const uint32_t di = (uint32_t)c->rec->draws.n;
  const uint32_t save_par = c->rec->n_par_sp;
  rc = record_nodes(c, di, nodes, n, d.kind == 0 && d.mode == 0, d.kind == 1 && d.dash_count > 0 && d.cap != Z2D_CAP_BUTT);
  const uint32_t save_unit = c->rec->n_unit_sp;
  if (rc == Z2D_OK && d.kind == 1 && d.mode == 0 && c->stroke_units) {
    // ordinary strokes (not hairlines; not the direct rasteriser with an unbounded operator, whose result depends on the order
    // of the call's edges): tessellated per join / cap by the unit stroker
    for (size_t k = save_sp; k < c->rec->subpaths.n; k++) c->rec->subpaths.p[k].flags |= kSpStrokeUnits;
    c->rec->n_unit_sp += (uint32_t)(c->rec->subpaths.n - save_sp);
  }
  DrawIn in;
  in.surface = d.surface;
  in.opts = pack_draw_opts(d.kind, d.aa, d.rule, d.op, d.precision, d.reduces, d.mode);
  in.paint_raw = d.paint_raw;
  in.px_rgba = d.src.px_rgba;
  in.tolerance = d.tolerance;
  in.src_index = kNoIndex;
  in.stroke_index = kNoIndex;
  if (rc == Z2D_OK && d.src.kind != Z2D_PARAM_PIXEL) {
    in.src_index = (uint32_t)c->rec->srcs.size();
    c->rec->srcs.push_back(d.src);
  }
  if (rc == Z2D_OK && d.kind == 1) {
    StrokeIn si;
    memset(&si, 0, sizeof si);
    si.cap = d.cap; si.join = d.join;
    si.dash_begin = d.dash_begin; si.dash_count = d.dash_count;
    si.pen_begin = d.pen_begin; si.pen_count = d.pen_count;
    si.hair_aa = d.hair_aa; si.hair_tolerance = d.hair_tolerance;
    si.thickness = d.thickness; si.miter_limit = d.miter_limit; si.dash_offset = d.dash_offset;
    for (int k = 0; k < 6; k++) { si.ctm[k] = d.ctm[k]; si.inv[k] = d.inv[k]; }
    in.stroke_index = (uint32_t)c->rec->strokes.size();
    c->rec->strokes.push_back(si);
  }
  if (d.mode == 1) {
    c->rec->iso_mode = d.mode;
    c->rec->iso_node_begin = (uint32_t)save_nodes;
    c->rec->iso_node_end = (uint32_t)c->rec->nodes.n;
  }
  if (rc == Z2D_OK && !c->rec->draws.push(in)) rc = Z2D_E_OUT_OF_MEMORY;
  if (rc) {  // roll back: the failed call draws nothing
    c->rec->nodes.n = save_nodes;
    c->rec->subpaths.n = save_sp;
    c->rec->n_par_sp = save_par;
    c->rec->n_unit_sp = save_unit;
    c->rec->strokes.resize(save_st);
    c->rec->srcs.resize(save_src);
    c->rec->grads.resize(save_g);
    c->rec->stop_offsets.resize(save_o);
    c->rec->stop_colors.resize(save_c);
    return rc;
  }
  if (d.mode == 1 || c->rec->nodes.n > kMaxBatchNodes || c->rec->draws.n > kMaxBatchDraws) return flush(c);
  // Pipelined recording: hand the batch to the worker as soon as it is idle (so the device starts early and is fed batches as
  // large as the time it took to execute the previous one), at the latest every chunk_draws draws.  Asynchronous: recording
  // continues into the other batch.
  if (c->chunk_draws && c->rec->draws.n >= kMinChunkDraws &&
      (c->rec->draws.n >= c->chunk_draws || !c->busy.load(std::memory_order_relaxed)))
    return kick(c);
  return Z2D_OK;
}

int32_t z2d_glyph_cache_add(z2d_ctx* c, const z2d_node* nodes, size_t n, uint32_t* out) {
  if (!c || !out || (n && !nodes)) return Z2D_E_INVALID_ARG;
  if (n == 0) {  // a glyph without outline (whitespace): an entry that expands to nothing
    c->glyphs.push_back({(uint32_t)c->glyph_nodes.size(), 0u, (uint32_t)c->glyph_sps.size(), 0u});
    *out = (uint32_t)c->glyphs.size() - 1;
    return Z2D_OK;
  }
  if (!is_closed_node_set(nodes, n)) return Z2D_E_PATH_NOT_CLOSED;  // painter.zig:82 on the run this glyph would be part of
  if (nodes[0].tag != Z2D_NODE_MOVE_TO) return Z2D_E_INVALID_STATE;  // (outlines start with a move_to: fill_plotter.zig:50,53)
  uint32_t n_sp = 0, n_par = 0;
  int rc = split_subpaths(nodes, n, 0, 0, false, nullptr, n_sp, n_par);
  if (rc) return rc;
  const size_t sp0 = c->glyph_sps.size();
  c->glyph_sps.resize(sp0 + n_sp);
  split_subpaths(nodes, n, 0, 0, false, c->glyph_sps.data() + sp0, n_sp, n_par);
  c->glyphs.push_back({(uint32_t)c->glyph_nodes.size(), (uint32_t)n, (uint32_t)sp0, n_sp});
  c->glyph_nodes.insert(c->glyph_nodes.end(), nodes, nodes + n);
  *out = (uint32_t)c->glyphs.size() - 1;
  return Z2D_OK;
}

int32_t z2d_fill_glyphs(z2d_ctx* c, z2d_sfc* s, const z2d_pattern* pattern, const z2d_glyph_instance* gi, size_t n, const z2d_fill_opts* o) {
  if (!c || !s || !pattern || !o || s->ctx != c || (n && !gi)) return Z2D_E_INVALID_ARG;
  if (o->op >= Z2D_OP_COUNT || o->fill_rule > 1 || o->precision > 1 || o->anti_aliasing_mode > Z2D_AA_SUPERSAMPLE_4X) return Z2D_E_INVALID_ARG;
  if (pattern->kind == Z2D_PATTERN_OPAQUE && !px_can_demultiply(pattern->pixel)) return Z2D_E_PIXEL_SOURCE_NOT_PREMULTIPLIED;
  size_t total = 0, total_sp = 0;
  for (size_t i = 0; i < n; i++) {
    if (gi[i].glyph >= c->glyphs.size()) return Z2D_E_INVALID_ARG;
    total += c->glyphs[gi[i].glyph].n_nodes;
    total_sp += c->glyphs[gi[i].glyph].n_sp;
  }
  if (total == 0) return Z2D_OK;  // painter.zig:81: an empty node list is a no-op
  DevDraw d;
  memset(&d, 0, sizeof d);
  uint32_t aa = (s->fmt == Z2D_FMT_ALPHA1) ? (uint32_t)Z2D_AA_NONE : o->anti_aliasing_mode;
  if (aa == Z2D_AA_DEFAULT) aa = Z2D_AA_MULTISAMPLE_4X;
  d.aa = aa;
  d.rule = o->fill_rule;
  d.op = o->op;
  d.precision = op_requires_float(o->op) ? (uint32_t)Z2D_PRECISION_FLOAT : o->precision;
  d.tolerance = o->tolerance > 0.001 ? o->tolerance : 0.001;
  if (aa == Z2D_AA_NONE && !op_is_bounded(o->op)) d.mode = 2;
  Batch& B = *c->rec;
  const size_t save_g = B.grads.size(), save_o = B.stop_offsets.size(), save_c = B.stop_colors.size();
  int rc = pattern_to_src(*pattern, d.src, B.grads, B.stop_offsets, B.stop_colors);
  if (rc) return rc;
  const uint32_t surface = batch_slot(c, s);
  const uint32_t reduces = (d.src.kind == Z2D_PARAM_PIXEL && (d.op == Z2D_OP_SRC || (d.op == Z2D_OP_SRC_OVER && px_is_opaque(pattern->pixel)))) ? 1u : 0u;
  const uint32_t paint_raw = d.src.kind == Z2D_PARAM_PIXEL
                                 ? pixel_to_raw(s->fmt, pattern->pixel.format, pattern->pixel.r, pattern->pixel.g, pattern->pixel.b, pattern->pixel.a)
                                 : 0u;
  const size_t node0 = B.nodes.n, sp0 = B.subpaths.n;
  if (!B.nodes.reserve(node0 + total) || !B.subpaths.reserve(sp0 + total_sp)) {
    B.grads.resize(save_g); B.stop_offsets.resize(save_o); B.stop_colors.resize(save_c);
    return Z2D_E_OUT_OF_MEMORY;
  }
  memset(B.nodes.p + node0, 0, total * sizeof(z2d_node));  // (overwritten on the device by k_expand_glyphs)
  const uint32_t di = (uint32_t)B.draws.n;
  size_t at = node0, sp_at = sp0;
  for (size_t i = 0; i < n; i++) {
    const z2d_ctx::GlyphEntry& g = c->glyphs[gi[i].glyph];
    if (!g.n_nodes) continue;
    GlyphInst gx;
    gx.src = g.node_off; gx.n_nodes = g.n_nodes; gx.dst = (uint32_t)at; gx._pad = 0;
    for (int k = 0; k < 6; k++) gx.m[k] = gi[i].m[k];
    B.ginst.push_back(gx);
    for (uint32_t k = 0; k < g.n_sp; k++) {  // every contour is plotted by the sequential fill plotter (exact for any transformation)
      DevSubPath sp = c->glyph_sps[g.sp_off + k];
      sp.draw = di;
      sp.node_begin += (uint32_t)at;
      sp.node_end += (uint32_t)at;
      sp.flags = 0;
      B.subpaths.p[sp_at++] = sp;
    }
    at += g.n_nodes;
  }
  if (sp_at > sp0) B.subpaths.p[sp_at - 1].flags |= kSpLastOfDraw;
  B.nodes.n = at;
  B.subpaths.n = sp_at;
  DrawIn in;
  in.surface = surface;
  in.opts = pack_draw_opts(0, d.aa, d.rule, d.op, d.precision, reduces, d.mode);
  in.paint_raw = paint_raw;
  in.px_rgba = d.src.px_rgba;
  in.tolerance = d.tolerance;
  in.src_index = kNoIndex;
  in.stroke_index = kNoIndex;
  if (d.src.kind != Z2D_PARAM_PIXEL) {
    in.src_index = (uint32_t)B.srcs.size();
    B.srcs.push_back(d.src);
  }
  if (!B.draws.push(in)) return Z2D_E_OUT_OF_MEMORY;
  if (B.nodes.n > kMaxBatchNodes || B.draws.n > kMaxBatchDraws) return flush(c);
  return Z2D_OK;
}

int32_t z2d_fill(z2d_ctx* c, z2d_sfc* s, const z2d_pattern* pattern, const z2d_node* nodes, size_t n, const z2d_fill_opts* o) {
  if (!c || !s || !pattern || !o || s->ctx != c || (n && !nodes)) return Z2D_E_INVALID_ARG;
  if (o->op >= Z2D_OP_COUNT || o->fill_rule > 1 || o->precision > 1 || o->anti_aliasing_mode > Z2D_AA_SUPERSAMPLE_4X) return Z2D_E_INVALID_ARG;
  if (pattern->kind == Z2D_PATTERN_OPAQUE && !px_can_demultiply(pattern->pixel)) return Z2D_E_PIXEL_SOURCE_NOT_PREMULTIPLIED;  // painter.zig:73-79
  if (n == 0) return Z2D_OK;                                         // painter.zig:81
  if (!is_closed_node_set(nodes, n)) return Z2D_E_PATH_NOT_CLOSED;   // painter.zig:82
  if (fill_misses_band(s, nodes, n, o->op)) return Z2D_OK;
  DevDraw d;
  memset(&d, 0, sizeof d);
  d.kind = 0;
  // painter.zig:88-97: 1-bit alpha surfaces are never anti-aliased; .default is MSAA
  uint32_t aa = (s->fmt == Z2D_FMT_ALPHA1) ? (uint32_t)Z2D_AA_NONE : o->anti_aliasing_mode;
  if (aa == Z2D_AA_DEFAULT) aa = Z2D_AA_MULTISAMPLE_4X;
  d.aa = aa;
  d.rule = o->fill_rule;
  d.op = o->op;
  d.precision = op_requires_float(o->op) ? (uint32_t)Z2D_PRECISION_FLOAT : o->precision;  // multisample.zig:36, compositor.zig:317-322
  d.scale = aa == Z2D_AA_NONE ? 1.0 : 4.0;
  d.tolerance = o->tolerance > 0.001 ? o->tolerance : 0.001;  // painter.zig:103
  if (aa == Z2D_AA_NONE && !op_is_bounded(o->op)) d.mode = 2;  // direct.zig clears per span pair: composited from row records
  return record_draw(c, s, pattern, nodes, n, d);
}

int32_t z2d_stroke(z2d_ctx* c, z2d_sfc* s, const z2d_pattern* pattern, const z2d_node* nodes, size_t n, const z2d_stroke_opts* o) {
  if (!c || !s || !pattern || !o || s->ctx != c || (n && !nodes)) return Z2D_E_INVALID_ARG;
  if (pattern->kind == Z2D_PATTERN_OPAQUE && !px_can_demultiply(pattern->pixel)) return Z2D_E_PIXEL_SOURCE_NOT_PREMULTIPLIED;
  {  // painter.zig:231: the CTM must be invertible (checked before the empty-node early-out)
    const double ax = o->ctm[0], by = o->ctm[1], cx = o->ctm[2], dy = o->ctm[3];
    if (by == 0 && cx == 0) {
      if (ax == 0 || dy == 0) return Z2D_E_INVALID_MATRIX;
    } else if (ax * dy - by * cx == 0) {
      return Z2D_E_INVALID_MATRIX;
    }
  }
  if (n == 0) return Z2D_OK;
  if (o->op >= Z2D_OP_COUNT || o->precision > 1 || o->anti_aliasing_mode > Z2D_AA_SUPERSAMPLE_4X || o->line_cap_mode > Z2D_CAP_SQUARE ||
      o->line_join_mode > Z2D_JOIN_BEVEL || (o->n_dashes && !o->dashes))
    return Z2D_E_INVALID_ARG;
  // Dasher.validate (tess/Dasher.zig:15-29): all >= 0 and at least one > 0, else the stroke is not dashed
  bool dashed = false;
  for (size_t i = 0; i < o->n_dashes; i++) {
    if (o->dashes[i] < 0) {
      dashed = false;
      break;
    }
    if (o->dashes[i] > 0) dashed = true;
  }
  if (o->hairline) {  // painter.zig:246-265: polyline plotter + hairline rasteriser, AA mode and tolerance as given
    bool have_pt = false;
    for (size_t i = 0; i < n; i++) {  // polyline_plotter.zig:62-85: line_to / curve_to before any move_to
      if (nodes[i].tag == Z2D_NODE_MOVE_TO) have_pt = true;
      else if (nodes[i].tag > Z2D_NODE_CLOSE_PATH) return Z2D_E_INVALID_ARG;
      else if (nodes[i].tag != Z2D_NODE_CLOSE_PATH && !have_pt) return Z2D_E_INVALID_STATE;
    }
    DevDraw d;
    memset(&d, 0, sizeof d);
    d.kind = 1;
    d.mode = 1;
    d.op = o->op;
    d.precision = op_requires_float(o->op) ? (uint32_t)Z2D_PRECISION_FLOAT : o->precision;
    d.hair_aa = o->anti_aliasing_mode;
    d.hair_tolerance = o->tolerance;
    d.dash_offset = o->dash_offset;
    int frc = flush(c);  // isolated draw: everything recorded so far lands first
    if (frc) return frc;
    const size_t save_d = c->rec->dashes.size();
    if (dashed) {
      d.dash_begin = (uint32_t)c->rec->dashes.size();
      d.dash_count = (uint32_t)o->n_dashes;
      c->rec->dashes.insert(c->rec->dashes.end(), o->dashes, o->dashes + o->n_dashes);
    }
    int32_t rc = record_draw(c, s, pattern, nodes, n, d);
    if (rc != Z2D_OK) c->rec->dashes.resize(std::min(save_d, c->rec->dashes.size()));
    return rc;
  }
  // plotter state machine errors (stroke_plotter.zig:113,134; dashed_plotter.zig:128,178,203): a line_to / curve_to
  // (dashed: also close_path) without a current point fails the whole call with InvalidState
  bool has_curve = false;
  {
    bool have_pt = false;
    for (size_t i = 0; i < n; i++) {
      switch (nodes[i].tag) {
        case Z2D_NODE_MOVE_TO: have_pt = true; break;
        case Z2D_NODE_CURVE_TO: has_curve = true;  // fallthrough
        case Z2D_NODE_LINE_TO:
          if (!have_pt) return Z2D_E_INVALID_STATE;
          break;
        case Z2D_NODE_CLOSE_PATH:
          if (dashed && !have_pt) return Z2D_E_INVALID_STATE;
          have_pt = false;
          break;
        default: return Z2D_E_INVALID_ARG;
      }
    }
  }
  DevDraw d;
  memset(&d, 0, sizeof d);
  d.kind = 1;
  uint32_t aa = (s->fmt == Z2D_FMT_ALPHA1) ? (uint32_t)Z2D_AA_NONE : o->anti_aliasing_mode;  // painter.zig:240-243
  if (aa == Z2D_AA_DEFAULT) aa = Z2D_AA_MULTISAMPLE_4X;
  d.aa = aa;
  d.rule = Z2D_FILL_NON_ZERO;  // painter.zig:308-343
  d.op = o->op;
  d.precision = op_requires_float(o->op) ? (uint32_t)Z2D_PRECISION_FLOAT : o->precision;
  d.scale = aa == Z2D_AA_NONE ? 1.0 : 4.0;
  d.tolerance = o->tolerance > 0.001 ? o->tolerance : 0.001;
  if (aa == Z2D_AA_NONE && !op_is_bounded(o->op)) d.mode = 2;
  // painter.zig:287-304: thin lines lose cap / join / miter settings; minimum width 1/256
  const double min_w = 0.00390625;
  d.cap = o->line_width >= 2 ? o->line_cap_mode : (uint32_t)Z2D_CAP_BUTT;
  d.join = o->line_width >= 2 ? o->line_join_mode : (uint32_t)Z2D_JOIN_MITER;
  d.miter_limit = o->line_width >= 2 ? o->miter_limit : 10.0;
  d.thickness = o->line_width >= min_w ? o->line_width : min_w;
  d.dash_offset = o->dash_offset;
  for (int i = 0; i < 6; i++) d.ctm[i] = o->ctm[i];
  invert_ctm(o->ctm, d.inv);
  const size_t save_d = c->rec->dashes.size(), save_p = c->rec->pens.size();
  if (dashed) {
    d.dash_begin = (uint32_t)c->rec->dashes.size();
    d.dash_count = (uint32_t)o->n_dashes;
    c->rec->dashes.insert(c->rec->dashes.end(), o->dashes, o->dashes + o->n_dashes);
  }
  // the pen exists when a round join / cap is requested, or lazily once a curve is stroked (stroke_plotter.zig:56-59,139-144)
  if (d.join == Z2D_JOIN_ROUND || d.cap == Z2D_CAP_ROUND || has_curve) add_pen(c, d);
  int32_t rc = record_draw(c, s, pattern, nodes, n, d);
  if (rc != Z2D_OK && rc != Z2D_E_DEVICE) {
    c->rec->dashes.resize(std::min(save_d, c->rec->dashes.size()));
    if (c->rec->pens.size() > save_p) {  // the failed call built a pen: it goes, and with it the cache entry that names it
      c->rec->pens.resize(save_p);
      c->rec->pen_cache_n = 0;
      c->rec->pen_cache_next = 0;
    }
  }
  return rc;
}

}  // extern "C"

// ---- parallel recording of runs of plain fills (single-pixel source, tile pipeline): recording is memory bound
// (~0.9 KB read + written per call), so several host threads split a run.  Three phases: (1) every call is validated and
// sized in parallel, (2) a serial prefix sum assigns node / sub-path / draw positions (so the batch is exactly what the
// one-by-one loop would have produced), (3) nodes, sub-paths and draw records are written in parallel.

static void plan_fill(z2d_ctx* c, const z2d_draw_cmd& k, FillPlan& pl) {
  memset(&pl, 0, sizeof pl);
  const z2d_sfc* s = k.surface;
  const z2d_fill_opts* o = k.fill;
  const z2d_pattern* pat = k.pattern;
  if (k.kind != 0 || !s || !pat || !o || s->ctx != c || (k.n_nodes && !k.nodes) || pat->kind != Z2D_PATTERN_OPAQUE ||
      pat->pixel.format > Z2D_FMT_ALPHA1 || o->op >= Z2D_OP_COUNT || o->fill_rule > 1 || o->precision > 1 ||
      o->anti_aliasing_mode > Z2D_AA_SUPERSAMPLE_4X)
    return;  // not eligible: the one-by-one path produces the status
  const uint32_t aa = (s->fmt == Z2D_FMT_ALPHA1) ? (uint32_t)Z2D_AA_NONE : o->anti_aliasing_mode;
  if (aa == Z2D_AA_NONE && !op_is_bounded(o->op)) return;  // row-record mode: recorded by z2d_fill
  pl.eligible = 1;
  if (!px_can_demultiply(pat->pixel)) {
    pl.status = Z2D_E_PIXEL_SOURCE_NOT_PREMULTIPLIED;
    return;
  }
  if (k.n_nodes == 0) {
    pl.skip = 1;
    return;
  }
  if (!is_closed_node_set(k.nodes, k.n_nodes)) {
    pl.status = Z2D_E_PATH_NOT_CLOSED;
    return;
  }
  if (fill_misses_band(s, k.nodes, k.n_nodes, o->op)) {
    pl.skip = 1;
    return;
  }
  size_t first = 0;
  while (first < k.n_nodes && k.nodes[first].tag != Z2D_NODE_MOVE_TO) {
    if (k.nodes[first].tag == Z2D_NODE_LINE_TO || k.nodes[first].tag == Z2D_NODE_CURVE_TO) {
      pl.status = Z2D_E_INVALID_STATE;
      return;
    }
    if (k.nodes[first].tag > Z2D_NODE_CLOSE_PATH) {
      pl.status = Z2D_E_INVALID_ARG;
      return;
    }
    first++;
  }
  pl.first = (uint32_t)first;
  pl.m = (uint32_t)(k.n_nodes - first);
  bool all_simple = true;
  pl.status = split_subpaths(k.nodes + first, pl.m, 0, 0, true, nullptr, pl.n_sp, pl.n_par, false, &all_simple);
  pl.all_simple = all_simple ? 1u : 0u;
  if (!all_simple) pl.n_par = 0;
}

static void write_fill(Batch& B, const z2d_draw_cmd& k, const FillPlan& pl) {
  const z2d_sfc* s = k.surface;
  const z2d_fill_opts* o = k.fill;
  const z2d_pixel& px = k.pattern->pixel;
  memcpy(B.nodes.p + pl.node_base, k.nodes + pl.first, (size_t)pl.m * sizeof(z2d_node));
  uint32_t n_sp, n_par;
  split_subpaths(k.nodes + pl.first, pl.m, pl.node_base, pl.draw_idx, pl.all_simple != 0, B.subpaths.p + pl.sp_base, n_sp, n_par);
  uint32_t aa = (s->fmt == Z2D_FMT_ALPHA1) ? (uint32_t)Z2D_AA_NONE : o->anti_aliasing_mode;  // as z2d_fill / record_draw
  if (aa == Z2D_AA_DEFAULT) aa = Z2D_AA_MULTISAMPLE_4X;
  const uint32_t precision = op_requires_float(o->op) ? (uint32_t)Z2D_PRECISION_FLOAT : o->precision;
  const uint32_t reduces = (o->op == Z2D_OP_SRC || (o->op == Z2D_OP_SRC_OVER && px_is_opaque(px))) ? 1u : 0u;
  const RGBA16 v = pixel_to_rgba16(px.format, px.r, px.g, px.b, px.a);
  DrawIn in;
  in.surface = pl.slot;
  in.opts = pack_draw_opts(0, aa, o->fill_rule, o->op, precision, reduces, 0);
  in.paint_raw = pixel_to_raw(s->fmt, px.format, px.r, px.g, px.b, px.a);
  in.px_rgba = (uint32_t)v.r | ((uint32_t)v.g << 8) | ((uint32_t)v.b << 16) | ((uint32_t)v.a << 24);
  in.src_index = kNoIndex;
  in.stroke_index = kNoIndex;
  in.tolerance = o->tolerance > 0.001 ? o->tolerance : 0.001;
  B.draws.p[pl.draw_idx] = in;
}

template <class F>
static void parallel_for(size_t n, unsigned threads, F&& fn) {  // fn(begin, end) on contiguous slices
  if (threads <= 1 || n < 2 * threads) return fn((size_t)0, n);
  std::vector<std::thread> pool;
  const size_t per = (n + threads - 1) / threads;
  for (unsigned t = 1; t < threads; t++) {
    const size_t b = std::min(n, t * per), e = std::min(n, b + per);
    if (b < e) pool.emplace_back([&fn, b, e] { fn(b, e); });
  }
  fn((size_t)0, std::min(n, per));
  for (auto& th : pool) th.join();
}

// Records cmds[0..n) if every one of them is a plain fill; returns false (nothing recorded) otherwise.
static bool submit_fills_parallel(z2d_ctx* c, const z2d_draw_cmd* cmds, size_t n, int32_t* statuses, int32_t& first_err) {
  std::vector<FillPlan>& plan = c->fill_plans;
  plan.resize(n);
  parallel_for(n, c->record_threads, [&](size_t b, size_t e) {
    for (size_t i = b; i < e; i++) plan_fill(c, cmds[i], plan[i]);
  });
  Batch& B = *c->rec;
  size_t nodes = B.nodes.n, sps = B.subpaths.n, draws = B.draws.n;
  uint32_t n_par = 0;
  for (size_t i = 0; i < n; i++) {
    FillPlan& pl = plan[i];
    if (!pl.eligible) return false;
    if (pl.status != Z2D_OK || pl.skip) continue;
    pl.node_base = (uint32_t)nodes;
    pl.sp_base = (uint32_t)sps;
    pl.draw_idx = (uint32_t)draws;
    nodes += pl.m;
    sps += pl.n_sp;
    draws += 1;
    n_par += pl.n_par;
  }
  if (nodes > kMaxBatchNodes || draws > kMaxBatchDraws) return false;
  if (!B.nodes.reserve(nodes) || !B.subpaths.reserve(sps) || !B.draws.reserve(draws)) return false;
  for (size_t i = 0; i < n; i++)  // surface slots are batch state: serial
    if (plan[i].status == Z2D_OK && !plan[i].skip) plan[i].slot = batch_slot(c, cmds[i].surface);
  parallel_for(n, c->record_threads, [&](size_t b, size_t e) {
    for (size_t i = b; i < e; i++)
      if (plan[i].status == Z2D_OK && !plan[i].skip) write_fill(B, cmds[i], plan[i]);
  });
  B.nodes.n = nodes;
  B.subpaths.n = sps;
  B.draws.n = draws;
  B.n_par_sp += n_par;
  for (size_t i = 0; i < n; i++) {
    if (statuses) statuses[i] = plan[i].status;
    if (plan[i].status != Z2D_OK && first_err == Z2D_OK) first_err = plan[i].status;
  }
  return true;
}

extern "C" {

int32_t z2d_submit(z2d_ctx* c, const z2d_draw_cmd* cmds, size_t n, int32_t* statuses) {
  if (!c || (n && !cmds)) return Z2D_E_INVALID_ARG;
  int32_t first = Z2D_OK;
  constexpr size_t kParallelMin = 2048;
  size_t i = 0;
  // Pipelined chunks of one long call: recording and executing run side by side.  A remainder shorter than half a chunk is
  // taken along instead of becoming a batch of its own (every batch has ~0.5 ms of fixed cost); 100 000 fills are
  // 32768 + 32768 + 34464.  (Starting with a smaller chunk and growing was measured too: no difference beyond the noise.)
  const size_t target = c->chunk_draws;
  while (i < n) {
    // a run that fits the current batch: recorded by several threads when it is long enough and all plain fills
    const size_t room = target ? (c->rec->draws.n < target ? target - c->rec->draws.n : 1) : n - i;
    size_t run = std::min(n - i, room);
    if (target && n - i - run < target / 2) run = std::min(n - i, kMaxBatchDraws / 2);  // absorb a short tail
    if (run >= kParallelMin && c->record_threads > 1 && submit_fills_parallel(c, cmds + i, run, statuses ? statuses + i : nullptr, first)) {
      i += run;
      if (target && c->rec->draws.n >= target && i < n) {
        const int rc = kick(c);  // asynchronous: the worker executes this batch while the next run is recorded
        if (rc != Z2D_OK && first == Z2D_OK) first = rc;
      }
      continue;
    }
    const size_t end = i + (run >= kParallelMin ? run : 1);  // (a mixed run falls back to the one-by-one path)
    for (; i < end; i++) {
      const z2d_draw_cmd& k = cmds[i];
      int32_t rc = k.kind == 0 ? z2d_fill(c, k.surface, k.pattern, k.nodes, k.n_nodes, k.fill)
                               : z2d_stroke(c, k.surface, k.pattern, k.nodes, k.n_nodes, k.stroke);
      if (statuses) statuses[i] = rc;
      if (rc != Z2D_OK && first == Z2D_OK) first = rc;
    }
  }
  return first;
}

// ------------------------------------------------------------------------- compositor
int32_t z2d_composite(z2d_ctx* c, z2d_sfc* dst, int32_t dst_x, int32_t dst_y, const z2d_comp_op* ops, size_t n_ops, uint32_t precision) {
  if (!c || !dst || dst->ctx != c || (n_ops && !ops) || precision > 1) return Z2D_E_INVALID_ARG;
  if (n_ops > kMaxCompOps) return Z2D_E_INVALID_ARG;
  cudaSetDevice(c->device);
  int rc = flush(c);  // ordering: everything recorded so far lands first
  if (rc) return rc;
  if (dst->y0 != 0 || dst->vh != dst->h) {  // band destination: whole-band operations with generated sources only
    if (dst_x != 0 || dst_y != 0) return Z2D_E_INVALID_ARG;
    for (size_t k = 0; k < n_ops; k++)
      if (ops[k].src.kind == Z2D_PARAM_SURFACE || ops[k].dst.kind == Z2D_PARAM_SURFACE) return Z2D_E_INVALID_ARG;
  }
  // compositor.zig:311-374
  if (n_ops == 0) return Z2D_OK;
  if (dst_x >= dst->w || dst_y >= dst->h) return Z2D_OK;
  for (size_t k = 0; k < n_ops; k++) {
    if (ops[k].op >= Z2D_OP_COUNT) return Z2D_E_INVALID_ARG;
    if (op_requires_float(ops[k].op)) precision = Z2D_PRECISION_FLOAT;
  }
  int src_w, src_h;
  switch (ops[0].src.kind) {
    case Z2D_PARAM_NONE: return Z2D_OK;
    case Z2D_PARAM_SURFACE: {
      const z2d_sfc* ss = (const z2d_sfc*)ops[0].src.surface;
      if (!ss) return Z2D_E_INVALID_ARG;
      src_w = ss->w;
      src_h = ss->h;
      break;
    }
    default:
      if (dst_x != 0 || dst_y != 0) return Z2D_OK;
      src_w = dst->w;
      src_h = dst->h;
  }
  const int src_start_x = dst_x < 0 ? -dst_x : 0, src_start_y = dst_y < 0 ? -dst_y : 0;
  const int width = (src_w + dst_x > dst->w) ? dst->w - dst_x : src_w;
  const int height = (src_h + dst_y > dst->h) ? dst->h - dst_y : src_h;
  if (src_start_x >= width || src_start_y >= height) return Z2D_OK;

  CompArgs A;
  memset(&A, 0, sizeof A);
  A.data = dst->data;
  A.fmt = dst->fmt;
  A.w = dst->w;
  A.h = dst->h;
  A.y_origin = dst->y0;
  A.src_start_x = src_start_x;
  A.src_start_y = src_start_y;
  A.dst_start_x = src_start_x + dst_x;
  A.dst_start_y = src_start_y + dst_y;
  A.scan_w = width - src_start_x;
  A.rows = height - src_start_y;
  A.n_ops = (uint32_t)n_ops;
  A.precision = precision;
  std::vector<DevGrad> grads;
  std::vector<float> offs;
  std::vector<float4> cols;
  auto conv = [&](const z2d_comp_param& p, DevSrc& s, uint32_t& has, bool is_dst) -> int {
    has = p.kind != Z2D_PARAM_NONE;
    if (!has) return Z2D_OK;
    if (p.kind == Z2D_PARAM_SURFACE) {
      const z2d_sfc* ss = (const z2d_sfc*)p.surface;
      if (!ss || ss->ctx != c) return Z2D_E_INVALID_ARG;
      // every surface parameter must cover the rectangle it is read over: sources in source space, dst overrides in
      // destination space (the clipping above only looked at ops[0].src)
      const int need_x = (is_dst ? A.dst_start_x : A.src_start_x) + A.scan_w, need_y = (is_dst ? A.dst_start_y : A.src_start_y) + A.rows;
      if (ss->w < need_x || ss->h < need_y) return Z2D_E_INVALID_ARG;
      // the destination itself as a parameter is only well defined pixel for pixel (the reference's result for a shifted
      // self-composite depends on its scanline / vector order)
      if (ss == dst && (dst_x != 0 || dst_y != 0)) return Z2D_E_INVALID_ARG;
      memset(&s, 0, sizeof s);
      s.kind = Z2D_PARAM_SURFACE;
      s.sdata = ss->data;
      s.sfmt = ss->fmt;
      s.sw = ss->w;
      s.sh = ss->h;
      return Z2D_OK;
    }
    return pattern_to_src(p.pattern, s, grads, offs, cols);
  };
  for (size_t k = 0; k < n_ops; k++) {
    A.ops[k].op = ops[k].op;
    if ((rc = conv(ops[k].dst, A.ops[k].dst, A.ops[k].has_dst, true))) return rc;
    if ((rc = conv(ops[k].src, A.ops[k].src, A.ops[k].has_src, false))) return rc;
  }
  NvtxRange r_comp("z2d K5 composite");
  CK(c, upload(c, c->d_comp_grads, grads.data(), grads.size() * sizeof(DevGrad)));
  CK(c, upload(c, c->d_comp_stop_off, offs.data(), offs.size() * 4));
  CK(c, upload(c, c->d_comp_stop_col, cols.data(), cols.size() * sizeof(float4)));
  for (const DevGrad& g : grads) A.max_stops = std::max(A.max_stops, g.n_stops);
  A.T = tables(c, c->d_comp_grads, c->d_comp_stop_off, c->d_comp_stop_col);
  launch_composite(A, c->sm_count, c->stream);
  CK(c, cudaGetLastError());
  c->stats.kernel_launches = 1;
  c->stats_pending = false;
  return Z2D_OK;
}

}  // extern "C"
