// ORACLE -- test infrastructure only (see ref_internal.h).
// Polygon edges, fill plotter, Bezier flattening and the three rasterisers:
// src/internal/tess/{Polygon,fill_plotter,Spline}.zig,
// src/internal/raster/{direct,multisample,supersample}.zig.
#include <algorithm>

#include "ref_internal.h"

namespace zref {

uint64_t g_assert_trips = 0;  // inputs on which the reference's own debug.assert would fire (ref_stroke.cpp)
uint64_t g_covered_px = 0;  // pixels with coverage > 0 composited by the MSAA rasteriser (benchmark statistic)

// ----------------------------------------------------------------- Polygon
void Polygon::add_edge(Pt p0, Pt p1) {  // Polygon.zig:61-109
  Pt a{p0.x * scale, p0.y * scale}, b{p1.x * scale, p1.y * scale};
  Edge e;
  if (a.y < b.y) {
    e = {a.y, b.y, a.x, (b.x - a.x) / (b.y - a.y)};
  } else if (a.y > b.y) {
    e = {a.y, b.y, b.x, (a.x - b.x) / (a.y - b.y)};
  } else {
    return;
  }
  double t = e.top(), bt = e.bottom();
  double l = a.x < b.x ? a.x : b.x, r = a.x < b.x ? b.x : a.x;
  if (edges.empty()) {
    ext_top = t; ext_bottom = bt; ext_left = l; ext_right = r;
  } else {
    if (t < ext_top) ext_top = t;
    if (bt > ext_bottom) ext_bottom = bt;
    if (l < ext_left) ext_left = l;
    if (r > ext_right) ext_right = r;
  }
  edges.push_back(e);
}

void Polygon::add_contour(const std::vector<Pt>& pts) {  // Polygon.zig:115-137
  if (pts.empty()) return;
  for (size_t i = 1; i < pts.size(); i++) add_edge(pts[i - 1], pts[i]);
  add_edge(pts.back(), pts.front());
}

bool Polygon::in_box(double sc, int bw, int bh) const {  // Polygon.zig:142-201
  if (ext_right < 0.0 || ext_bottom < 0.0) return false;
  int sx = (int)std::floor(ext_left / sc), sy = (int)std::floor(ext_top / sc);
  int ex = (int)std::ceil(ext_right / sc), ey = (int)std::ceil(ext_bottom / sc);
  int pw = ex - sx, ph = ey - sy;
  if (pw == 0 || ph == 0) return false;
  if (sx + pw < 0 || sy + ph < 0) return false;
  if (sx >= bw || sy >= bh) return false;
  return true;
}

// ------------------------------------------------------------------ Spline
static inline double dot_sq(double x, double y) { return x * x + y * y; }

double Knots::error_sq() const {  // Spline.zig:83-123
  double bx = b.x - a.x, by = b.y - a.y, cx = c.x - a.x, cy = c.y - a.y;
  if (a.x != d.x || a.y != d.y) {
    double dx = d.x - a.x, dy = d.y - a.y;
    double dd = dot_sq(dx, dy);
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
  double be = dot_sq(bx, by), ce = dot_sq(cx, cy);
  return be > ce ? be : ce;
}

static inline Pt lerp_half(Pt a, Pt b) { return {a.x + ((b.x - a.x) / 2), a.y + ((b.y - a.y) / 2)}; }

Knots Knots::de_casteljau() {  // Spline.zig:128-151
  Pt ab = lerp_half(a, b), bc = lerp_half(b, c), cd = lerp_half(c, d);
  Pt abbc = lerp_half(ab, bc), bccd = lerp_half(bc, cd);
  Pt fin = lerp_half(abbc, bccd);
  Knots r{fin, bccd, cd, d};
  b = ab;
  c = abbc;
  d = fin;
  return r;
}

// ------------------------------------------------------------ fill_plotter
namespace {
struct PointBuf13 {  // PointBuffer(1, 3) (point_buffer.zig:14-28)
  Pt items[3];
  size_t len = 0;
  void add(Pt p) {
    if (len < 3)
      items[len++] = p;
    else {
      items[1] = items[2];
      items[2] = p;
    }
  }
  void reset() { len = 0; }
};
}  // namespace

int fill_plot(const z2d_node* nodes, size_t n, double scale, double tol, Polygon& out) {  // fill_plotter.zig:21-97
  out.scale = scale;
  PointBuf13 pts;
  for (size_t i = 0; i < n; i++) {
    const z2d_node& nd = nodes[i];
    switch (nd.tag) {
      case Z2D_NODE_MOVE_TO:
        if (i == n - 1) return Z2D_OK;  // trailing auto move_to
        pts.reset();
        pts.add({nd.p[0], nd.p[1]});
        break;
      case Z2D_NODE_LINE_TO: {
        if (pts.len == 0) return Z2D_E_INVALID_STATE;
        Pt p{nd.p[0], nd.p[1]}, last = pts.items[pts.len - 1];
        if (!pt_eq(last, p)) {
          out.add_edge(last, p);
          pts.add(p);
        }
        break;
      }
      case Z2D_NODE_CURVE_TO: {
        if (pts.len == 0) return Z2D_E_INVALID_STATE;
        Pt a = pts.items[pts.len - 1];
        spline_decompose(a, {nd.p[0], nd.p[1]}, {nd.p[2], nd.p[3]}, {nd.p[4], nd.p[5]}, tol, [&](Pt p) {
          Pt last = pts.items[pts.len - 1];
          if (!pt_eq(last, p)) {
            out.add_edge(last, p);
            pts.add(p);
          }
        });
        break;
      }
      default:  // close_path (fill_plotter.zig:72-92)
        if (pts.len >= 3) {
          Pt last = pts.items[pts.len - 1], first = pts.items[0];
          if (pt_eq(last, first)) break;
          out.add_edge(last, first);
          pts.add(first);
        }
    }
  }
  return Z2D_OK;
}

// ---------------------------------------------------------- WorkingEdgeSet
namespace {
struct WorkingEdgeSet {  // Polygon.zig:203-354
  std::vector<Edge>& all;
  size_t n_active = 0;
  std::vector<int32_t> xs;
  std::vector<int32_t> bps;
  explicit WorkingEdgeSet(Polygon& p) : all(p.edges) {
    xs.resize(all.size());
    for (const Edge& e : all) {  // breakpoints(): sorted unique
      bps.push_back((int32_t)zround(e.top()));
      bps.push_back((int32_t)zround(e.bottom()));
    }
    std::sort(bps.begin(), bps.end());
    bps.erase(std::unique(bps.begin(), bps.end()), bps.end());
  }
  void rescan(int line_y) {
    double mid = (double)line_y + 0.5;
    size_t to = 0;
    for (size_t from = 0; from < all.size(); from++) {
      if (all[from].top() < mid && all[from].bottom() >= mid) {
        if (from != to) std::swap(all[to], all[from]);
        to++;
      }
    }
    n_active = to;
  }
  void inc(int y) {
    double mid = (double)y + 0.5;
    for (size_t i = 0; i < n_active; i++) {
      const Edge& e = all[i];
      xs[i] = (int32_t)zround(e.x_start + (e.x_inc * (mid - e.top())));
    }
  }
  // sort by x; edges and x values move together.  (pdq is unstable; ties do
  // not change coverage for closed contours.  We sort stably.)
  std::vector<std::pair<int32_t, Edge>> tmp;
  void sort() {
    tmp.resize(n_active);
    for (size_t i = 0; i < n_active; i++) tmp[i] = {xs[i], all[i]};
    std::stable_sort(tmp.begin(), tmp.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
    for (size_t i = 0; i < n_active; i++) {
      xs[i] = tmp[i].first;
      all[i] = tmp[i].second;
    }
  }
  // returns number of filtered x values (in xs[0..n))
  size_t filter(uint32_t rule) {
    if (rule == Z2D_FILL_EVEN_ODD) return n_active;
    int wind = 0;
    size_t to = 0;
    for (size_t from = 0; from < n_active; from++) {
      xs[to] = xs[from];
      if (wind == 0) {
        wind += all[from].dir();
        to++;
      } else {
        wind += all[from].dir();
        if (wind == 0) to++;
      }
    }
    return to;
  }
  // index of the first breakpoint to use: (first bp >= start) -| 1; -1 => none
  long first_bp(int start) const {
    for (size_t i = 0; i < bps.size(); i++)
      if (bps[i] >= start) return i == 0 ? 0 : (long)i - 1;
    return -1;
  }
};
inline int clampi(int v, int lo, int hi) { return std::max(lo, std::min(v, hi)); }
}  // namespace

// ------------------------------------------------------------ direct.zig
void raster_direct(Sfc& s, const Src& pat, Polygon& poly, uint32_t rule, uint32_t op, uint32_t prec) {
  const int W = s.w, H = s.h;
  if (op_requires_float(op)) prec = Z2D_PRECISION_FLOAT;
  const bool bounded = op_is_bounded(op);
  if (!poly.in_box(1.0, W, H)) return;
  int py0 = bounded ? (int)std::floor(poly.ext_top) : 0;
  int py1 = bounded ? (int)std::ceil(poly.ext_bottom) : H - 1;
  int y0 = clampi(py0, 0, H - 1), y1 = clampi(py1, y0, H - 1);
  WorkingEdgeSet wes(poly);
  long bpi = wes.first_bp(y0);
  if (bpi < 0) return;
  for (int y = y0; y <= y1; y++) {
    if (y >= wes.bps[bpi]) {
      wes.rescan(y);
      if ((size_t)bpi < wes.bps.size() - 1) bpi++;
    }
    wes.inc(y);
    wes.sort();
    size_t nf = wes.filter(rule);
    if (!bounded && nf == 0) {
      sfc_clear_stride(s, 0, y, (size_t)W);
      continue;
    }
    for (size_t p = 0; p < nf / 2; p++) {
      int sx = std::max(0, wes.xs[p * 2]);
      if (sx >= W) break;
      int ex = clampi(wes.xs[p * 2 + 1], sx, W);
      int fl = ex - sx, ecl = W - ex;
      if (!bounded && sx > 0) sfc_clear_stride(s, 0, y, (size_t)sx);
      if (fl > 0) composite_opaque(op, s, pat, sx, y, (size_t)fl, prec);
      if (!bounded && ecl > 0) sfc_clear_stride(s, ex, y, (size_t)ecl);
    }
  }
}

// -------------------------------------------------------- multisample.zig
void raster_multisample(Sfc& s, const Src& pat, Polygon& poly, uint32_t rule, uint32_t op, uint32_t prec) {
  const int scale = 4, cov_full = 16, alpha_scale = 16;
  const int W = s.w, H = s.h;
  if (op_requires_float(op)) prec = Z2D_PRECISION_FLOAT;
  if (!poly.in_box(scale, W, H)) return;
  int y0 = clampi((int)std::floor(poly.ext_top / scale), 0, H - 1);
  int y1 = clampi((int)std::ceil(poly.ext_bottom / scale), y0, H - 1);
  int x0 = clampi((int)std::floor(poly.ext_left / scale), 0, W - 1);
  int x1 = clampi((int)std::ceil(poly.ext_right / scale), x0, W);
  int dw = x1 - x0;
  if (dw < 1) return;  // reference panics; cannot happen after in_box
  std::vector<uint8_t> cov((size_t)dw);
  const int x0s = x0 * scale, dws = dw * scale;

  if (!op_is_bounded(op)) {  // multisample.zig:96-110 (quirks preserved)
    for (int y = 0; y < std::max(0, y0); y++) sfc_clear_stride(s, 0, y, (size_t)W);
    for (int y = std::max(0, y1 + 1); y < std::max(0, W); y++) sfc_clear_stride(s, 0, y, (size_t)W);  // sic: sfc_width
    for (int y = std::max(0, y0); y <= std::max(0, y1); y++) {
      if (x0 > 0) sfc_clear_stride(s, 0, y, (size_t)x0);
      if (x1 < W) sfc_clear_stride(s, 0, y, (size_t)(W - x1));  // sic: starts at 0
    }
  }

  WorkingEdgeSet wes(poly);
  long bpi = wes.first_bp(y0);
  if (bpi < 0) return;
  for (int y = y0; y <= y1; y++) {
    std::fill(cov.begin(), cov.end(), 0);
    int cov_len = 0;
    for (int yo = 0; yo < 4; yo++) {
      int ys = y * scale + yo;
      if (ys >= wes.bps[bpi]) {
        wes.rescan(ys);
        if ((size_t)bpi < wes.bps.size() - 1) bpi++;
      }
      wes.inc(ys);
      wes.sort();
      size_t nf = wes.filter(rule);
      int x_min = 0;
      for (size_t p = 0; p < nf / 2; p++) {
        int sx = std::max(x_min, wes.xs[p * 2] - x0s);
        if (sx >= dws) break;
        int ex = clampi(wes.xs[p * 2 + 1] - x0s, sx, dws);
        if (ex - sx > 0) {
          // addSpan (multisample.zig:239-279): per-pixel coverage += samples covered
          for (int sxx = std::max(0, sx); sxx < ex; sxx++) cov[(size_t)(sxx / scale)]++;
          cov_len = std::max(cov_len, (ex + scale - 1) / scale);
        }
        x_min = ex;
      }
    }
    for (int cx = 0; cx < std::min(cov_len, dw); cx++) {
      int x = cx + x0;
      int c = std::min<int>(cov[(size_t)cx], cov_full);
      if (x >= W) break;
      if (c == 0) continue;
      g_covered_px++;
      if (c == cov_full)
        composite_opaque(op, s, pat, x, y, 1, prec);
      else
        composite_opacity(op, s, pat, x, y, 1, prec, (uint8_t)clampi(c * alpha_scale - 1, 0, 255));
    }
  }
}

// -------------------------------------------------------- supersample.zig
void raster_supersample(Sfc& s, const Src& pat, Polygon& poly, uint32_t rule, uint32_t op, uint32_t prec) {
  const int scale = 4;
  const int W = s.w, H = s.h;
  if (!poly.in_box(scale, W, H)) return;
  const bool bounded = op_is_bounded(op);
  int x0 = bounded ? (int)std::floor(poly.ext_left / scale) : 0;
  int y0 = bounded ? (int)std::floor(poly.ext_top / scale) : 0;
  int x1 = bounded ? (int)std::ceil(poly.ext_right / scale) : W;
  int y1 = bounded ? (int)std::ceil(poly.ext_bottom / scale) : H;
  const int tws = W * scale, ths = H * scale;
  int bx0 = clampi(x0 * scale, 0, tws - 1), by0 = clampi(y0 * scale, 0, ths - 1);
  int bx1 = clampi(x1 * scale, bx0, tws - 1), by1 = clampi(y1 * scale, by0, ths - 1);
  int mw = (bx1 + 1) - bx0, mh = (by1 + 1) - by0;
  if (mw < 1 || mh < 1) return;
  uint32_t mfmt = (s.fmt == Z2D_FMT_ALPHA4 || s.fmt == Z2D_FMT_ALPHA2 || s.fmt == Z2D_FMT_ALPHA1) ? s.fmt : (uint32_t)Z2D_FMT_ALPHA8;
  z2d_pixel opaque_px{mfmt, 0, 0, 0, (uint8_t)(mfmt == Z2D_FMT_ALPHA8 ? 255 : mfmt == Z2D_FMT_ALPHA4 ? 15 : mfmt == Z2D_FMT_ALPHA2 ? 3 : 1)};
  std::vector<uint8_t> mbuf(sfc_byte_len(mfmt, mw, mh), 0);
  Sfc mask{mbuf.data(), mfmt, mw, mh};

  WorkingEdgeSet wes(poly);
  long bpi = wes.first_bp(by0);
  if (bpi < 0) return;
  for (int y = 0; y < mh; y++) {
    int dev_y = y + by0;
    if (dev_y >= wes.bps[bpi]) {
      wes.rescan(dev_y);
      if ((size_t)bpi < wes.bps.size() - 1) bpi++;
    }
    wes.inc(dev_y);
    wes.sort();
    size_t nf = wes.filter(rule);
    for (size_t p = 0; p < nf / 2; p++) {
      int sx = std::max(0, wes.xs[p * 2] - bx0);
      if (sx >= mw) break;
      int ex = clampi(wes.xs[p * 2 + 1] - bx0, sx, mw);
      if (ex - sx > 0) sfc_paint_stride(mask, sx, y, (size_t)(ex - sx), opaque_px);
    }
  }
  sfc_downsample(mask);

  SurfOp ops[2];
  ops[0].op = Z2D_OP_DST_IN;
  ops[0].dst = pat;
  ops[0].src.kind = Z2D_PARAM_SURFACE;
  ops[0].src.sfc = &mask;
  ops[1].op = op;
  surface_run(s, std::max(0, x0), std::max(0, y0), ops, 2, prec);
}

}  // namespace zref
