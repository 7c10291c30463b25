// ORACLE -- test infrastructure only (see ref_internal.h).
// Hairline stroking: src/internal/tess/polyline_plotter.zig and
// src/internal/raster/hairline.zig.
#include <algorithm>

#include "ref_internal.h"

namespace zref {
namespace {

struct HDasher {  // tess/Dasher.zig (same state machine as in ref_stroke.cpp)
  const double* dashes;
  size_t n;
  double offset;
  size_t idx;
  bool on;
  double remain;
  void reset() {
    idx = 0;
    on = true;
    remain = dashes[0];
    remain -= offset;
    while (remain < 0 || remain > dashes[idx]) {
      if (remain < 0) {
        remain += dashes[idx];
        idx = (idx >= n - 1) ? 0 : idx + 1;
      } else {
        remain -= dashes[idx];
        idx = (idx == 0) ? n - 1 : idx - 1;
      }
      on = !on;
    }
  }
  bool step(double len) {
    remain -= len;
    if (remain <= 0) {
      on = !on;
      idx += 1;
      if (idx >= n) idx = 0;
      remain = dashes[idx];
      return true;
    }
    return false;
  }
};

bool dashes_valid(const double* d, size_t n) {
  bool valid = false;
  for (size_t i = 0; i < n; i++) {
    if (d[i] < 0) return false;
    if (d[i] > 0) valid = true;
  }
  return valid;
}

// Slope.normalize (tess/Slope.zig:174-214)
double normalize(double& dx, double& dy) {
  double mag;
  if (dx == 0.0) {
    if (dy > 0.0) { mag = dy; dy = 1.0; } else { mag = -dy; dy = -1.0; }
    dx = 0.0;
  } else if (dy == 0.0) {
    if (dx > 0.0) { mag = dx; dx = 1.0; } else { mag = -dx; dx = -1.0; }
    dy = 0.0;
  } else {
    mag = std::hypot(dx, dy);
    dx = dx / mag;
    dy = dy / mag;
  }
  return mag;
}

struct PB12 {  // PointBuffer(1, 2)
  Pt items[2];
  size_t len = 0;
  void add(Pt p) {
    if (len < 2) items[len++] = p; else items[1] = p;
  }
  void reset() { len = 0; }
};

using Contours = std::vector<std::vector<Pt>>;

void dashed_line_to(Pt p0, Pt p1, HDasher& d, PB12& pts, std::vector<Pt>& cur, Contours& res) {  // polyline_plotter.zig:183-221
  double dx = p1.x - p0.x, dy = p1.y - p0.y;
  const double total = normalize(dx, dy);
  double remaining = total;
  double step = std::min(d.remain, remaining);
  while (remaining > 0) {
    remaining -= step;
    Pt dp{p0.x + dx * (total - remaining), p0.y + dy * (total - remaining)};
    if (!pt_eq(dp, pts.items[pts.len - 1])) pts.add(dp);
    if (d.on) cur.push_back(dp);
    if (d.step(step)) {
      res.push_back(cur);  // appended even when empty (len 0 contours are skipped by the rasteriser)
      cur.clear();
      if (d.on) cur.push_back(dp);
    }
    step = std::min(d.remain, remaining);
  }
}

int polyline_plot(const z2d_node* nodes, size_t n, double tol, const double* dashes, size_t nd, double doff, Contours& res) {
  std::vector<Pt> cur;
  PB12 pts;
  HDasher dasher{dashes, nd, doff, 0, true, 0};
  const bool dashed = dashes_valid(dashes, nd);
  if (dashed) dasher.reset();
  auto line_to = [&](Pt p) {
    Pt last = pts.items[pts.len - 1];
    if (!pt_eq(last, p)) {
      if (dashed) {
        dashed_line_to(last, p, dasher, pts, cur, res);
      } else {
        cur.push_back(p);
        pts.add(p);
      }
    }
  };
  for (size_t i = 0; i < n; i++) {
    const z2d_node& nd_ = nodes[i];
    switch (nd_.tag) {
      case Z2D_NODE_MOVE_TO:
        if (!cur.empty()) {
          res.push_back(cur);
          cur.clear();
        }
        pts.reset();
        if (dashed) dasher.reset();
        if (i == n - 1) goto done;
        cur.push_back({nd_.p[0], nd_.p[1]});
        pts.add({nd_.p[0], nd_.p[1]});
        break;
      case Z2D_NODE_LINE_TO:
        if (pts.len == 0) return Z2D_E_INVALID_STATE;
        line_to({nd_.p[0], nd_.p[1]});
        break;
      case Z2D_NODE_CURVE_TO:
        if (pts.len == 0) return Z2D_E_INVALID_STATE;
        spline_decompose(pts.items[pts.len - 1], {nd_.p[0], nd_.p[1]}, {nd_.p[2], nd_.p[3]}, {nd_.p[4], nd_.p[5]}, tol, line_to);
        break;
      default:
        if (pts.len >= 2) {
          Pt last = pts.items[pts.len - 1], first = pts.items[0];
          if (pt_eq(last, first)) break;
          if (dashed) {
            dashed_line_to(last, first, dasher, pts, cur, res);
          } else {
            cur.push_back(first);
            pts.add(first);
          }
        }
    }
  }
done:
  if (!cur.empty()) res.push_back(cur);
  return Z2D_OK;
}

struct DrawOpts {
  Sfc& s;
  const Src& pat;
  uint32_t op, prec, aa;
};

void delta_step(int a, int b, int& d, int& st) {
  int c = b - a;
  if (c < 0) { d = -c; st = -1; } else { d = c; st = 1; }
}
uint16_t err_inc(int a, int b) {  // hairline.zig:383-396
  if (a == b) return 0xFFFF;
  return (uint16_t)((((uint32_t)a) << 16) / (uint32_t)b);
}

void bres_h(const DrawOpts& o, int x0, int y0, int x1, int y1) {
  if (x0 > x1) return bres_h(o, x1, y1, x0, y0);
  int dx = x1 - x0, dy, sy;
  delta_step(y0, y1, dy, sy);
  int y = y0, d = 2 * dy - dx;
  for (int x = x0; x <= x1; x++) {
    composite_opaque(o.op, o.s, o.pat, x, y, 1, o.prec);
    if (d > 0) { y += sy; d -= 2 * dx; }
    d += 2 * dy;
  }
}
void bres_v(const DrawOpts& o, int x0, int y0, int x1, int y1) {
  if (y0 > y1) return bres_v(o, x1, y1, x0, y0);
  int dy = y1 - y0, dx, sx;
  delta_step(x0, x1, dx, sx);
  int x = x0, d = 2 * dx - dy;
  for (int y = y0; y <= y1; y++) {
    composite_opaque(o.op, o.s, o.pat, x, y, 1, o.prec);
    if (d > 0) { x += sx; d -= 2 * dy; }
    d += 2 * dx;
  }
}
void wu_h(const DrawOpts& o, int x0, int y0, int x1, int y1) {
  if (x0 > x1) return wu_h(o, x1, y1, x0, y0);
  int dx = x1 - x0, dy, sy;
  delta_step(y0, y1, dy, sy);
  int x = x0, y = y0;
  uint16_t err = 0;
  const uint16_t inc = err_inc(dy, dx);
  composite_opaque(o.op, o.s, o.pat, x, y, 1, o.prec);
  x += 1;
  for (; x < x1; x++) {
    uint32_t sum = (uint32_t)err + inc;
    err = (uint16_t)sum;
    if (sum > 0xFFFF) y += sy;
    uint8_t oc = (uint8_t)(err >> 8);
    composite_opacity(o.op, o.s, o.pat, x, y, 1, o.prec, (uint8_t)(oc ^ 0xFF));
    composite_opacity(o.op, o.s, o.pat, x, y + sy, 1, o.prec, oc);
  }
  composite_opaque(o.op, o.s, o.pat, x1, y1, 1, o.prec);
}
void wu_v(const DrawOpts& o, int x0, int y0, int x1, int y1) {
  if (y0 > y1) return wu_v(o, x1, y1, x0, y0);
  int dy = y1 - y0, dx, sx;
  delta_step(x0, x1, dx, sx);
  int x = x0, y = y0;
  uint16_t err = 0;
  const uint16_t inc = err_inc(dx, dy);
  composite_opaque(o.op, o.s, o.pat, x, y, 1, o.prec);
  y += 1;
  for (; y < y1; y++) {
    uint32_t sum = (uint32_t)err + inc;
    err = (uint16_t)sum;
    if (sum > 0xFFFF) x += sx;
    uint8_t oc = (uint8_t)(err >> 8);
    composite_opacity(o.op, o.s, o.pat, x, y, 1, o.prec, (uint8_t)(oc ^ 0xFF));
    composite_opacity(o.op, o.s, o.pat, x + sx, y, 1, o.prec, oc);
  }
  composite_opaque(o.op, o.s, o.pat, x1, y1, 1, o.prec);
}

void draw_line(const DrawOpts& o, int x0, int y0, int x1, int y1) {  // hairline.zig:94-123
  const int W = o.s.w, H = o.s.h;
  if ((x0 < 0 || x0 >= W) && (x1 < 0 || x1 >= W)) return;
  if ((y0 < 0 || y0 >= H) && (y1 < 0 || y1 >= H)) return;
  unsigned dx = (unsigned)std::abs(x1 - x0), dy = (unsigned)std::abs(y1 - y0);
  if (dx == 0) {
    int sy = std::max(0, std::min(std::min(y0, y1), H - 1)), ey = std::max(0, std::min(std::max(y0, y1), H - 1));
    for (int y = sy; y <= ey; y++) composite_opaque(o.op, o.s, o.pat, x0, y, 1, o.prec);
  } else if (dy == 0) {
    int sx = std::max(0, std::min(std::min(x0, x1), W - 1)), ex = std::max(0, std::min(std::max(x0, x1), W - 1));
    int len = ex - sx + 1;
    if (len > 0) composite_opaque(o.op, o.s, o.pat, sx, y0, (size_t)len, o.prec);
  } else if (dx < dy) {
    if (o.aa == Z2D_AA_NONE) bres_v(o, x0, y0, x1, y1); else wu_v(o, x0, y0, x1, y1);
  } else {
    if (o.aa == Z2D_AA_NONE) bres_h(o, x0, y0, x1, y1); else wu_h(o, x0, y0, x1, y1);
  }
}

}  // namespace

int hairline_stroke(Sfc& s, const Src& pat, const z2d_node* nodes, size_t n, double tol, const double* dashes, size_t n_dashes,
                    double dash_offset, uint32_t op, uint32_t prec, uint32_t aa) {
  Contours cs;
  int rc = polyline_plot(nodes, n, tol, dashes, n_dashes, dash_offset, cs);
  if (rc) return rc;
  if (op_requires_float(op)) prec = Z2D_PRECISION_FLOAT;
  DrawOpts o{s, pat, op, prec, aa};
  for (const auto& c : cs) {
    if (c.empty()) continue;
    if (c.size() == 1) {
      composite_opaque(op, s, pat, (int)zround(c[0].x), (int)zround(c[0].y), 1, prec);
      continue;
    }
    Pt pts[2];
    pts[0] = c[0];
    int idx = 1;
    for (size_t k = 1; k < c.size(); k++) {  // hairline.zig:58-78 (always draws slot 0 -> slot 1)
      pts[idx] = c[k];
      draw_line(o, (int)zround(pts[0].x), (int)zround(pts[0].y), (int)zround(pts[1].x), (int)zround(pts[1].y));
      idx ^= 1;
    }
  }
  return Z2D_OK;
}

}  // namespace zref
