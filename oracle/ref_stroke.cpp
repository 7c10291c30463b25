// ORACLE -- test infrastructure only (see ref_internal.h).
// Stroke tessellation: src/internal/tess/{stroke_plotter,dashed_plotter,Face,
// Pen,Slope,Dasher,point_buffer}.zig and arc.zig:92-267
// (transformed_circle_major_axis).  Contours are kept as linked lists exactly
// like the reference (append / prepend / insert-before), then turned into
// polygon edges.
#include <algorithm>
#include <list>

#include "ref_internal.h"

namespace zref {
namespace {

constexpr double kEps = 2.220446049250313e-16;  // math.floatEps(f64)

inline int sgn(double v) { return (v > 0) - (v < 0); }

struct Slope {  // tess/Slope.zig
  double dx, dy;
  static Slope init(Pt a, Pt b) { return {b.x - a.x, b.y - a.y}; }
  double normalize() {  // Slope.zig:174-214
    double rdx, rdy, mag;
    if (dx == 0.0) {
      rdx = 0.0;
      if (dy > 0.0) {
        mag = dy;
        rdy = 1.0;
      } else {
        mag = -dy;
        rdy = -1.0;
      }
    } else if (dy == 0.0) {
      rdy = 0.0;
      if (dx > 0.0) {
        mag = dx;
        rdx = 1.0;
      } else {
        mag = -dx;
        rdx = -1.0;
      }
    } else {
      mag = std::hypot(dx, dy);
      rdx = dx / mag;
      rdy = dy / mag;
    }
    dx = rdx;
    dy = rdy;
    return mag;
  }
};

int slope_compare(Slope a, Slope b) {  // Slope.zig:46-85
  double bdy = std::fabs(b.dy - a.dy) > kEps ? b.dy : a.dy;
  double bdx = std::fabs(b.dx - a.dx) > kEps ? b.dx : a.dx;
  int cmp = sgn(a.dy * bdx - bdy * a.dx);
  if (cmp != 0) return cmp;
  if (a.dx == 0 && a.dy == 0 && bdx == 0 && bdy == 0) return 0;
  if (a.dx == 0 && a.dy == 0) return 1;
  if (bdx == 0 && bdy == 0) return -1;
  if (sgn(a.dx) != sgn(bdx) || sgn(a.dy) != sgn(bdy)) return (a.dx > 0 || (a.dx == 0 && a.dy > 0)) ? -1 : 1;
  return 0;
}

bool compare_for_miter_limit(Slope in, Slope out, double ml) {  // Slope.zig:150-172
  in.normalize();
  out.normalize();
  double d = in.dx * out.dx + in.dy * out.dy;
  return 2 <= ml * ml * (1 + d);
}

double major_axis(const Xf& m, double radius) {  // arc.zig:92-267
  const double eps = 0.00390625;
  double det = m.ax * m.dy - m.by * m.cx;
  if (std::fabs(det * det - 1.0) < eps) {
    if (std::fabs(m.by) < eps && std::fabs(m.cx) < eps) return radius;
    if (std::fabs(m.ax) < eps && std::fabs(m.dy) < eps) return radius;
  }
  double i = m.ax * m.ax + m.by * m.by, j = m.cx * m.cx + m.dy * m.dy;
  double f = 0.5 * (i + j), g = 0.5 * (i - j), h = m.ax * m.cx + m.by * m.dy;
  return radius * std::sqrt(f + std::hypot(g, h));
}

struct PenVertex {
  Pt point;
  Slope cw, ccw;
};
struct Pen {  // tess/Pen.zig
  std::vector<PenVertex> v;
  void init(double thickness, double tol, const Xf& ctm) {  // Pen.zig:36-129
    double radius = thickness / 2;
    int n;
    double major = major_axis(ctm, radius);
    if (tol >= major * 4) {
      n = 1;
    } else if (tol >= major) {
      n = 4;
    } else {
      double delta = std::acos(1 - tol / major);
      if (delta == 0) {
        n = 4;
      } else {
        n = (int)std::ceil(2 * M_PI / delta);
        if (n < 4)
          n = 4;
        else if (n % 2 != 0)
          n = n + 1;
      }
    }
    bool reflect = ctm.det() < 0;
    v.resize((size_t)n);
    for (int i = 0; i < n; i++) {
      double t = 2 * M_PI * (double)i / (double)n;
      if (reflect) t = -t;
      double dx = radius * std::cos(t), dy = radius * std::sin(t);
      ctm.dist(dx, dy);
      v[(size_t)i].point = {dx, dy};
    }
    for (int i = 0; i < n; i++) {
      int next = (i >= n - 1) ? 0 : i + 1;
      int prev = std::max(0, i == 0 ? n - 1 : i - 1);
      v[(size_t)i].cw = Slope::init(v[(size_t)prev].point, v[(size_t)i].point);
      v[(size_t)i].ccw = Slope::init(v[(size_t)i].point, v[(size_t)next].point);
    }
  }
  // vertexIteratorFor (Pen.zig:138-232): returns [start,end) walk order
  void range(Slope from, Slope to, bool clockwise, size_t& start_o, size_t& end_o) const {
    int n = (int)v.size();
    int start = 0, end = 0;
    if (clockwise) {
      int low = 0, high = n, i = (low + high) >> 1;
      while (high - low > 1) {
        if (slope_compare(v[(size_t)i].cw, from) < 0)
          low = i;
        else
          high = i;
        i = (low + high) >> 1;
      }
      if (slope_compare(v[(size_t)i].cw, from) < 0) {
        i += 1;
        if (i == n) i = 0;
      }
      start = i;
      if (slope_compare(to, v[(size_t)i].ccw) >= 0) {
        low = i;
        high = i + n;
        i = (low + high) >> 1;
        while (high - low > 1) {
          int j = i >= n ? i - n : i;
          if (slope_compare(v[(size_t)j].cw, to) > 0)
            high = i;
          else
            low = i;
          i = (low + high) >> 1;
        }
        if (i >= n) i -= n;
      }
      end = i;
    } else {
      int low = 0, high = n, i = (low + high) >> 1;
      while (high - low > 1) {
        if (slope_compare(from, v[(size_t)i].ccw) < 0)
          low = i;
        else
          high = i;
        i = (low + high) >> 1;
      }
      if (slope_compare(from, v[(size_t)i].ccw) < 0) {
        i += 1;
        if (i == n) i = 0;
      }
      start = i;
      if (slope_compare(v[(size_t)i].cw, to) <= 0) {
        low = i;
        high = i + n;
        i = (low + high) >> 1;
        while (high - low > 1) {
          int j = i >= n ? i - n : i;
          if (slope_compare(to, v[(size_t)j].ccw) > 0)
            high = i;
          else
            low = i;
          i = (low + high) >> 1;
        }
        if (i >= n) i -= n;
      }
      end = i;
    }
    start_o = (size_t)std::max(0, start);
    end_o = (size_t)std::max(0, end);
  }
  template <class F>
  void walk(Slope from, Slope to, bool clockwise, F&& f) const {  // VertexIterator.next
    size_t idx, end;
    range(from, to, clockwise, idx, end);
    while (idx != end) {
      const PenVertex& r = v[idx];
      if (clockwise) {
        idx += 1;
        if (idx == v.size()) idx = 0;
      } else {
        if (idx == 0) idx = v.size();
        idx -= 1;
      }
      f(r);
    }
  }
};

struct Face {  // tess/Face.zig
  Pt p0, p1;
  double width, half_width;
  Slope dev_slope, user_slope;
  Pt p0_cw, p0_ccw, p1_cw, p1_ccw;
  Xf ctm;
  static Face make(Pt p0, Pt p1, Slope dev_slope, double thickness, const Xf& ctm) {  // Face.zig:65-115
    Face f;
    double hw = thickness / 2, ox, oy;
    Slope us = dev_slope;
    if (!ctm.is_identity()) {
      double dx = dev_slope.dx, dy = dev_slope.dy;
      Xf inv;
      ctm.inverse(inv);
      inv.dist(dx, dy);
      us = {dx, dy};
      us.normalize();
      if (ctm.det() >= 0) {
        ox = -us.dy * hw;
        oy = us.dx * hw;
      } else {
        ox = us.dy * hw;
        oy = -us.dx * hw;
      }
      ctm.dist(ox, oy);
    } else {
      ox = -dev_slope.dy * hw;
      oy = dev_slope.dx * hw;
    }
    double cx = ox, cy = oy, ccx = -cx, ccy = -cy;
    f.p0 = p0; f.p1 = p1; f.width = thickness; f.half_width = hw;
    f.dev_slope = dev_slope; f.user_slope = us;
    f.p0_cw = {p0.x + cx, p0.y + cy};
    f.p0_ccw = {p0.x + ccx, p0.y + ccy};
    f.p1_cw = {p1.x + cx, p1.y + cy};
    f.p1_ccw = {p1.x + ccx, p1.y + ccy};
    f.ctm = ctm;
    return f;
  }
  static Face init(Pt p0, Pt p1, double thickness, const Xf& ctm) {
    Slope s = Slope::init(p0, p1);
    s.normalize();
    return make(p0, p1, s, thickness, ctm);
  }
  Pt intersect(const Face& out, bool clockwise) const {  // Face.zig:117-152
    Pt ip = clockwise ? p1_ccw : p1_cw;
    Pt op = clockwise ? out.p0_ccw : out.p0_cw;
    Slope is = dev_slope, os = out.dev_slope;
    is.normalize();
    os.normalize();
    double ry = ((op.x - ip.x) * is.dy * os.dy - op.y * os.dx * is.dy + ip.y * is.dx * os.dy) / (is.dx * os.dy - os.dx * is.dy);
    double rx = (std::fabs(is.dy) >= std::fabs(os.dy)) ? (ry - ip.y) * is.dx / is.dy + ip.x : (ry - op.y) * os.dx / os.dy + op.x;
    return {rx, ry};
  }
  template <class F>
  void cap(uint32_t mode, bool clockwise, const Pen* pen, F&& line_to) const {  // Face.zig:186-284
    switch (mode) {
      case Z2D_CAP_BUTT:
        if (clockwise) {
          line_to(p1_ccw);
          line_to(p1_cw);
        } else {
          line_to(p1_cw);
          line_to(p1_ccw);
        }
        break;
      case Z2D_CAP_SQUARE: {
        double ox = user_slope.dx * half_width, oy = user_slope.dy * half_width;
        ctm.dist(ox, oy);
        if (clockwise) {
          line_to(p1_ccw);
          line_to({p1_ccw.x + ox, p1_ccw.y + oy});
          line_to({p1_cw.x + ox, p1_cw.y + oy});
          line_to(p1_cw);
        } else {
          line_to(p1_cw);
          line_to({p1_cw.x + ox, p1_cw.y + oy});
          line_to({p1_ccw.x + ox, p1_ccw.y + oy});
          line_to(p1_ccw);
        }
        break;
      }
      default: {
        line_to(clockwise ? p1_ccw : p1_cw);
        pen->walk(dev_slope, Slope{-dev_slope.dx, -dev_slope.dy}, clockwise,
                  [&](const PenVertex& v) { line_to({p1.x + v.point.x, p1.y + v.point.y}); });
        line_to(clockwise ? p1_cw : p1_ccw);
      }
    }
  }
  template <class F>
  void cap_p0(uint32_t mode, bool clockwise, const Pen* pen, F&& line_to) const {
    Face rev = init(p1, p0, width, ctm);
    rev.cap(mode, clockwise, pen, line_to);
  }
  template <class F>
  void cap_p1(uint32_t mode, bool clockwise, const Pen* pen, F&& line_to) const {
    cap(mode, clockwise, pen, line_to);
  }
};

struct PointBuf25 {  // PointBuffer(2, 5)
  Pt items[5];
  size_t len = 0;
  void add(Pt p) {
    if (len < 5)
      items[len++] = p;
    else {
      items[2] = items[3];
      items[3] = items[4];
      items[4] = p;
    }
  }
  void reset() { len = 0; }
  Pt head(size_t n) const { return items[n]; }
  Pt tail(size_t n) const { return items[len - n]; }
  Pt first() const { return items[0]; }
  Pt last() const { return items[len - 1]; }
};

using Contour = std::list<Pt>;  // Polygon.Contour; points are stored pre-scaled

// The part of the plotter state the generic helpers (join, plotSingle,
// plotOpenJoined, plotClosedJoined) touch -- shared by Plotter and
// dashed_plotter.InitialPolygon.
struct PlotState {
  const StrokeParams* opts = nullptr;
  const Pen* pen = nullptr;
  int clockwise_ = -1;  // ?bool
  Polygon* result = nullptr;
  Contour outer, inner;
  void plot(Contour& c, Pt p, const Contour::iterator* before) {  // Contour.plot
    Pt s{p.x * opts->scale, p.y * opts->scale};
    if (before)
      c.insert(*before, s);
    else
      c.push_back(s);
  }
  void plot_reverse(Contour& c, Pt p) { c.push_front({p.x * opts->scale, p.y * opts->scale}); }
  void flush(Contour& c) {
    std::vector<Pt> v(c.begin(), c.end());
    result->add_contour(v);
  }
};

// stroke_plotter.join (stroke_plotter.zig:410-561).  `before` == insert-before
// node for the *outer contour* (null -> append); inner emission always prepends.
void join(PlotState& st, uint32_t join_mode, Pt p0, Pt p1, Pt p2, const Contour::iterator* before) {
  if (pt_eq(p0, p1) || pt_eq(p1, p2)) {
    if (st.clockwise_ < 0) st.clockwise_ = 0;
    return;
  }
  const StrokeParams& o = *st.opts;
  Face in = Face::init(p0, p1, o.thickness, o.ctm), out = Face::init(p1, p2, o.thickness, o.ctm);
  bool join_cw = slope_compare(in.dev_slope, out.dev_slope) < 0;
  bool poly_cw = st.clockwise_ >= 0 ? (st.clockwise_ != 0) : join_cw;
  bool switched = join_cw != poly_cw;
  auto plot_outer = [&](Pt p) { st.plot(st.outer, p, before); };
  auto plot_inner = [&](Pt p) { st.plot_reverse(st.inner, p); };
  auto outer_j = [&](Pt p) { switched ? plot_inner(p) : plot_outer(p); };
  auto inner_j = [&](Pt p) { switched ? plot_outer(p) : plot_inner(p); };

  if (slope_compare(in.dev_slope, out.dev_slope) == 0) {
    outer_j(join_cw ? in.p1_ccw : in.p1_cw);
    inner_j(join_cw ? in.p1_cw : in.p1_ccw);
    if (st.clockwise_ < 0) st.clockwise_ = poly_cw;
    return;
  }
  switch (join_mode) {
    case Z2D_JOIN_MITER:
    case Z2D_JOIN_BEVEL:
      if (join_mode == Z2D_JOIN_MITER && compare_for_miter_limit(in.dev_slope, out.dev_slope, o.miter_limit)) {
        outer_j(in.intersect(out, join_cw));
      } else {
        outer_j(join_cw ? in.p1_ccw : in.p1_cw);
        outer_j(join_cw ? out.p0_ccw : out.p0_cw);
      }
      break;
    default:
      outer_j(join_cw ? in.p1_ccw : in.p1_cw);
      st.pen->walk(in.dev_slope, out.dev_slope, join_cw,
                   [&](const PenVertex& v) { outer_j({p1.x + v.point.x, p1.y + v.point.y}); });
      outer_j(join_cw ? out.p0_ccw : out.p0_cw);
  }
  inner_j(join_cw ? in.p1_cw : in.p1_ccw);
  inner_j(p1);
  inner_j(join_cw ? out.p0_cw : out.p0_ccw);
  if (st.clockwise_ < 0) st.clockwise_ = poly_cw;
}

void plot_single(PlotState& st, Pt start, Pt end) {  // stroke_plotter.zig:251-294
  const StrokeParams& o = *st.opts;
  if (!st.inner.empty()) g_assert_trips++;  // stroke_plotter.zig:254 debug.assert(self.inner.len == 0)
  Face f = Face::init(start, end, o.thickness, o.ctm);
  auto lt = [&](Pt p) { st.plot(st.outer, p, nullptr); };
  f.cap_p0(o.cap, true, st.pen, lt);
  f.cap_p1(o.cap, true, st.pen, lt);
  st.flush(st.outer);
  st.outer.clear();
  st.clockwise_ = -1;
}

void plot_open_joined(PlotState& st, Pt start0, Pt end0, Pt start1, Pt end1) {  // stroke_plotter.zig:296-364
  const StrokeParams& o = *st.opts;
  Face fs = Face::init(start0, end0, o.thickness, o.ctm), fe = Face::init(start1, end1, o.thickness, o.ctm);
  bool cw = st.clockwise_ >= 0 ? (st.clockwise_ != 0) : true;
  if (st.outer.empty()) {
    fs.cap_p0(o.cap, cw, st.pen, [&](Pt p) { st.plot(st.outer, p, nullptr); });
  } else {
    Contour::iterator first = st.outer.begin();
    fs.cap_p0(o.cap, cw, st.pen, [&](Pt p) { st.plot(st.outer, p, &first); });
  }
  fe.cap_p1(o.cap, cw, st.pen, [&](Pt p) { st.plot(st.outer, p, nullptr); });
  st.outer.splice(st.outer.end(), st.inner);
  st.flush(st.outer);
  st.outer.clear();
  st.inner.clear();
  st.clockwise_ = -1;
}

void plot_closed_joined(PlotState& st, Pt initial0, Pt initial1, Pt p1, Pt p2) {  // stroke_plotter.zig:366-408
  const StrokeParams& o = *st.opts;
  if (!pt_eq(p2, initial0)) {
    join(st, o.join, p1, p2, initial0, nullptr);
    join(st, o.join, p2, initial0, initial1, nullptr);
  } else {
    join(st, o.join, p1, initial0, initial1, nullptr);
  }
  st.flush(st.outer);
  st.flush(st.inner);
  st.outer.clear();
  st.inner.clear();
  st.clockwise_ = -1;
}

// ------------------------------------------------------------- undashed
struct Plotter {
  PlotState st;
  Pen pen_storage;
  bool have_pen = false;
  PointBuf25 points;
  void ensure_pen() {
    if (!have_pen) {
      pen_storage.init(st.opts->thickness, st.opts->tolerance, st.opts->ctm);
      have_pen = true;
      st.pen = &pen_storage;
    }
  }
  int line_to(uint32_t join_mode, Pt p) {  // _runLineTo (112-131)
    if (points.len == 0) return Z2D_E_INVALID_STATE;
    if (pt_eq(p, points.last())) return Z2D_OK;
    points.add(p);
    if (points.len > 2) join(st, join_mode, points.tail(3), points.tail(2), points.tail(1), nullptr);
    return Z2D_OK;
  }
  void plot_dotted(Pt point) {  // 202-237
    if (!st.inner.empty()) g_assert_trips++;  // stroke_plotter.zig:211 debug.assert(self.inner.len == 0)
    if (st.opts->cap == Z2D_CAP_ROUND) {
      for (const PenVertex& v : st.pen->v) st.plot(st.outer, {point.x + v.point.x, point.y + v.point.y}, nullptr);
      st.flush(st.outer);
      st.outer.clear();
      st.clockwise_ = -1;
    }
  }
  void finish() {  // 182-200
    switch (points.len) {
      case 0:
      case 1: break;
      case 2: plot_single(st, points.head(0), points.head(1)); break;
      default: plot_open_joined(st, points.head(0), points.head(1), points.tail(2), points.tail(1));
    }
  }
  int run(const z2d_node* nodes, size_t n) {
    for (size_t i = 0; i < n; i++) {
      const z2d_node& nd = nodes[i];
      switch (nd.tag) {
        case Z2D_NODE_MOVE_TO:
          if (points.len > 0) finish();
          points.reset();
          points.add({nd.p[0], nd.p[1]});
          break;
        case Z2D_NODE_LINE_TO: {
          int rc = line_to(st.opts->join, {nd.p[0], nd.p[1]});
          if (rc) return rc;
          break;
        }
        case Z2D_NODE_CURVE_TO: {
          if (points.len == 0) return Z2D_E_INVALID_STATE;
          ensure_pen();
          Pt a = points.last();
          spline_decompose(a, {nd.p[0], nd.p[1]}, {nd.p[2], nd.p[3]}, {nd.p[4], nd.p[5]}, st.opts->tolerance,
                           [&](Pt p) { line_to(Z2D_JOIN_ROUND, p); });
          break;
        }
        default:  // close_path (157-180)
          switch (points.len) {
            case 0: break;
            case 1: plot_dotted(points.first()); break;
            case 2: plot_single(st, points.head(0), points.head(1)); break;
            default: plot_closed_joined(st, points.head(0), points.head(1), points.tail(2), points.tail(1));
          }
          points.reset();
      }
    }
    finish();
    return Z2D_OK;
  }
};

// ---------------------------------------------------------------- dashed
struct Dasher {  // tess/Dasher.zig
  const double* dashes;
  size_t n;
  double offset;
  size_t idx;
  bool on;
  double remain;
  static bool validate(const double* d, size_t n) {
    bool valid = false;
    for (size_t i = 0; i < n; i++) {
      if (d[i] < 0) return false;
      if (d[i] > 0) valid = true;
    }
    return valid;
  }
  void apply_offset() {
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
  void reset() {
    idx = 0;
    on = true;
    remain = dashes[0];
    apply_offset();
  }
  bool step(double len) {
    bool stepped = false;
    remain -= len;
    if (remain <= 0) {
      stepped = true;
      on = !on;
      idx += 1;
      if (idx >= n) idx = 0;
      remain = dashes[idx];
    }
    return stepped;
  }
};

struct DashedPlotter {
  PlotState st;
  Pen pen_storage;
  bool have_pen = false;
  PointBuf25 points;
  Dasher dasher;
  Slope current_slope{0, 0};
  enum { NONE, OFF, ON } initial_kind = NONE;
  Pt initial_off{0, 0};
  struct Initial {
    PlotState st;
    PointBuf25 points;
    Slope current_slope;
  } initial;

  void ensure_pen() {
    if (!have_pen) {
      pen_storage.init(st.opts->thickness, st.opts->tolerance, st.opts->ctm);
      have_pen = true;
      st.pen = &pen_storage;
    }
  }

  void plot_dotted(PlotState& s, Pt point, Slope slope) {  // dashed_plotter.zig:369-465 (always on self)
    const StrokeParams& o = *st.opts;
    (void)s;
    // dashed_plotter.zig:381 `debug.assert(self.inner.len == 0); // should have not been used`: the reference panics here in
    // Debug / ReleaseSafe builds (the modes its spec suite runs in) and has undefined behaviour in ReleaseFast.  What follows
    // is the literal continuation; the trip is counted so that tests can tell such calls apart (z2d_ref_assert_trips).
    if (!st.inner.empty()) g_assert_trips++;
    switch (o.cap) {
      case Z2D_CAP_ROUND:
        for (const PenVertex& v : st.pen->v) st.plot(st.outer, {point.x + v.point.x, point.y + v.point.y}, nullptr);
        st.flush(st.outer);
        break;
      case Z2D_CAP_SQUARE: {
        Face f = Face::make(point, point, slope, o.thickness, o.ctm);
        double ox = f.user_slope.dx * f.half_width, oy = f.user_slope.dy * f.half_width;
        o.ctm.dist(ox, oy);
        st.plot(st.outer, {f.p1_cw.x - ox, f.p1_cw.y - oy}, nullptr);
        st.plot(st.outer, {f.p1_cw.x + ox, f.p1_cw.y + oy}, nullptr);
        st.plot(st.outer, {f.p1_ccw.x + ox, f.p1_ccw.y + oy}, nullptr);
        st.plot(st.outer, {f.p1_ccw.x - ox, f.p1_ccw.y - oy}, nullptr);
        st.flush(st.outer);
        break;
      }
      default: break;
    }
    st.outer.clear();
    st.clockwise_ = -1;
  }

  void save_initial() {  // 467-520
    if (!dasher.on) {
      initial_kind = ON;
      initial.st.opts = st.opts;
      initial.st.pen = st.pen;
      initial.st.clockwise_ = st.clockwise_;
      initial.st.result = st.result;
      initial.st.outer.clear();
      initial.st.inner.clear();
      initial.st.outer.splice(initial.st.outer.end(), st.outer);
      initial.st.inner.splice(initial.st.inner.end(), st.inner);
      initial.points = points;
      initial.current_slope = current_slope;
    } else {
      initial_kind = OFF;
      initial_off = points.first();
    }
    st.outer.clear();
    st.inner.clear();
    st.clockwise_ = -1;
  }

  void emit_current() {  // the switch shared by nextSegment / finish
    switch (points.len) {
      case 0: break;
      case 1: plot_dotted(st, points.first(), current_slope); break;
      case 2: plot_single(st, points.head(0), points.head(1)); break;
      default: plot_open_joined(st, points.head(0), points.head(1), points.tail(2), points.tail(1));
    }
  }

  void next_segment(Pt point) {  // 307-332
    if (initial_kind == NONE)
      save_initial();
    else if (!dasher.on)
      emit_current();
    points.reset();
    points.add(point);
  }

  int finish_initial_dotted() {  // 522-531
    if (initial_kind != ON) return Z2D_E_INVALID_STATE;
    plot_dotted(st, initial.points.first(), initial.current_slope);
    initial_kind = NONE;
    return Z2D_OK;
  }
  int finish_initial(Pt last_point, Pt second_to_last) {  // 533-552
    if (initial_kind != ON) return Z2D_E_INVALID_STATE;
    if (initial.points.len < 2) return Z2D_E_INVALID_STATE;
    plot_open_joined(initial.st, last_point, second_to_last, initial.points.tail(2), initial.points.tail(1));
    initial_kind = NONE;
    return Z2D_OK;
  }

  int join_and_cap_initial() {  // 554-628
    if (initial_kind != ON) return Z2D_E_INVALID_STATE;
    if (points.len > 2) {
      if (initial.points.len < 2) return Z2D_E_INVALID_STATE;
      join(st, st.opts->join, points.tail(2), initial.points.head(0), initial.points.head(1), nullptr);
      // self.outer.concat(initial.outer); initial.inner.concat(self.inner); initial.outer = self.outer
      st.outer.splice(st.outer.end(), initial.st.outer);
      initial.st.inner.splice(initial.st.inner.end(), st.inner);
      initial.st.outer.clear();
      initial.st.outer.splice(initial.st.outer.end(), st.outer);
      plot_open_joined(initial.st, points.head(0), points.head(1), initial.points.tail(2), initial.points.tail(1));
    } else {
      if (initial.points.len < 2) return Z2D_E_INVALID_STATE;
      if (initial.st.outer.empty()) {
        join(initial.st, st.opts->join, points.tail(2), initial.points.head(0), initial.points.head(1), nullptr);
      } else {
        Contour::iterator first = initial.st.outer.begin();
        join(initial.st, st.opts->join, points.tail(2), initial.points.head(0), initial.points.head(1), &first);
      }
      plot_open_joined(initial.st, points.first(), initial.points.first(), initial.points.tail(2), initial.points.tail(1));
    }
    initial_kind = NONE;
    st.outer.clear();
    st.inner.clear();
    st.clockwise_ = -1;
    return Z2D_OK;
  }

  int line_to(uint32_t join_mode, Pt target) {  // _runLineTo (123-173)
    if (points.len == 0) return Z2D_E_INVALID_STATE;
    Pt current = points.last();
    if (pt_eq(target, current)) return Z2D_OK;
    const StrokeParams& o = *st.opts;
    Pt first_dash_point = current;
    Slope slope = Slope::init(first_dash_point, target);
    current_slope = slope;
    current_slope.normalize();
    Xf inv;
    o.ctm.inverse(inv);
    inv.dist(slope.dx, slope.dy);
    const double total_len = slope.normalize();
    double remaining = total_len;
    double step_len = std::min(dasher.remain, remaining);
    while (remaining > 0) {
      remaining -= step_len;
      double xo = slope.dx * (total_len - remaining), yo = slope.dy * (total_len - remaining);
      o.ctm.dist(xo, yo);
      Pt dp{first_dash_point.x + xo, first_dash_point.y + yo};
      if (!pt_eq(dp, points.last())) points.add(dp);
      if (dasher.on) {
        if (points.len > 2) join(st, join_mode, points.tail(3), points.tail(2), points.tail(1), nullptr);
      }
      if (dasher.step(step_len)) next_segment(dp);
      step_len = std::min(dasher.remain, remaining);
    }
    return Z2D_OK;
  }

  int finish() {  // 334-367
    switch (initial_kind) {
      case ON: {
        int rc;
        switch (initial.points.len) {
          case 0: return Z2D_E_INVALID_STATE;
          case 1: rc = finish_initial_dotted(); break;
          default: rc = finish_initial(initial.points.head(0), initial.points.head(1));
        }
        if (rc) return rc;
        break;
      }
      case OFF: initial_kind = NONE; break;
      default: break;
    }
    if (dasher.on) emit_current();
    return Z2D_OK;
  }

  int close_path() {  // 202-305
    if (points.len == 0) return Z2D_E_INVALID_STATE;
    Pt target;
    switch (initial_kind) {
      case ON:
        if (initial.points.len == 0) return Z2D_E_INVALID_STATE;
        target = initial.points.first();
        break;
      case OFF: target = initial_off; break;
      default: target = points.first();
    }
    int rc = line_to(st.opts->join, target);
    if (rc) return rc;
    switch (initial_kind) {
      case ON:
        if (dasher.on && points.len > 1) {
          if (initial.points.len == 1) {
            plot_open_joined(st, points.head(0), points.head(1), points.tail(2), points.tail(1));
            initial_kind = NONE;
          } else {
            rc = join_and_cap_initial();
            if (rc) return rc;
          }
        } else {
          switch (initial.points.len) {
            case 0: return Z2D_E_INVALID_STATE;
            case 1: rc = finish_initial_dotted(); break;
            default: rc = finish_initial(initial.points.head(0), initial.points.head(1));
          }
          if (rc) return rc;
        }
        break;
      case OFF: initial_kind = NONE; break;
      default:
        switch (points.len) {
          case 1: plot_dotted(st, points.first(), current_slope); break;
          case 2: plot_single(st, points.head(0), points.head(1)); break;
          default:
            join(st, st.opts->join, points.tail(2), points.head(0), points.head(1), nullptr);
            st.flush(st.outer);
            st.flush(st.inner);
            st.outer.clear();
            st.inner.clear();
            st.clockwise_ = -1;
        }
    }
    points.reset();
    return Z2D_OK;
  }

  int run(const z2d_node* nodes, size_t n) {
    for (size_t i = 0; i < n; i++) {
      const z2d_node& nd = nodes[i];
      int rc = Z2D_OK;
      switch (nd.tag) {
        case Z2D_NODE_MOVE_TO:
          rc = finish();
          dasher.reset();
          points.reset();
          points.add({nd.p[0], nd.p[1]});
          break;
        case Z2D_NODE_LINE_TO: rc = line_to(st.opts->join, {nd.p[0], nd.p[1]}); break;
        case Z2D_NODE_CURVE_TO: {
          if (points.len == 0) return Z2D_E_INVALID_STATE;
          ensure_pen();
          Pt a = points.last();
          spline_decompose(a, {nd.p[0], nd.p[1]}, {nd.p[2], nd.p[3]}, {nd.p[4], nd.p[5]}, st.opts->tolerance,
                           [&](Pt p) { line_to(Z2D_JOIN_ROUND, p); });
          break;
        }
        default: rc = close_path();
      }
      if (rc) return rc;
    }
    return finish();
  }
};

}  // namespace

int stroke_plot(const z2d_node* nodes, size_t n, const StrokeParams& sp, Polygon& out) {  // stroke_plotter.zig:40-75
  out.scale = 1;  // contours are pre-scaled (Polygon.zig:388-391); result polygon scale stays 1
  if (Dasher::validate(sp.dashes, sp.n_dashes)) {
    DashedPlotter p;
    p.st.opts = &sp;
    p.st.result = &out;
    if (sp.join == Z2D_JOIN_ROUND || sp.cap == Z2D_CAP_ROUND) p.ensure_pen();
    p.dasher.dashes = sp.dashes;
    p.dasher.n = sp.n_dashes;
    p.dasher.offset = sp.dash_offset;
    p.dasher.reset();  // Dasher.init
    return p.run(nodes, n);
  }
  Plotter p;
  p.st.opts = &sp;
  p.st.result = &out;
  if (sp.join == Z2D_JOIN_ROUND || sp.cap == Z2D_CAP_ROUND) p.ensure_pen();
  return p.run(nodes, n);
}

}  // namespace zref
