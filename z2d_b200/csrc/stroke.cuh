// K2: stroke tessellation on the device, one thread per sub-path (included by kernels.cu).
//
// Offsets a polyline / flattened Bezier chain into the stroke outline exactly like the
// reference (tess/stroke_plotter.zig, dashed_plotter.zig, Face.zig, Slope.zig, Pen.zig,
// Dasher.zig) and feeds the outline's edges to an EdgeSink.  The reference keeps the outer
// and inner contours as linked lists (append / prepend / insert-before) and converts them to
// edges at the end; only *consecutive pairs* of contour points matter, so the device version
// streams edges as points arrive and keeps O(1) state per contour (first, last, length in
// list order).  The same walk runs twice (count + extents, then emit).
//
// The sequential state inside a sub-path (clockwise latch, dash phase, initial-dash splice)
// is preserved; parallelism is across sub-paths.  Pen vertices (sin/cos/acos) are computed
// on the host; the only transcendental needed here is hypot, implemented with the exact
// operation sequence of glibc's generic kernel so that it is bit-identical to the CPU.
#pragma once

namespace z2d {

struct PenV {  // tess/Pen.zig PenVertex
  double px, py, cwx, cwy, ccwx, ccwy;
};

struct StrokeCtx {
  uint32_t cap, join;
  double thickness, miter_limit, scale, tolerance, dash_offset;
  double ctm[6], inv[6];
  bool ctm_identity, det_nonneg;
  const PenV* pen;
  int npen;
  const double* dashes;
  int ndash;
};

struct Slope {
  double dx, dy;
};

// glibc 2.39 sysdeps/ieee754/dbl-64/e_hypot.c, non-FMA kernel, common (no scaling) case
Z2D_D double z_hypot(double x, double y) {
  x = fabs(x);
  y = fabs(y);
  const double ax = x < y ? y : x, ay = x < y ? x : y;
  if (ax * 0x1p-54 >= ay) return ax + ay;
  double t1, t2;
  double h = sqrt(ax * ax + ay * ay);
  if (h <= 2.0 * ay) {
    const double delta = h - ay;
    t1 = ax * (2.0 * delta - ax);
    t2 = (delta - 2.0 * (ax - ay)) * delta;
  } else {
    const double delta = h - ax;
    t1 = 2.0 * delta * (ax - 2.0 * ay);
    t2 = (4.0 * delta - ay) * ay + delta * delta;
  }
  h -= (t1 + t2) / (2.0 * h);
  return h;
}

// The stroker's arithmetic helpers are real functions, called with values only (no pointers into the caller's frame): inlined
// at every use they made k_flatten_count 173 000 instructions long, and its divergent threads then spent 68 % of their
// stall samples waiting for instruction fetch (profiles/r01_stroke_flatten_ncu.json).
struct NormSlope {
  double dx, dy, mag;
};
Z2D_DN NormSlope slope_normalized(double dx, double dy) {  // Slope.zig:174-214
  double rdx, rdy, mag;
  if (dx == 0.0) {
    rdx = 0.0;
    if (dy > 0.0) { mag = dy; rdy = 1.0; } else { mag = -dy; rdy = -1.0; }
  } else if (dy == 0.0) {
    rdy = 0.0;
    if (dx > 0.0) { mag = dx; rdx = 1.0; } else { mag = -dx; rdx = -1.0; }
  } else {
    mag = z_hypot(dx, dy);
    rdx = dx / mag;
    rdy = dy / mag;
  }
  return {rdx, rdy, mag};
}
Z2D_D double slope_normalize(Slope& s) {
  const NormSlope r = slope_normalized(s.dx, s.dy);
  s.dx = r.dx;
  s.dy = r.dy;
  return r.mag;
}

Z2D_D int sgn(double v) { return (v > 0.0) - (v < 0.0); }

Z2D_DN int slope_compare(Slope a, Slope b) {  // Slope.zig:46-85
  const double eps = 2.220446049250313e-16;
  const double bdy = fabs(b.dy - a.dy) > eps ? b.dy : a.dy;
  const double bdx = fabs(b.dx - a.dx) > eps ? b.dx : a.dx;
  const int cmp = sgn(a.dy * bdx - bdy * a.dx);
  if (cmp != 0) return cmp;
  if (a.dx == 0 && a.dy == 0 && bdx == 0 && bdy == 0) return 0;
  if (a.dx == 0 && a.dy == 0) return 1;
  if (bdx == 0 && bdy == 0) return -1;
  if (sgn(a.dx) != sgn(bdx) || sgn(a.dy) != sgn(bdy)) return (a.dx > 0 || (a.dx == 0 && a.dy > 0)) ? -1 : 1;
  return 0;
}

Z2D_D bool miter_within_limit(Slope in, Slope out, double ml) {  // Slope.zig:150-172
  slope_normalize(in);
  slope_normalize(out);
  const double d = in.dx * out.dx + in.dy * out.dy;
  return 2 <= ml * ml * (1 + d);
}

Z2D_D void xf_dist(const double* m, double& x, double& y) {  // Transformation.userToDeviceDistance
  const double ix = x, iy = y;
  x = m[0] * ix + m[1] * iy;
  y = m[2] * ix + m[3] * iy;
}

struct Face {  // tess/Face.zig
  Pt p0, p1;
  Slope dev, user;
  double half_width;
  Pt p0_cw, p0_ccw, p1_cw, p1_ccw;
};

Z2D_D Face face_make(Pt p0, Pt p1, Slope dev, const StrokeCtx& c) {  // Face.zig:65-115
  Face f;
  const double hw = c.thickness / 2;
  double ox, oy;
  Slope us = dev;
  if (!c.ctm_identity) {
    double dx = dev.dx, dy = dev.dy;
    xf_dist(c.inv, dx, dy);
    us = {dx, dy};
    slope_normalize(us);
    if (c.det_nonneg) {
      ox = -us.dy * hw;
      oy = us.dx * hw;
    } else {
      ox = us.dy * hw;
      oy = -us.dx * hw;
    }
    xf_dist(c.ctm, ox, oy);
  } else {
    ox = -dev.dy * hw;
    oy = dev.dx * hw;
  }
  const double ccx = -ox, ccy = -oy;
  f.p0 = p0;
  f.p1 = p1;
  f.dev = dev;
  f.user = us;
  f.half_width = hw;
  f.p0_cw = {p0.x + ox, p0.y + oy};
  f.p0_ccw = {p0.x + ccx, p0.y + ccy};
  f.p1_cw = {p1.x + ox, p1.y + oy};
  f.p1_ccw = {p1.x + ccx, p1.y + ccy};
  return f;
}
Z2D_D Face face_init(Pt p0, Pt p1, const StrokeCtx& c) {
  Slope s{p1.x - p0.x, p1.y - p0.y};
  slope_normalize(s);
  return face_make(p0, p1, s, c);
}
Z2D_D Pt face_intersect(const Face& in, const Face& out, bool clockwise) {  // Face.zig:117-152
  const Pt ip = clockwise ? in.p1_ccw : in.p1_cw;
  const Pt op = clockwise ? out.p0_ccw : out.p0_cw;
  Slope is = in.dev, os = out.dev;
  slope_normalize(is);
  slope_normalize(os);
  const double ry = ((op.x - ip.x) * is.dy * os.dy - op.y * os.dx * is.dy + ip.y * is.dx * os.dy) / (is.dx * os.dy - os.dx * is.dy);
  const double rx = (fabs(is.dy) >= fabs(os.dy)) ? (ry - ip.y) * is.dx / is.dy + ip.x : (ry - op.y) * os.dx / os.dy + op.x;
  return {rx, ry};
}

// Pen.vertexIteratorFor (Pen.zig:138-232).  The reference has one copy of the two binary searches per direction; here the
// direction only selects the operands and the sense of each Slope.compare (same calls, same operand order), so the lanes of a
// warp that search clockwise and counter-clockwise run one loop together.
Z2D_DN int2 pen_range_of(const PenV* __restrict__ v, int n, Slope from, Slope to, bool clockwise) {
  auto cw = [&](int i) Z2D_LAMBDA { return Slope{v[i].cwx, v[i].cwy}; };
  auto ccw = [&](int i) Z2D_LAMBDA { return Slope{v[i].ccwx, v[i].ccwy}; };
  // clockwise: compare(cw(i), from) < 0        counter-clockwise: compare(from, ccw(i)) < 0
  auto before_from = [&](int i) Z2D_LAMBDA { return slope_compare(clockwise ? cw(i) : from, clockwise ? from : ccw(i)) < 0; };
  int low = 0, high = n, i = (low + high) >> 1;
  while (high - low > 1) {
    if (before_from(i)) low = i; else high = i;
    i = (low + high) >> 1;
  }
  if (before_from(i)) {
    i += 1;
    if (i == n) i = 0;
  }
  const int start = i;
  // clockwise: compare(to, ccw(i)) >= 0        counter-clockwise: compare(cw(i), to) <= 0
  const int r0 = slope_compare(clockwise ? to : cw(i), clockwise ? ccw(i) : to);
  if (clockwise ? r0 >= 0 : r0 <= 0) {
    low = i;
    high = i + n;
    i = (low + high) >> 1;
    while (high - low > 1) {
      const int j = i >= n ? i - n : i;
      // clockwise: compare(cw(j), to) > 0      counter-clockwise: compare(to, ccw(j)) > 0
      if (slope_compare(clockwise ? cw(j) : to, clockwise ? to : ccw(j)) > 0) high = i; else low = i;
      i = (low + high) >> 1;
    }
    if (i >= n) i -= n;
  }
  return make_int2(max(0, start), max(0, i));
}
Z2D_D void pen_range(const StrokeCtx& c, Slope from, Slope to, bool clockwise, int& start_o, int& end_o) {
  const int2 r = pen_range_of(c.pen, c.npen, from, to, clockwise);
  start_o = r.x;
  end_o = r.y;
}

struct PointBuf25 {  // PointBuffer(2, 5) (point_buffer.zig)
  Pt items[5];
  int len = 0;
  Z2D_D void add(Pt p) {
    if (len < 5) {
      items[len++] = p;
    } else {
      items[2] = items[3];
      items[3] = items[4];
      items[4] = p;
    }
  }
  Z2D_D void reset() { len = 0; }
  Z2D_D Pt head(int n) const { return items[n]; }
  Z2D_D Pt tail(int n) const { return items[len - n]; }
  Z2D_D Pt first() const { return items[0]; }
  Z2D_D Pt last() const { return items[len - 1]; }
};

// A contour in list order, reduced to what edge generation needs.
struct Contour {
  Pt first, last;
  uint32_t len = 0;
};

struct PlotState {  // the part of the plotter the generic helpers use (Plotter and InitialPolygon)
  Contour outer, inner;
  int clockwise = -1;  // ?bool
};

// The points one line_to / curve_to node feeds to the plotter, one per call, so that a plotter has ONE line_to site for lines,
// curves and the closing segment: a line yields its end point; a curve yields what Spline.decompose (Spline.zig:37-71) passes to
// its callback, in the same order and computed by the same operations as spline_decompose (kernels.cu).
struct SegIter {
  Knots stack[kSplineStack];
  Knots k;
  Pt a, d;
  double tol_sq;
  int sp, phase;  // phase 0: subdividing, 1: end point pending, 2: done
  Z2D_D void line(Pt p) {
    d = p;
    phase = 1;
  }
  Z2D_D void curve(Pt a_, Pt b, Pt c, Pt d_, double tsq) {
    a = a_;
    d = d_;
    tol_sq = tsq;
    sp = 0;
    k = Knots{a_, b, c, d_};
    phase = (pt_eq(a_, b) && pt_eq(c, d_)) ? 1 : 0;  // Spline.zig:39-42
  }
  Z2D_D bool next(Pt& p) {
#pragma unroll 1
    while (phase == 0) {
      while (!(knots_error_sq(k) < tol_sq || sp >= kSplineStack - 2)) stack[sp++] = knots_split(k);  // k becomes the left half
      const Pt ka = k.a;
      if (sp == 0) phase = 1; else k = stack[--sp];
      if (!pt_eq(ka, a)) {
        p = ka;
        return true;
      }
    }
    if (phase == 1) {
      p = d;
      phase = 2;
      return true;
    }
    return false;
  }
};

struct Dasher {  // tess/Dasher.zig
  const double* d;
  int n;
  double offset;
  int idx;
  bool on;
  double remain;
  Z2D_D void reset() {
    idx = 0;
    on = true;
    remain = d[0];
    remain -= offset;
    while (remain < 0 || remain > d[idx]) {
      if (remain < 0) {
        remain += d[idx];
        idx = (idx >= n - 1) ? 0 : idx + 1;
      } else {
        remain -= d[idx];
        idx = (idx == 0) ? n - 1 : idx - 1;
      }
      on = !on;
    }
  }
  Z2D_D bool step(double len) {
    remain -= len;
    if (remain <= 0) {
      on = !on;
      idx += 1;
      if (idx >= n) idx = 0;
      remain = d[idx];
      return true;
    }
    return false;
  }
};

// ---- caps (Face.zig:154-284); `emit(p)` receives the cap points in order.  Written as one loop over "fixed head points, pen
// vertices, fixed tail point" so that `emit` (which ends in the edge sink) is instantiated once per cap, not once per point.
template <class F>
Z2D_D void stroke_cap(const StrokeCtx& c, const Face& f, bool clockwise, F&& emit) {
  Pt h0, h1, h2{}, h3{};
  int n_head, idx = 0, end = 0;
  bool tail = false;
  switch (c.cap) {
    case Z2D_CAP_BUTT:
      n_head = 2;
      h0 = clockwise ? f.p1_ccw : f.p1_cw;
      h1 = clockwise ? f.p1_cw : f.p1_ccw;
      break;
    case Z2D_CAP_SQUARE: {
      double ox = f.user.dx * f.half_width, oy = f.user.dy * f.half_width;
      xf_dist(c.ctm, ox, oy);
      n_head = 4;
      h0 = clockwise ? f.p1_ccw : f.p1_cw;
      h3 = clockwise ? f.p1_cw : f.p1_ccw;
      h1 = {h0.x + ox, h0.y + oy};
      h2 = {h3.x + ox, h3.y + oy};
      break;
    }
    default:
      n_head = 1;
      h0 = clockwise ? f.p1_ccw : f.p1_cw;
      h1 = clockwise ? f.p1_cw : f.p1_ccw;  // the tail point
      tail = true;
      pen_range(c, f.dev, Slope{-f.dev.dx, -f.dev.dy}, clockwise, idx, end);
  }
  #pragma unroll 1
  for (int k = 0;;) {
    Pt p;
    if (k < n_head) {
      p = k == 0 ? h0 : k == 1 ? h1 : k == 2 ? h2 : h3;
      k++;
    } else if (idx != end) {  // VertexIterator.next
      const PenV v = c.pen[idx];
      if (clockwise) {
        idx += 1;
        if (idx == c.npen) idx = 0;
      } else {
        if (idx == 0) idx = c.npen;
        idx -= 1;
      }
      p = {f.p1.x + v.px, f.p1.y + v.py};
    } else if (tail) {
      p = h1;
      tail = false;
    } else {
      break;
    }
    emit(p);
  }
}

template <class Sink>
struct Stroker {
  Sink& sink;
  const StrokeCtx& c;
  Z2D_D Stroker(Sink& s, const StrokeCtx& ctx) : sink(s), c(ctx) {}

  Z2D_D Pt scaled(Pt p) const { return {p.x * c.scale, p.y * c.scale}; }  // Contour.plot pre-scales (Polygon.zig:388-391)

  // Contour.plot(point, null): append
  Z2D_D void append(Contour& ct, Pt p) {
    const Pt s = scaled(p);
    if (ct.len == 0) ct.first = s; else sink.add(ct.last, s);
    ct.last = s;
    ct.len++;
  }
  // Contour.plotReverse: prepend
  Z2D_D void prepend(Contour& ct, Pt p) {
    const Pt s = scaled(p);
    if (ct.len == 0) ct.last = s; else sink.add(s, ct.first);
    ct.first = s;
    ct.len++;
  }
  // a run of Contour.plot(point, before = the node that was first when the run began)
  struct Block {
    bool active = false, any = false;
    Pt first, prev;
  };
  Z2D_D void block_push(Contour& ct, Block& b, Pt p) {
    const Pt s = scaled(p);
    if (!b.any) b.first = s; else sink.add(b.prev, s);
    b.prev = s;
    b.any = true;
    ct.len++;
  }
  Z2D_D void block_end(Contour& ct, Block& b) {
    if (b.any) {
      sink.add(b.prev, ct.first);
      ct.first = b.first;
    }
    b.active = b.any = false;
  }
  Z2D_D void concat(Contour& a, Contour& b) {  // a.concat(&b)
    if (b.len == 0) return;
    if (a.len == 0) {
      a = b;
    } else {
      sink.add(a.last, b.first);
      a.last = b.last;
      a.len += b.len;
    }
    b.len = 0;
  }
  Z2D_D void close_contour(Contour& ct) {  // Polygon.addEdgesFromContour closing edge
    if (ct.len > 0) sink.add(ct.last, ct.first);
    ct.len = 0;
  }

  template <class F>
  Z2D_D void cap(const Face& f, bool clockwise, F&& emit) {
    stroke_cap(c, f, clockwise, emit);
  }

  // ---- stroke_plotter.join (stroke_plotter.zig:410-561); use_before: outer points are inserted
  // before the outer contour's first node (dashed_plotter.zig:603-611) instead of appended
  Z2D_D void join(PlotState& st, uint32_t join_mode, Pt p0, Pt p1, Pt p2, bool use_before) {
    if (pt_eq(p0, p1) || pt_eq(p1, p2)) {
      if (st.clockwise < 0) st.clockwise = 0;
      return;
    }
    const Face in = face_init(p0, p1, c), out = face_init(p1, p2, c);
    const int cmp = slope_compare(in.dev, out.dev);
    const bool join_cw = cmp < 0;
    const bool poly_cw = st.clockwise >= 0 ? (st.clockwise != 0) : join_cw;
    const bool switched = join_cw != poly_cw;
    Block blk;
    blk.active = use_before && st.outer.len > 0;
    // The join's points in emission order: outer side = o0, pen vertices [idx, end), o1; inner side = i0, p1, i1 (only i0 when
    // the faces are parallel).  One loop with one plotting site, so the edge sink is instantiated once per join.
    Pt o0, o1{};
    int n_o0 = 1, idx = 0, end = 0;
    bool has_o1 = false;
    int n_inner = 3;
    if (cmp == 0) {
      o0 = join_cw ? in.p1_ccw : in.p1_cw;
      n_inner = 1;
    } else if (join_mode == Z2D_JOIN_ROUND) {
      o0 = join_cw ? in.p1_ccw : in.p1_cw;
      pen_range(c, in.dev, out.dev, join_cw, idx, end);
      o1 = join_cw ? out.p0_ccw : out.p0_cw;
      has_o1 = true;
    } else if (join_mode == Z2D_JOIN_MITER && miter_within_limit(in.dev, out.dev, c.miter_limit)) {
      o0 = face_intersect(in, out, join_cw);
    } else {
      o0 = join_cw ? in.p1_ccw : in.p1_cw;
      o1 = join_cw ? out.p0_ccw : out.p0_cw;
      has_o1 = true;
    }
    const Pt i0 = join_cw ? in.p1_cw : in.p1_ccw, i1 = join_cw ? out.p0_cw : out.p0_ccw;
    #pragma unroll 1
    for (int ki = 0;;) {
      Pt p;
      bool outer_side = true;
      if (n_o0) {
        p = o0;
        n_o0 = 0;
      } else if (idx != end) {
        const PenV v = c.pen[idx];
        if (join_cw) {
          idx += 1;
          if (idx == c.npen) idx = 0;
        } else {
          if (idx == 0) idx = c.npen;
          idx -= 1;
        }
        p = {p1.x + v.px, p1.y + v.py};
      } else if (has_o1) {
        p = o1;
        has_o1 = false;
      } else if (ki < n_inner) {
        p = ki == 0 ? i0 : ki == 1 ? p1 : i1;
        outer_side = false;
        ki++;
      } else {
        break;
      }
      // outer-side points go to the outer contour (appended, or inserted before its first node) unless the join turns against
      // the polygon's direction, in which case the sides swap; inner-side points are prepended to the inner contour
      const Pt sp = scaled(p);
      Pt a, b;
      bool have;
      if (outer_side == switched) {  // prepend(st.inner, p)
        have = st.inner.len != 0;
        a = sp;
        b = st.inner.first;
        if (!have) st.inner.last = sp;
        st.inner.first = sp;
        st.inner.len++;
      } else if (blk.active) {  // block_push(st.outer, blk, p)
        have = blk.any;
        a = blk.prev;
        b = sp;
        if (!have) blk.first = sp;
        blk.prev = sp;
        blk.any = true;
        st.outer.len++;
      } else {  // append(st.outer, p)
        have = st.outer.len != 0;
        a = st.outer.last;
        b = sp;
        if (!have) st.outer.first = sp;
        st.outer.last = sp;
        st.outer.len++;
      }
      if (have) sink.add(a, b);
    }
    if (blk.active) block_end(st.outer, blk);
    if (st.clockwise < 0) st.clockwise = poly_cw ? 1 : 0;
  }

  Z2D_D void plot_single(PlotState& st, Pt start, Pt end) {  // stroke_plotter.zig:251-294
    const Face f = face_init(start, end, c);
    const Face rev = face_init(end, start, c);  // cap_p0 caps the reversed face
    cap(rev, true, [&](Pt p) Z2D_LAMBDA { append(st.outer, p); });
    cap(f, true, [&](Pt p) Z2D_LAMBDA { append(st.outer, p); });
    close_contour(st.outer);
    st.clockwise = -1;
  }

  Z2D_D void plot_open_joined(PlotState& st, Pt start0, Pt end0, Pt start1, Pt end1) {  // stroke_plotter.zig:296-364
    const Face fs_rev = face_init(end0, start0, c);
    const Face fe = face_init(start1, end1, c);
    const bool cw = st.clockwise >= 0 ? (st.clockwise != 0) : true;
    if (st.outer.len == 0) {
      cap(fs_rev, cw, [&](Pt p) Z2D_LAMBDA { append(st.outer, p); });
    } else {
      Block blk;
      blk.active = true;
      cap(fs_rev, cw, [&](Pt p) Z2D_LAMBDA { block_push(st.outer, blk, p); });
      block_end(st.outer, blk);
    }
    cap(fe, cw, [&](Pt p) Z2D_LAMBDA { append(st.outer, p); });
    concat(st.outer, st.inner);
    close_contour(st.outer);
    st.inner.len = 0;
    st.clockwise = -1;
  }

  Z2D_D void plot_closed_joined(PlotState& st, Pt initial0, Pt initial1, Pt p1, Pt p2) {  // stroke_plotter.zig:366-408
    if (!pt_eq(p2, initial0)) {
      join(st, c.join, p1, p2, initial0, false);
      join(st, c.join, p2, initial0, initial1, false);
    } else {
      join(st, c.join, p1, initial0, initial1, false);
    }
    close_contour(st.outer);
    close_contour(st.inner);
    st.clockwise = -1;
  }

  Z2D_D void pen_circle(PlotState& st, Pt point) {
    #pragma unroll 1
    for (int i = 0; i < c.npen; i++) append(st.outer, {point.x + c.pen[i].px, point.y + c.pen[i].py});
    close_contour(st.outer);
  }

  // iterative Spline.decompose (Spline.zig:37-71)
  template <class F>
  Z2D_D void spline(Pt a, Pt b, Pt cc, Pt d, F&& line_to) {
    spline_decompose(a, b, cc, d, c.tolerance * c.tolerance, line_to);
  }

  // =============================== undashed (stroke_plotter.zig:77-249)
  Z2D_D void run_plain(const z2d_node* __restrict__ nodes, uint32_t begin, uint32_t end) {
    PlotState st;
    PointBuf25 pts;
    auto line_to = [&](uint32_t join_mode, Pt p) Z2D_LAMBDA {
      if (pts.len == 0 || pt_eq(p, pts.last())) return;
      pts.add(p);
      if (pts.len > 2) join(st, join_mode, pts.tail(3), pts.tail(2), pts.tail(1), false);
    };
    auto finish = [&]() Z2D_LAMBDA {
      if (pts.len == 2) plot_single(st, pts.head(0), pts.head(1));
      else if (pts.len > 2) plot_open_joined(st, pts.head(0), pts.head(1), pts.tail(2), pts.tail(1));
    };
    SegIter it;
#pragma unroll 1
    for (uint32_t i = begin; i <= end; i++) {  // one extra trip: the end of the node list finishes like a move_to
      const bool at_end = i == end;
      const z2d_node nd = nodes[at_end ? begin : i];
      const uint32_t tag = at_end ? (uint32_t)Z2D_NODE_MOVE_TO : nd.tag;
      if (tag == Z2D_NODE_MOVE_TO) {
        finish();
        if (at_end) break;
        pts.reset();
        pts.add({nd.p[0], nd.p[1]});
      } else if (tag == Z2D_NODE_LINE_TO || tag == Z2D_NODE_CURVE_TO) {
        if (pts.len == 0) continue;
        uint32_t jm = c.join;
        if (tag == Z2D_NODE_LINE_TO) {
          it.line({nd.p[0], nd.p[1]});
        } else {
          it.curve(pts.last(), {nd.p[0], nd.p[1]}, {nd.p[2], nd.p[3]}, {nd.p[4], nd.p[5]}, c.tolerance * c.tolerance);
          jm = Z2D_JOIN_ROUND;
        }
        Pt p;
#pragma unroll 1
        while (it.next(p)) line_to(jm, p);
      } else {  // close_path (stroke_plotter.zig:157-180)
        if (pts.len == 1) {
          if (c.cap == Z2D_CAP_ROUND) {  // plotDotted (202-237)
            pen_circle(st, pts.first());
            st.clockwise = -1;
          }
        } else if (pts.len == 2) {
          plot_single(st, pts.head(0), pts.head(1));
        } else if (pts.len > 2) {
          plot_closed_joined(st, pts.head(0), pts.head(1), pts.tail(2), pts.tail(1));
        }
        pts.reset();
      }
    }
  }

  // =============================== dashed (dashed_plotter.zig)
  Z2D_D void plot_dotted_dashed(PlotState& st, Pt point, Slope slope) {  // dashed_plotter.zig:369-465 (always on the main state)
    if (c.cap == Z2D_CAP_ROUND) {
      pen_circle(st, point);
    } else if (c.cap == Z2D_CAP_SQUARE) {
      const Face f = face_make(point, point, slope, c);
      double ox = f.user.dx * f.half_width, oy = f.user.dy * f.half_width;
      xf_dist(c.ctm, ox, oy);
      append(st.outer, {f.p1_cw.x - ox, f.p1_cw.y - oy});
      append(st.outer, {f.p1_cw.x + ox, f.p1_cw.y + oy});
      append(st.outer, {f.p1_ccw.x + ox, f.p1_ccw.y + oy});
      append(st.outer, {f.p1_ccw.x - ox, f.p1_ccw.y - oy});
      close_contour(st.outer);
    }
    st.outer.len = 0;  // contour reset (also discards pending outer points, as the reference does)
    st.clockwise = -1;
  }

  Z2D_D void run_dashed(const z2d_node* __restrict__ nodes, uint32_t begin, uint32_t end) {
    PlotState st, ist;     // main and "initial polygon" states
    PointBuf25 pts, ipts;  // ...and their point buffers
    Slope cur_slope{0, 0}, islope{0, 0};
    int initial_kind = 0;  // 0 none, 1 off, 2 on
    Pt initial_off{0, 0};
    Dasher dasher{c.dashes, c.ndash, c.dash_offset, 0, true, 0.0};
    dasher.reset();

    auto emit_current = [&]() Z2D_LAMBDA {
      if (pts.len == 1) plot_dotted_dashed(st, pts.first(), cur_slope);
      else if (pts.len == 2) plot_single(st, pts.head(0), pts.head(1));
      else if (pts.len > 2) plot_open_joined(st, pts.head(0), pts.head(1), pts.tail(2), pts.tail(1));
    };
    auto save_initial = [&]() Z2D_LAMBDA {  // 467-520
      if (!dasher.on) {
        initial_kind = 2;
        ist = st;
        ipts = pts;
        islope = cur_slope;
      } else {
        initial_kind = 1;
        initial_off = pts.first();
      }
      st.outer.len = 0;
      st.inner.len = 0;
      st.clockwise = -1;
    };
    auto seg_mark = sink.mark();  // sink position where the dash in progress began
    auto next_segment = [&](Pt point) Z2D_LAMBDA {  // 307-332
      if (initial_kind == 0) save_initial();
      else if (!dasher.on) emit_current();
      pts.reset();
      pts.add(point);
      seg_mark = sink.mark();
    };
    auto finish_initial = [&]() Z2D_LAMBDA {  // finishInitialDotted / finishInitial (522-552)
      if (ipts.len == 1) {
        plot_dotted_dashed(st, ipts.first(), islope);
      } else if (ipts.len >= 2) {
        plot_open_joined(ist, ipts.head(0), ipts.head(1), ipts.tail(2), ipts.tail(1));
      }
      initial_kind = 0;
    };
    auto line_to = [&](uint32_t join_mode, Pt target) Z2D_LAMBDA {  // _runLineTo (123-173)
      if (pts.len == 0) return;
      const Pt current = pts.last();
      if (pt_eq(target, current)) return;
      const Pt first_dash_point = current;
      Slope slope{target.x - first_dash_point.x, target.y - first_dash_point.y};
      cur_slope = slope;
      slope_normalize(cur_slope);
      xf_dist(c.inv, slope.dx, slope.dy);
      const double total_len = slope_normalize(slope);
      double remaining = total_len;
      double step_len = fmin(dasher.remain, remaining);
      #pragma unroll 1
      while (remaining > 0) {
        remaining -= step_len;
        double xo = slope.dx * (total_len - remaining), yo = slope.dy * (total_len - remaining);
        xf_dist(c.ctm, xo, yo);
        const Pt dp{first_dash_point.x + xo, first_dash_point.y + yo};
        if (!pt_eq(dp, pts.last())) pts.add(dp);
        if (dasher.on && pts.len > 2) join(st, join_mode, pts.tail(3), pts.tail(2), pts.tail(1), false);
        if (dasher.step(step_len)) next_segment(dp);
        step_len = fmin(dasher.remain, remaining);
      }
    };
    auto finish = [&]() Z2D_LAMBDA {  // 334-367
      if (initial_kind == 2) {
        if (ipts.len >= 1) finish_initial();
      } else if (initial_kind == 1) {
        initial_kind = 0;
      }
      if (dasher.on) emit_current();
    };
    auto join_and_cap_initial = [&]() Z2D_LAMBDA {  // 554-628
      if (pts.len > 2) {
        join(st, c.join, pts.tail(2), ipts.head(0), ipts.head(1), false);
        concat(st.outer, ist.outer);  // self.outer.concat(&initial.outer)
        concat(ist.inner, st.inner);  // initial.inner.concat(&self.inner)
        ist.outer = st.outer;         // initial.outer = self.outer
        plot_open_joined(ist, pts.head(0), pts.head(1), ipts.tail(2), ipts.tail(1));
      } else {
        join(ist, c.join, pts.tail(2), ipts.head(0), ipts.head(1), true);
        plot_open_joined(ist, pts.first(), ipts.first(), ipts.tail(2), ipts.tail(1));
      }
      initial_kind = 0;
      st.outer.len = 0;
      st.inner.len = 0;
      st.clockwise = -1;
    };

    SegIter it;
#pragma unroll 1
    for (uint32_t i = begin; i <= end; i++) {  // one extra trip: the end of the node list finishes like a move_to
      const bool at_end = i == end;
      const z2d_node nd = nodes[at_end ? begin : i];
      const uint32_t tag = at_end ? (uint32_t)Z2D_NODE_MOVE_TO : nd.tag;
      if (tag == Z2D_NODE_MOVE_TO) {
        finish();
        if (at_end) break;
        dasher.reset();
        pts.reset();
        pts.add({nd.p[0], nd.p[1]});
        seg_mark = sink.mark();
        continue;
      }
      if (pts.len == 0) continue;  // line_to / curve_to / close_path without a current point
      uint32_t jm = c.join;
      if (tag == Z2D_NODE_LINE_TO) {
        it.line({nd.p[0], nd.p[1]});
      } else if (tag == Z2D_NODE_CURVE_TO) {
        it.curve(pts.last(), {nd.p[0], nd.p[1]}, {nd.p[2], nd.p[3]}, {nd.p[4], nd.p[5]}, c.tolerance * c.tolerance);
        jm = Z2D_JOIN_ROUND;
      } else {  // close_path (202-305): first a line to where the sub-path began
        it.line(initial_kind == 2 ? ipts.first() : (initial_kind == 1 ? initial_off : pts.first()));
      }
      Pt p;
#pragma unroll 1
      while (it.next(p)) line_to(jm, p);
      if (tag != Z2D_NODE_LINE_TO && tag != Z2D_NODE_CURVE_TO) {
        if (initial_kind == 2) {
          if (dasher.on && pts.len > 1) {
            if (ipts.len == 1) {
              plot_open_joined(st, pts.head(0), pts.head(1), pts.tail(2), pts.tail(1));
              initial_kind = 0;
            } else {
              join_and_cap_initial();
            }
          } else {
            finish_initial();
          }
        } else if (initial_kind == 1) {
          // "we've already drawn back to the initial point" (dashed_plotter.zig:258-262): the reference resets the point buffer
          // here, so a dash still in progress is never capped and its buffered join points never become edges
          initial_kind = 0;
          sink.rewind(seg_mark);
          st.outer.len = 0;
          st.inner.len = 0;
          st.clockwise = -1;
        } else {
          if (pts.len == 1) {
            plot_dotted_dashed(st, pts.first(), cur_slope);
          } else if (pts.len == 2) {
            plot_single(st, pts.head(0), pts.head(1));
          } else {
            join(st, c.join, pts.tail(2), pts.head(0), pts.head(1), false);
            close_contour(st.outer);
            close_contour(st.inner);
            st.clockwise = -1;
          }
        }
        pts.reset();
      }
    }
  }
};

#ifndef Z2D_HOST_TEST
Z2D_D void stroke_ctx_of(StrokeCtx& c, const DevDraw& d, const PenV* pens, const double* dashes) {
  c.cap = d.cap;
  c.join = d.join;
  c.thickness = d.thickness;
  c.miter_limit = d.miter_limit;
  c.scale = d.scale;
  c.tolerance = d.tolerance;
  c.dash_offset = d.dash_offset;
  for (int i = 0; i < 6; i++) {
    c.ctm[i] = d.ctm[i];
    c.inv[i] = d.inv[i];
  }
  c.ctm_identity = d.ctm[0] == 1 && d.ctm[1] == 0 && d.ctm[2] == 0 && d.ctm[3] == 1 && d.ctm[4] == 0 && d.ctm[5] == 0;
  c.det_nonneg = (d.ctm[0] * d.ctm[3] - d.ctm[1] * d.ctm[2]) >= 0;
  c.pen = pens + d.pen_begin;
  c.npen = (int)d.pen_count;
  c.dashes = dashes + d.dash_begin;
  c.ndash = (int)d.dash_count;
}

template <bool EMIT>
Z2D_D void stroke_subpath(const z2d_node* __restrict__ nodes, uint32_t begin, uint32_t end, const DevDraw& d, const PenV* pens,
                          const double* dashes, EdgeSink<EMIT>& sink) {
  StrokeCtx c;
  stroke_ctx_of(c, d, pens, dashes);
  sink.scale = 1.0;  // contour points are pre-scaled; the result polygon's own scale is 1
  Stroker<EdgeSink<EMIT>> s(sink, c);
  if (d.dash_count > 0) s.run_dashed(nodes, begin, end); else s.run_plain(nodes, begin, end);
}
#endif

}  // namespace z2d
