// Stroke tessellation in parallel over UNITS (joins, caps, dots) instead of over sub-paths.
//
// The sub-path stroker (stroke.cuh) is one thread per sub-path: config 3 (50 000 strokes, 360 edges each) runs 1 563 warps with
// 255 registers and ~8 active lanes, twice (count + emit) -- latency bound at 4 ms per pass.  Almost all of that work is
// independent per join / cap: a join's points depend only on its three polyline points, the join mode and the polygon's
// direction latch; a cap's only on its segment.  What IS sequential -- curve subdivision order, the duplicate-point filter, the
// dash phase (Dasher.zig: non-associative f64 bookkeeping), the direction latch (stroke_plotter.zig:420-426), and which
// contour every point list is appended / prepended / spliced into (stroke_plotter.zig:296-408, dashed_plotter.zig:307-628) --
// is cheap.  So the work is split in three:
//
//   k_stroke_walk   one thread per sub-path: the plotter's control flow WITHOUT its geometry.  Where the plotter computes a
//                   join or a cap and plots the points, the walker records a 64-byte unit {kind, p0, p1, p2, flags}; where
//                   the plotter connects two point lists (append after the contour's last point, prepend before its first,
//                   splice before the first node, concat, close) the walker records a link between two PORTS -- first / last
//                   point of a unit's appended list (A) or prepended list (B) -- because only consecutive pairs of contour
//                   points ever become edges.
//   k_stroke_units  one thread per unit: Face / Slope / Pen arithmetic (the same functions as stroke.cuh), the unit's internal
//                   edges, and its four port points.
//   k_stroke_links  one thread per link: the edge between two stored ports.
//
// Edges go to a pool with a warp-aggregated cursor; their order is irrelevant downstream (winding is a sum), each carries its
// draw index, horizontal edges (dropped by Polygon.addEdge) leave a dead slot (edge_draw = ~0).  All three arrays are sized
// optimistically by the host and the batch is redone with larger ones if a cursor ran past its capacity.
//
// tools/stroke_units_host_test.cpp compiles this file and stroke.cuh for the host and checks that both strokers produce the same
// multiset of edges for every sub-path of a large random corpus.
#pragma once

namespace z2d {

constexpr uint32_t kUnitDead = 0u, kUnitJoin = 1u, kUnitCap = 2u, kUnitDotRound = 3u, kUnitDotSquare = 4u;
constexpr uint32_t kUnitKindMask = 7u;
constexpr uint32_t kUnitCw = 8u;         // join: the polygon's direction latch; cap: the `clockwise` argument
constexpr uint32_t kUnitJoinShift = 4u;  // join mode (Z2D_JOIN_*)
constexpr uint32_t kPortAF = 0u, kPortAL = 1u, kPortBF = 2u, kPortBL = 3u;  // first / last point of the appended (A) / prepended (B) list
constexpr uint32_t kNoUnit = 0xffffffffu;

struct StrokeUnit {  // 64 bytes
  uint32_t kind;     // kUnit* | kUnitCw | join mode << kUnitJoinShift
  uint32_t draw;
  uint32_t prev;     // previous unit recorded by the same walker thread (rewind chain)
  uint32_t _pad;
  double p[6];       // join: p0, p1, p2; cap: the face's p0, p1; dots: centre, (square) slope
};
struct StrokeLink {  // edge from port `from` to port `to` (port = unit * 4 + kPort*)
  uint32_t from, to, draw, prev;
};

// -------------------------------------------------------------------------------------------------- the walker
struct WContour {  // a contour in list order, reduced to its two end ports
  uint32_t first = 0, last = 0;
  bool has = false;
};
struct WState {
  WContour outer, inner;
  int clockwise = -1;  // ?bool
};

template <class Rec>
struct StrokeWalker {
  Rec& rec;
  const StrokeCtx& c;
  Z2D_D StrokeWalker(Rec& r, const StrokeCtx& ctx) : rec(r), c(ctx) {}

  // ---- contour operations (stroke.cuh Stroker::append / prepend / block_* / concat / close_contour on whole point lists)
  Z2D_D void append_a(WContour& ct, uint32_t u) {
    if (ct.has) rec.link(ct.last, u * 4 + kPortAF); else ct.first = u * 4 + kPortAF;
    ct.last = u * 4 + kPortAL;
    ct.has = true;
  }
  Z2D_D void insert_a(WContour& ct, uint32_t u) {  // before the contour's first node (ct.has)
    rec.link(u * 4 + kPortAL, ct.first);
    ct.first = u * 4 + kPortAF;
  }
  Z2D_D void prepend_b(WContour& ct, uint32_t u) {
    if (ct.has) rec.link(u * 4 + kPortBF, ct.first); else ct.last = u * 4 + kPortBF;
    ct.first = u * 4 + kPortBL;
    ct.has = true;
  }
  Z2D_D void concat(WContour& a, WContour& b) {
    if (!b.has) return;
    if (!a.has) {
      a = b;
    } else {
      rec.link(a.last, b.first);
      a.last = b.last;
    }
    b.has = false;
  }
  Z2D_D void close_contour(WContour& ct) {
    if (ct.has) rec.link(ct.last, ct.first);
    ct.has = false;
  }

  Z2D_D void join(WState& st, uint32_t join_mode, Pt p0, Pt p1, Pt p2, bool use_before) {
    if (pt_eq(p0, p1) || pt_eq(p1, p2)) {
      if (st.clockwise < 0) st.clockwise = 0;
      return;
    }
    bool poly_cw;
    if (st.clockwise >= 0) {
      poly_cw = st.clockwise != 0;
    } else {  // the latch takes the direction of this join (Face.init slopes, Slope.compare)
      Slope a{p1.x - p0.x, p1.y - p0.y}, b{p2.x - p1.x, p2.y - p1.y};
      slope_normalize(a);
      slope_normalize(b);
      poly_cw = slope_compare(a, b) < 0;
    }
    const uint32_t u = rec.unit(kUnitJoin | (poly_cw ? kUnitCw : 0u) | (join_mode << kUnitJoinShift), p0, p1, p2);
    if (use_before && st.outer.has) insert_a(st.outer, u); else append_a(st.outer, u);
    prepend_b(st.inner, u);
    if (st.clockwise < 0) st.clockwise = poly_cw ? 1 : 0;
  }
  Z2D_D uint32_t cap_unit(Pt p0, Pt p1, bool clockwise) { return rec.unit(kUnitCap | (clockwise ? kUnitCw : 0u), p0, p1, Pt{0, 0}); }

  Z2D_D void plot_single(WState& st, Pt start, Pt end) {
    append_a(st.outer, cap_unit(end, start, true));
    append_a(st.outer, cap_unit(start, end, true));
    close_contour(st.outer);
    st.clockwise = -1;
  }
  Z2D_D void plot_open_joined(WState& st, Pt start0, Pt end0, Pt start1, Pt end1) {
    const bool cw = st.clockwise >= 0 ? (st.clockwise != 0) : true;
    const uint32_t us = cap_unit(end0, start0, cw);
    if (!st.outer.has) append_a(st.outer, us); else insert_a(st.outer, us);
    append_a(st.outer, cap_unit(start1, end1, cw));
    concat(st.outer, st.inner);
    close_contour(st.outer);
    st.inner.has = false;
    st.clockwise = -1;
  }
  Z2D_D void plot_closed_joined(WState& st, Pt initial0, Pt initial1, Pt p1, Pt p2) {
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
  Z2D_D void pen_circle(WState& st, Pt point) {
    if (c.npen > 0) append_a(st.outer, rec.unit(kUnitDotRound, point, Pt{0, 0}, Pt{0, 0}));
    close_contour(st.outer);
  }

  // =============================== undashed (stroke_plotter.zig:77-249; stroke.cuh run_plain)
  Z2D_D void run_plain(const z2d_node* __restrict__ nodes, uint32_t begin, uint32_t end) {
    WState st;
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
    for (uint32_t i = begin; i <= end; i++) {
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
      } else {  // close_path
        if (pts.len == 1) {
          if (c.cap == Z2D_CAP_ROUND) {
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

  // =============================== dashed (dashed_plotter.zig; stroke.cuh run_dashed)
  Z2D_D void plot_dotted_dashed(WState& st, Pt point, Slope slope) {
    if (c.cap == Z2D_CAP_ROUND) {
      pen_circle(st, point);
    } else if (c.cap == Z2D_CAP_SQUARE) {
      append_a(st.outer, rec.unit(kUnitDotSquare, point, Pt{slope.dx, slope.dy}, Pt{0, 0}));
      close_contour(st.outer);
    }
    st.outer.has = false;
    st.clockwise = -1;
  }

  Z2D_D void run_dashed(const z2d_node* __restrict__ nodes, uint32_t begin, uint32_t end) {
    WState st, ist;
    PointBuf25 pts, ipts;
    // current_slope (dashed_plotter.zig:98) is only read by a square dot: kept un-normalised, normalised where it is used
    Slope cur_raw{0, 0}, iraw{0, 0};
    bool cur_set = false, iset = false;
    auto unit_slope = [&](Slope raw, bool set) Z2D_LAMBDA {
      if (set && c.cap == Z2D_CAP_SQUARE) slope_normalize(raw);
      return raw;
    };
    int initial_kind = 0;  // 0 none, 1 off, 2 on
    Pt initial_off{0, 0};
    Dasher dasher{c.dashes, c.ndash, c.dash_offset, 0, true, 0.0};
    dasher.reset();

    auto emit_current = [&]() Z2D_LAMBDA {
      if (pts.len == 1) plot_dotted_dashed(st, pts.first(), unit_slope(cur_raw, cur_set));
      else if (pts.len == 2) plot_single(st, pts.head(0), pts.head(1));
      else if (pts.len > 2) plot_open_joined(st, pts.head(0), pts.head(1), pts.tail(2), pts.tail(1));
    };
    auto save_initial = [&]() Z2D_LAMBDA {
      if (!dasher.on) {
        initial_kind = 2;
        ist = st;
        ipts = pts;
        iraw = cur_raw;
        iset = cur_set;
      } else {
        initial_kind = 1;
        initial_off = pts.first();
      }
      st.outer.has = false;
      st.inner.has = false;
      st.clockwise = -1;
    };
    auto seg_mark = rec.mark();
    auto next_segment = [&](Pt point) Z2D_LAMBDA {
      if (initial_kind == 0) save_initial();
      else if (!dasher.on) emit_current();
      pts.reset();
      pts.add(point);
      seg_mark = rec.mark();
    };
    auto finish_initial = [&]() Z2D_LAMBDA {
      if (ipts.len == 1) {
        plot_dotted_dashed(st, ipts.first(), unit_slope(iraw, iset));
      } else if (ipts.len >= 2) {
        plot_open_joined(ist, ipts.head(0), ipts.head(1), ipts.tail(2), ipts.tail(1));
      }
      initial_kind = 0;
    };
    auto line_to = [&](uint32_t join_mode, Pt target) Z2D_LAMBDA {
      if (pts.len == 0) return;
      const Pt current = pts.last();
      if (pt_eq(target, current)) return;
      const Pt first_dash_point = current;
      Slope slope{target.x - first_dash_point.x, target.y - first_dash_point.y};
      cur_raw = slope;
      cur_set = true;
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
    auto finish = [&]() Z2D_LAMBDA {
      if (initial_kind == 2) {
        if (ipts.len >= 1) finish_initial();
      } else if (initial_kind == 1) {
        initial_kind = 0;
      }
      if (dasher.on) emit_current();
    };
    auto join_and_cap_initial = [&]() Z2D_LAMBDA {
      if (pts.len > 2) {
        join(st, c.join, pts.tail(2), ipts.head(0), ipts.head(1), false);
        concat(st.outer, ist.outer);
        concat(ist.inner, st.inner);
        ist.outer = st.outer;
        plot_open_joined(ist, pts.head(0), pts.head(1), ipts.tail(2), ipts.tail(1));
      } else {
        join(ist, c.join, pts.tail(2), ipts.head(0), ipts.head(1), true);
        plot_open_joined(ist, pts.first(), ipts.first(), ipts.tail(2), ipts.tail(1));
      }
      initial_kind = 0;
      st.outer.has = false;
      st.inner.has = false;
      st.clockwise = -1;
    };

    SegIter it;
#pragma unroll 1
    for (uint32_t i = begin; i <= end; i++) {
      const bool at_end = i == end;
      const z2d_node nd = nodes[at_end ? begin : i];
      const uint32_t tag = at_end ? (uint32_t)Z2D_NODE_MOVE_TO : nd.tag;
      if (tag == Z2D_NODE_MOVE_TO) {
        finish();
        if (at_end) break;
        dasher.reset();
        pts.reset();
        pts.add({nd.p[0], nd.p[1]});
        seg_mark = rec.mark();
        continue;
      }
      if (pts.len == 0) continue;
      uint32_t jm = c.join;
      if (tag == Z2D_NODE_LINE_TO) {
        it.line({nd.p[0], nd.p[1]});
      } else if (tag == Z2D_NODE_CURVE_TO) {
        it.curve(pts.last(), {nd.p[0], nd.p[1]}, {nd.p[2], nd.p[3]}, {nd.p[4], nd.p[5]}, c.tolerance * c.tolerance);
        jm = Z2D_JOIN_ROUND;
      } else {
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
          initial_kind = 0;
          rec.rewind(seg_mark);
          st.outer.has = false;
          st.inner.has = false;
          st.clockwise = -1;
        } else {
          if (pts.len == 1) {
            plot_dotted_dashed(st, pts.first(), unit_slope(cur_raw, cur_set));
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

// -------------------------------------------------------------------------------------------------- one unit
// A unit's points as two lists: the first = fixed head points, a run of pen vertices around `centre`, a fixed tail point
// (join: the outer side -- o0, Pen vertices between the two faces, o1 -- stroke_plotter.zig:470-540; cap: Face.zig:154-284;
// dots: the whole pen / the four corners); the second (joins only) = the inner side i0, p1, i1 (or i0 alone when the faces are
// parallel).  `swap`: the join turns against the polygon's direction, the first list is prepended to the inner contour and the
// second appended to the outer one.
struct UnitPlan {
  Pt h[4];
  Pt centre, t;
  Pt in[3];
  int n_head = 0, n_tail = 0, n_in = 0;
  int run_start = 0, run_len = 0;
  bool run_cw = true, swap = false;
  Z2D_D uint32_t slots() const {  // edges between consecutive points of each list
    const int n1 = n_head + run_len + n_tail;
    return (uint32_t)((n1 > 0 ? n1 - 1 : 0) + (n_in > 0 ? n_in - 1 : 0));
  }
};

Z2D_D int pen_run_len(int idx, int end, int npen, bool cw) {
  int d = cw ? end - idx : idx - end;
  if (d < 0) d += npen;
  return d;
}

Z2D_D void unit_plan(const StrokeCtx& c, uint32_t kind, const double* __restrict__ q, UnitPlan& P) {
  const uint32_t k = kind & kUnitKindMask;
  const bool flag_cw = (kind & kUnitCw) != 0;
  const bool is_join = k == kUnitJoin, is_cap = k == kUnitCap;
  // Joins and caps sit next to each other in the unit array (a dash = cap, cap or join, cap, cap): the face of the first
  // segment and the pen search are done at ONE place for both kinds so that the lanes of a warp run them together.
  const Pt p0{q[0], q[1]}, p1{q[2], q[3]}, p2{q[4], q[5]};
  Face fa{}, fb{};
  if (is_join || is_cap) fa = face_init(p0, p1, c);
  if (is_join) fb = face_init(p1, p2, c);
  bool need_run = false, run_cw = true;
  Slope from{0, 0}, to{0, 0};
  if (is_join) {
    const Face &in = fa, &out = fb;
    const uint32_t join_mode = kind >> kUnitJoinShift;
    const int cmp = slope_compare(in.dev, out.dev);
    const bool join_cw = cmp < 0;
    P.swap = join_cw != flag_cw;
    P.n_head = 1;
    P.n_in = 3;
    if (cmp == 0) {
      P.h[0] = join_cw ? in.p1_ccw : in.p1_cw;
      P.n_in = 1;
    } else if (join_mode == Z2D_JOIN_ROUND) {
      P.h[0] = join_cw ? in.p1_ccw : in.p1_cw;
      need_run = true;
      from = in.dev;
      to = out.dev;
      run_cw = join_cw;
      P.centre = p1;
      P.t = join_cw ? out.p0_ccw : out.p0_cw;
      P.n_tail = 1;
    } else if (join_mode == Z2D_JOIN_MITER && miter_within_limit(in.dev, out.dev, c.miter_limit)) {
      P.h[0] = face_intersect(in, out, join_cw);
    } else {
      P.h[0] = join_cw ? in.p1_ccw : in.p1_cw;
      P.t = join_cw ? out.p0_ccw : out.p0_cw;
      P.n_tail = 1;
    }
    P.in[0] = join_cw ? in.p1_cw : in.p1_ccw;
    P.in[1] = p1;
    P.in[2] = join_cw ? out.p0_cw : out.p0_ccw;
  } else if (is_cap) {
    const Face& f = fa;
    const bool clockwise = flag_cw;
    if (c.cap == Z2D_CAP_BUTT) {
      P.n_head = 2;
      P.h[0] = clockwise ? f.p1_ccw : f.p1_cw;
      P.h[1] = clockwise ? f.p1_cw : f.p1_ccw;
    } else if (c.cap == Z2D_CAP_SQUARE) {
      double ox = f.user.dx * f.half_width, oy = f.user.dy * f.half_width;
      xf_dist(c.ctm, ox, oy);
      P.n_head = 4;
      P.h[0] = clockwise ? f.p1_ccw : f.p1_cw;
      P.h[3] = clockwise ? f.p1_cw : f.p1_ccw;
      P.h[1] = {P.h[0].x + ox, P.h[0].y + oy};
      P.h[2] = {P.h[3].x + ox, P.h[3].y + oy};
    } else {
      P.n_head = 1;
      P.h[0] = clockwise ? f.p1_ccw : f.p1_cw;
      P.t = clockwise ? f.p1_cw : f.p1_ccw;
      P.n_tail = 1;
      need_run = true;
      from = f.dev;
      to = Slope{-f.dev.dx, -f.dev.dy};
      run_cw = clockwise;
      P.centre = f.p1;
    }
  } else if (k == kUnitDotRound) {
    P.centre = p0;
    P.run_start = 0;
    P.run_len = c.npen;
    P.run_cw = true;
  } else if (k == kUnitDotSquare) {
    const Face f = face_make(p0, p0, Slope{q[2], q[3]}, c);
    double ox = f.user.dx * f.half_width, oy = f.user.dy * f.half_width;
    xf_dist(c.ctm, ox, oy);
    P.n_head = 4;
    P.h[0] = {f.p1_cw.x - ox, f.p1_cw.y - oy};
    P.h[1] = {f.p1_cw.x + ox, f.p1_cw.y + oy};
    P.h[2] = {f.p1_ccw.x + ox, f.p1_ccw.y + oy};
    P.h[3] = {f.p1_ccw.x - ox, f.p1_ccw.y - oy};
  }
  if (need_run) {
    int idx, end;
    pen_range(c, from, to, run_cw, idx, end);
    P.run_start = idx;
    P.run_len = pen_run_len(idx, end, c.npen, run_cw);
    P.run_cw = run_cw;
  }
}

// Walk the unit's points in plotting order; `edge(a, b)` receives every internal edge (pre-scaled contour points, the
// direction Polygon.addEdge sees), `ports` the four end points.
template <class E>
Z2D_D void unit_emit(const StrokeCtx& c, const UnitPlan& P, Pt* __restrict__ ports, E&& edge) {
  Pt a_prev{0, 0}, b_prev{0, 0};
  bool a_has = false, b_has = false;
  int k_head = 0, k_run = 0, k_tail = 0, k_in = 0, idx = P.run_start;
#pragma unroll 1
  for (;;) {
    Pt p;
    bool first_list = true;
    if (k_head < P.n_head) {
      p = k_head == 0 ? P.h[0] : k_head == 1 ? P.h[1] : k_head == 2 ? P.h[2] : P.h[3];
      k_head++;
    } else if (k_run < P.run_len) {
      const PenV v = c.pen[idx];
      if (P.run_cw) {
        idx += 1;
        if (idx == c.npen) idx = 0;
      } else {
        if (idx == 0) idx = c.npen;
        idx -= 1;
      }
      p = {P.centre.x + v.px, P.centre.y + v.py};
      k_run++;
    } else if (k_tail < P.n_tail) {
      p = P.t;
      k_tail++;
    } else if (k_in < P.n_in) {
      p = k_in == 0 ? P.in[0] : k_in == 1 ? P.in[1] : P.in[2];
      first_list = false;
      k_in++;
    } else {
      break;
    }
    const Pt sp{p.x * c.scale, p.y * c.scale};  // Contour.plot pre-scales (Polygon.zig:388-391)
    Pt a, b;
    bool have;
    if (first_list == P.swap) {  // prepended to the inner contour
      have = b_has;
      a = sp;
      b = b_prev;
      if (!b_has) ports[kPortBF] = sp;
      ports[kPortBL] = sp;
      b_prev = sp;
      b_has = true;
    } else {  // appended to (or spliced into) the outer contour
      have = a_has;
      a = a_prev;
      b = sp;
      if (!a_has) ports[kPortAF] = sp;
      ports[kPortAL] = sp;
      a_prev = sp;
      a_has = true;
    }
    if (have) edge(a, b);
  }
}


#ifndef Z2D_HOST_TEST
// -------------------------------------------------------------------------------------------------- device side
// The walker is latency bound (dependent f64 chains, 8 of 32 lanes active on average): only every 4th lane of a warp takes a
// sub-path, which shortens the serialised divergent paths of each warp and spreads the work over 4x the warps (config 5, 256
// scenes: flatten 1.80 -> 1.21 ms; config 3: 2.51 -> 2.35 ms).
#ifndef Z2D_WALK_SPREAD
#define Z2D_WALK_SPREAD 4
#endif
#ifndef Z2D_WALK_MIN_CTAS
#define Z2D_WALK_MIN_CTAS 1
#endif
#ifndef Z2D_WALK_THREADS
#define Z2D_WALK_THREADS 32
#endif
constexpr uint32_t kUnitChunk = 8u, kLinkChunk = 16u;  // ids a walker thread takes from the global cursor at a time

// ctr[0] units taken, ctr[1] links taken, ctr[2] edge slots taken (each may run past its capacity: nothing is written there
// and the host redoes the batch with larger arrays); ctr[3], ctr[4]: node ranges of the single-pass fill flattening (kernels.cu).
// The caller zeroes ctr[0..8) before the first kernel of a batch.
struct PoolRec {
  StrokeUnit* units;
  StrokeLink* links;
  uint32_t unit_cap, link_cap;
  uint32_t* ctr;
  uint32_t draw;
  uint32_t u_next = 0, u_end = 0, l_next = 0, l_end = 0;
  uint32_t last_unit = kNoUnit, last_link = kNoUnit;
  struct Mark {
    uint32_t u, l;
  };
  Z2D_D uint32_t unit(uint32_t kind, Pt a, Pt b, Pt c) {
    if (u_next == u_end) {
      u_next = atomicAdd(&ctr[0], kUnitChunk);
      u_end = u_next + kUnitChunk;
    }
    const uint32_t id = u_next++;
    if (id < unit_cap) {
      double2* o = reinterpret_cast<double2*>(units + id);
      uint4 h = make_uint4(kind, draw, last_unit, 0u);
      *reinterpret_cast<uint4*>(o) = h;
      o[1] = make_double2(a.x, a.y);
      o[2] = make_double2(b.x, b.y);
      o[3] = make_double2(c.x, c.y);
    }
    last_unit = id;
    return id;
  }
  Z2D_D void link(uint32_t from, uint32_t to) {
    if (l_next == l_end) {
      l_next = atomicAdd(&ctr[1], kLinkChunk);
      l_end = l_next + kLinkChunk;
    }
    const uint32_t id = l_next++;
    if (id < link_cap) *reinterpret_cast<uint4*>(links + id) = make_uint4(from, to, draw, last_link);
    last_link = id;
  }
  Z2D_D Mark mark() const { return Mark{last_unit, last_link}; }
  // take back everything recorded since the mark (the reference throws the unfinished contour away, dashed_plotter.zig:258-262)
  Z2D_D void rewind(const Mark& m) {
    while (last_unit != m.u && last_unit < unit_cap) {
      const uint32_t pv = units[last_unit].prev;
      units[last_unit].kind = kUnitDead;
      last_unit = pv;
    }
    last_unit = m.u;
    while (last_link != m.l && last_link < link_cap) {
      const uint32_t pv = links[last_link].prev;
      links[last_link].from = kNoUnit;
      last_link = pv;
    }
    last_link = m.l;
  }
  Z2D_D void finish() {  // the ids of the last chunks that were never used
    // (whole 16-byte headers: the consumers read them with one vector load)
    for (; u_next < u_end; u_next++)
      if (u_next < unit_cap) *reinterpret_cast<uint4*>(units + u_next) = make_uint4(kUnitDead, 0u, kNoUnit, 0u);
    for (; l_next < l_end; l_next++)
      if (l_next < link_cap) *reinterpret_cast<uint4*>(links + l_next) = make_uint4(kNoUnit, kNoUnit, 0u, kNoUnit);
  }
};

__global__ void __launch_bounds__(Z2D_WALK_THREADS, Z2D_WALK_MIN_CTAS) k_stroke_walk(const DevSubPath* __restrict__ sps, uint32_t n_sp, const z2d_node* __restrict__ nodes,
                                                    const DevDraw* __restrict__ draws, const PenV* __restrict__ pens,
                                                    const double* __restrict__ dashes, const uint32_t* __restrict__ order,
                                                    StrokeUnit* __restrict__ units, uint32_t unit_cap, StrokeLink* __restrict__ links,
                                                    uint32_t link_cap, uint32_t* __restrict__ ctr) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
#if Z2D_WALK_SPREAD > 1
  if (i % Z2D_WALK_SPREAD) return;
  i /= Z2D_WALK_SPREAD;
#endif
  if (i >= n_sp) return;
  if (order) i = order[i];
  const DevSubPath sp = sps[i];
  if (!(sp.flags & kSpStrokeUnits)) return;
  const DevDraw& d = draws[sp.draw];
  StrokeCtx c;
  stroke_ctx_of(c, d, pens, dashes);
  PoolRec rec{units, links, unit_cap, link_cap, ctr, sp.draw};
  StrokeWalker<PoolRec> w(rec, c);
  if (d.dash_count > 0) w.run_dashed(nodes, sp.node_begin, sp.node_end); else w.run_plain(nodes, sp.node_begin, sp.node_end);
  rec.finish();
}

// Edge slots: `n` per lane, one atomicAdd per warp.  Every lane of the warp must call it.
Z2D_D uint32_t warp_take(uint32_t* cursor, uint32_t n) {
  const uint32_t lane = threadIdx.x & 31u;
  uint32_t incl = n;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const uint32_t v = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= (uint32_t)off) incl += v;
  }
  const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
  uint32_t base = 0;
  if (lane == 31 && total) base = atomicAdd(cursor, total);
  base = __shfl_sync(0xffffffffu, base, 31);
  return base + incl - n;
}

struct PoolSink {  // Polygon.addEdge into pool slots; horizontal edges leave a dead slot (the slots were taken before plotting)
  DevEdge* edges;
  uint32_t* edge_draw;
  uint32_t draw, pos, cap;
  uint32_t n_live = 0;
  double top = INFINITY, bottom = -INFINITY, left = INFINITY, right = -INFINITY;
  double scale = 1.0;  // strokes plot pre-scaled contour points; fills scale here like Polygon.addEdge's caller
  Z2D_D void add(Pt p0, Pt p1) {
    const double ax = p0.x * scale, ay = p0.y * scale, bx = p1.x * scale, by = p1.y * scale;
    const uint32_t at = pos++;
    if (ay == by) {
      if (at < cap) edge_draw[at] = kNoUnit;
      return;
    }
    // Polygon.addEdge: {y0, y1, x of the upper end, dx/dy}.  (bx - ax) / (by - ay) and (ax - bx) / (ay - by) are the same
    // double (both operands negated exactly), so one division serves both directions.
    const DevEdge e{ay, by, ay < by ? ax : bx, (bx - ax) / (by - ay)};
    if (at < cap) {
      edges[at] = e;
      edge_draw[at] = draw;
    }
    const double t = ay < by ? ay : by, b = ay < by ? by : ay;
    const double l = ax < bx ? ax : bx, r = ax < bx ? bx : ax;
    top = t < top ? t : top;
    bottom = b > bottom ? b : bottom;
    left = l < left ? l : left;
    right = r > right ? r : right;
    n_live++;
  }
  Z2D_D void commit(DevDraw* __restrict__ draws) const {
    if (!n_live) return;
    DevDraw& d = draws[draw];
    // most units lie inside what earlier units of the draw already reported: look before paying for an atomic
    const volatile long long* ext = d.ext;
    const long long t = f64_order(top), b = f64_order(bottom), l = f64_order(left), r = f64_order(right);
    if (t < ext[0]) atomicMin(&d.ext[0], t);
    if (b > ext[1]) atomicMax(&d.ext[1], b);
    if (l < ext[2]) atomicMin(&d.ext[2], l);
    if (r > ext[3]) atomicMax(&d.ext[3], r);
    if (*(const volatile uint32_t*)&d.n_edges == 0u) atomicAdd(&d.n_edges, n_live);  // only "any edge at all" is read back
  }
};

#ifndef Z2D_UNITS_MIN_CTAS
#define Z2D_UNITS_MIN_CTAS 4
#endif
__global__ void __launch_bounds__(128, Z2D_UNITS_MIN_CTAS) k_stroke_units(const StrokeUnit* __restrict__ units, uint32_t unit_cap, uint32_t* __restrict__ ctr,
                                                      DevDraw* __restrict__ draws, const PenV* __restrict__ pens,
                                                      const double* __restrict__ dashes, Pt* __restrict__ ports,
                                                      DevEdge* __restrict__ edges, uint32_t* __restrict__ edge_draw, uint32_t edge_cap) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t n_units = min(ctr[0], unit_cap);
  if (blockIdx.x * blockDim.x >= n_units) return;  // whole block past the end
  StrokeCtx c;
  UnitPlan P;
  uint32_t slots = 0, draw = 0;
  bool live = false;
  if (i < n_units) {
    const uint4 h = *reinterpret_cast<const uint4*>(units + i);
    if ((h.x & kUnitKindMask) != kUnitDead) {
      live = true;
      draw = h.y;
      stroke_ctx_of(c, draws[draw], pens, dashes);
      unit_plan(c, h.x, units[i].p, P);
      slots = P.slots();
    }
  }
  const uint32_t base = warp_take(&ctr[2], slots);
  if (!live) return;
  PoolSink sink{edges, edge_draw, draw, base, edge_cap};
  Pt pp[4] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
  unit_emit(c, P, pp, [&](Pt a, Pt b) Z2D_LAMBDA { sink.add(a, b); });
  double2* po = reinterpret_cast<double2*>(ports + (size_t)i * 4);
#pragma unroll
  for (int k = 0; k < 4; k++) po[k] = make_double2(pp[k].x, pp[k].y);
  sink.commit(draws);
}

__global__ void __launch_bounds__(256) k_stroke_links(const StrokeLink* __restrict__ links, uint32_t link_cap, uint32_t unit_cap, uint32_t* __restrict__ ctr,
                                                      DevDraw* __restrict__ draws, const Pt* __restrict__ ports,
                                                      DevEdge* __restrict__ edges, uint32_t* __restrict__ edge_draw, uint32_t edge_cap) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t n_links = min(ctr[1], link_cap);
  if (blockIdx.x * blockDim.x >= n_links) return;
  bool live = false;
  Pt a{0, 0}, b{0, 0};
  uint32_t draw = 0;
  if (i < n_links) {
    const uint4 l = *reinterpret_cast<const uint4*>(links + i);
    // (a port of a unit beyond the capacity was never computed: the host redoes the batch, nothing here is kept)
    if (l.x != kNoUnit && (l.x >> 2) < unit_cap && (l.y >> 2) < unit_cap) {
      const double2 pa = *reinterpret_cast<const double2*>(ports + l.x), pb = *reinterpret_cast<const double2*>(ports + l.y);
      a = {pa.x, pa.y};
      b = {pb.x, pb.y};
      draw = l.z;
      live = a.y != b.y;  // Polygon.addEdge drops horizontal edges
    }
  }
  const uint32_t base = warp_take(&ctr[2], live ? 1u : 0u);
  if (!live) return;
  PoolSink sink{edges, edge_draw, draw, base, edge_cap};
  sink.add(a, b);
  sink.commit(draws);
}
#endif  // Z2D_HOST_TEST

}  // namespace z2d
