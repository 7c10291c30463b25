// Host-side equivalence check of the two stroke tessellators (no GPU needed):
//   stroke.cuh        Stroker: one walk per sub-path, streams edges           (the path every parity test has pinned)
//   stroke_units.cuh  StrokeWalker -> units + links -> unit_plan / unit_emit   (parallel over joins / caps)
// Both are compiled for the host from the very files the library compiles for the device and must produce the same multiset
// of edges for every sub-path of a random corpus (open / closed polylines and Beziers, dashes incl. zero-length ones and
// offsets, every cap / join, duplicate points, several move_tos per sub-path, identity / general / mirrored CTMs).
//
// Also checked here: the merged Pen vertex search of stroke.cuh against the reference's two-copy form, and the upper bound on a
// cubic's segment count (geom.cuh curve_edge_bound, used by the single-pass fill flattening) against Spline.decompose.
//
//   g++ -O1 -std=c++17 -ffp-contract=off -I/usr/local/cuda/include -Iinclude tools/stroke_units_host_test.cpp -o /tmp/sut && /tmp/sut [n]
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include <vector_types.h>
static inline int2 make_int2(int x, int y) { int2 r; r.x = x; r.y = y; return r; }

#include "z2d_cuda.h"

#define Z2D_HOST_TEST 1
#define Z2D_D inline
#define Z2D_DN __attribute__((noinline))
#define Z2D_LAMBDA
using std::max;
using std::min;

namespace z2d {
struct DevEdge {
  double y0, y1, x_start, x_inc;
};
}  // namespace z2d
#include "../z2d_b200/csrc/geom.cuh"
#include "../z2d_b200/csrc/stroke.cuh"
#include "../z2d_b200/csrc/stroke_units.cuh"

using namespace z2d;

// Pen.vertexIteratorFor written the way the reference has it (Pen.zig:138-232): one copy of the two binary searches per
// direction.  stroke.cuh merges the two copies into one loop; this is the yardstick for that rewrite.
static int2 pen_range_reference(const PenV* __restrict__ v, int n, Slope from, Slope to, bool clockwise) {
  int start = 0, end = 0;
  auto cw = [&](int i) Z2D_LAMBDA { return Slope{v[i].cwx, v[i].cwy}; };
  auto ccw = [&](int i) Z2D_LAMBDA { return Slope{v[i].ccwx, v[i].ccwy}; };
  if (clockwise) {
    int low = 0, high = n, i = (low + high) >> 1;
    while (high - low > 1) {
      if (slope_compare(cw(i), from) < 0) low = i; else high = i;
      i = (low + high) >> 1;
    }
    if (slope_compare(cw(i), from) < 0) {
      i += 1;
      if (i == n) i = 0;
    }
    start = i;
    if (slope_compare(to, ccw(i)) >= 0) {
      low = i;
      high = i + n;
      i = (low + high) >> 1;
      while (high - low > 1) {
        const int j = i >= n ? i - n : i;
        if (slope_compare(cw(j), to) > 0) high = i; else low = i;
        i = (low + high) >> 1;
      }
      if (i >= n) i -= n;
    }
    end = i;
  } else {
    int low = 0, high = n, i = (low + high) >> 1;
    while (high - low > 1) {
      if (slope_compare(from, ccw(i)) < 0) low = i; else high = i;
      i = (low + high) >> 1;
    }
    if (slope_compare(from, ccw(i)) < 0) {
      i += 1;
      if (i == n) i = 0;
    }
    start = i;
    if (slope_compare(cw(i), to) <= 0) {
      low = i;
      high = i + n;
      i = (low + high) >> 1;
      while (high - low > 1) {
        const int j = i >= n ? i - n : i;
        if (slope_compare(to, ccw(j)) > 0) high = i; else low = i;
        i = (low + high) >> 1;
      }
      if (i >= n) i -= n;
    }
    end = i;
  }
  return make_int2(max(0, start), max(0, end));
}

struct HostRec {
  std::vector<StrokeUnit> units;
  std::vector<StrokeLink> links;
  struct Mark { size_t u, l; };
  uint32_t unit(uint32_t kind, Pt a, Pt b, Pt c) {
    StrokeUnit u{};
    u.kind = kind;
    u.p[0] = a.x; u.p[1] = a.y; u.p[2] = b.x; u.p[3] = b.y; u.p[4] = c.x; u.p[5] = c.y;
    units.push_back(u);
    return (uint32_t)units.size() - 1;
  }
  void link(uint32_t from, uint32_t to) { links.push_back(StrokeLink{from, to, 0, 0}); }
  Mark mark() const { return Mark{units.size(), links.size()}; }
  void rewind(const Mark& m) {
    for (size_t i = m.u; i < units.size(); i++) units[i].kind = kUnitDead;
    for (size_t i = m.l; i < links.size(); i++) links[i].from = kNoUnit;
  }
};

struct E4 {
  double v[4];
  bool operator<(const E4& o) const { return memcmp(v, o.v, sizeof v) < 0; }
  bool operator==(const E4& o) const { return memcmp(v, o.v, sizeof v) == 0; }
};

static void add_edge(std::vector<E4>& out, Pt p0, Pt p1) {  // EdgeSink::add with scale 1
  const double ax = p0.x, ay = p0.y, bx = p1.x, by = p1.y;
  if (ay < by) out.push_back(E4{{ay, by, ax, (bx - ax) / (by - ay)}});
  else if (ay > by) out.push_back(E4{{ay, by, bx, (ax - bx) / (ay - by)}});
}

static std::vector<PenV> make_pen(double thickness, double tolerance, const double* ctm) {  // Pen.zig init (shape only matters for consistency)
  const double radius = thickness / 2;
  const double major = radius * std::max(std::hypot(ctm[0], ctm[2]), std::hypot(ctm[1], ctm[3]));
  int n;
  if (tolerance >= major * 4) n = 1;
  else if (tolerance >= major) n = 4;
  else {
    n = (int)std::ceil(2.0 * M_PI / std::acos(1 - tolerance / major));
    if (n % 2) n++;
    if (n < 4) n = 4;
  }
  std::vector<PenV> v(n);
  const bool reflect = ctm[0] * ctm[3] - ctm[1] * ctm[2] < 0;
  for (int i = 0; i < n; i++) {
    double theta = 2.0 * M_PI * i / n;
    if (reflect) theta = -theta;
    double dx = radius * std::cos(theta), dy = radius * std::sin(theta);
    const double x = ctm[0] * dx + ctm[1] * dy, y = ctm[2] * dx + ctm[3] * dy;
    v[i].px = x;
    v[i].py = y;
  }
  for (int i = 0; i < n; i++) {
    const int next = (i + 1) % n, prev = (i + n - 1) % n;
    v[i].cwx = v[i].px - v[prev].px; v[i].cwy = v[i].py - v[prev].py;
    v[i].ccwx = v[next].px - v[i].px; v[i].ccwy = v[next].py - v[i].py;
  }
  return v;
}

int main(int argc, char** argv) {
  const int n_cases = argc > 1 ? atoi(argv[1]) : 20000;
  std::mt19937_64 rng(0x7a326433);
  auto uni = [&](double a, double b) { return std::uniform_real_distribution<double>(a, b)(rng); };
  auto pick = [&](int n) { return (int)(rng() % (uint64_t)n); };
  std::vector<DevEdge> buf(1 << 20);
  std::vector<uint32_t> dbuf(1 << 20);
  size_t total_edges = 0, total_units = 0, total_links = 0;
  int bad = 0;
  for (int cs = 0; cs < n_cases; cs++) {
    // ---- a random sub-path
    std::vector<z2d_node> nodes;
    auto node = [&](uint32_t tag, double a = 0, double b = 0, double c = 0, double d = 0, double e = 0, double f = 0) {
      z2d_node n{};
      n.tag = tag;
      n.p[0] = a; n.p[1] = b; n.p[2] = c; n.p[3] = d; n.p[4] = e; n.p[5] = f;
      nodes.push_back(n);
    };
    const double span = pick(4) == 0 ? 8.0 : 120.0;
    const bool snap = pick(3) == 0;
    auto coord = [&]() { double v = uni(0, span); return snap ? std::round(v * 2) / 2 : v; };
    const int n_moves = pick(8) == 0 ? 1 + pick(3) : 1;
    for (int mv = 0; mv < n_moves; mv++) {
      double x = coord(), y = coord();
      const double sx = x, sy = y;
      node(Z2D_NODE_MOVE_TO, x, y);
      const int segs = pick(10) == 0 ? 0 : 1 + pick(9);
      for (int s = 0; s < segs; s++) {
        const int r = pick(12);
        if (r == 0) {
          node(Z2D_NODE_LINE_TO, x, y);  // duplicate point
        } else if (r == 1 && s > 1) {
          x = sx; y = sy;
          node(Z2D_NODE_LINE_TO, x, y);  // back to the start
        } else if (r < 8) {
          x = coord(); y = coord();
          node(Z2D_NODE_LINE_TO, x, y);
        } else if (r == 8) {  // collinear continuation / reversal
          const double k = pick(2) ? 2.0 : -1.0;
          const z2d_node& pv = nodes.back();
          const double px = pv.tag == Z2D_NODE_CURVE_TO ? pv.p[4] : pv.p[0], py = pv.tag == Z2D_NODE_CURVE_TO ? pv.p[5] : pv.p[1];
          (void)px; (void)py;
          x = x + k * 3.0; y = y + k * 1.5;
          node(Z2D_NODE_LINE_TO, x, y);
        } else {
          const double a = coord(), b = coord(), c2 = coord(), d = coord();
          x = coord(); y = coord();
          if (pick(10) == 0) node(Z2D_NODE_CURVE_TO, nodes.back().p[0], nodes.back().p[1], x, y, x, y);
          else node(Z2D_NODE_CURVE_TO, a, b, c2, d, x, y);
        }
      }
      if (pick(3) == 0) {
        node(Z2D_NODE_CLOSE_PATH);
        if (pick(6) == 0) node(Z2D_NODE_LINE_TO, coord(), coord());  // ignored: no current point after a close
      }
    }
    // ---- random stroke parameters
    StrokeCtx c{};
    c.cap = (uint32_t)pick(3);
    c.join = (uint32_t)pick(3);
    c.thickness = pick(5) == 0 ? uni(0.2, 1.0) : uni(1.0, 14.0);
    c.miter_limit = pick(2) ? 10.0 : uni(1.0, 4.0);
    c.scale = pick(2) ? 4.0 : 1.0;
    c.tolerance = 0.1;
    const int ct = pick(5);
    double m[6] = {1, 0, 0, 1, 0, 0};
    if (ct == 1) { m[0] = 2.0; m[3] = 0.5; }
    if (ct == 2) { const double t = uni(0, 6.28); m[0] = std::cos(t); m[1] = -std::sin(t); m[2] = std::sin(t); m[3] = std::cos(t); }
    if (ct == 3) { m[0] = -1.5; m[1] = 0.3; m[2] = 0.2; m[3] = 1.1; }
    if (ct == 4) { m[0] = 1.0; m[1] = 0.7; m[2] = 0.0; m[3] = 1.0; }
    const double det = m[0] * m[3] - m[1] * m[2];
    const double inv[6] = {m[3] / det, -m[1] / det, -m[2] / det, m[0] / det, 0, 0};
    for (int i = 0; i < 6; i++) { c.ctm[i] = m[i]; c.inv[i] = inv[i]; }
    c.ctm_identity = ct == 0;
    c.det_nonneg = det >= 0;
    std::vector<PenV> pen = make_pen(c.thickness, c.tolerance, m);
    c.pen = pen.data();
    c.npen = (int)pen.size();
    double dashes[4];
    int nd = 0;
    if (pick(2)) {
      nd = 1 + pick(4);
      for (int i = 0; i < nd; i++) dashes[i] = pick(7) == 0 ? 0.0 : uni(0.5, 30.0);
      bool all_zero = true;
      for (int i = 0; i < nd; i++) all_zero &= dashes[i] == 0.0;
      if (all_zero) dashes[0] = 3.0;
    }
    c.dashes = dashes;
    c.ndash = nd;
    c.dash_offset = pick(3) == 0 ? uni(-40, 40) : 0.0;

    // ---- sub-path stroker
    EdgeSink<true> sink;
    sink.scale = 1.0;
    sink.out = buf.data();
    sink.out_draw = dbuf.data();
    sink.limit = (uint32_t)buf.size();
    {
      Stroker<EdgeSink<true>> s(sink, c);
      if (nd > 0) s.run_dashed(nodes.data(), 0, (uint32_t)nodes.size()); else s.run_plain(nodes.data(), 0, (uint32_t)nodes.size());
    }
    std::vector<E4> ref(sink.n);
    for (uint32_t i = 0; i < sink.n; i++) ref[i] = E4{{buf[i].y0, buf[i].y1, buf[i].x_start, buf[i].x_inc}};

    // ---- unit stroker
    HostRec rec;
    {
      StrokeWalker<HostRec> w(rec, c);
      if (nd > 0) w.run_dashed(nodes.data(), 0, (uint32_t)nodes.size()); else w.run_plain(nodes.data(), 0, (uint32_t)nodes.size());
    }
    std::vector<E4> got;
    std::vector<Pt> ports(rec.units.size() * 4, Pt{NAN, NAN});
    size_t slots = 0;
    for (size_t u = 0; u < rec.units.size(); u++) {
      if ((rec.units[u].kind & kUnitKindMask) == kUnitDead) continue;
      UnitPlan P;
      unit_plan(c, rec.units[u].kind, rec.units[u].p, P);
      size_t before = got.size(), calls = 0;
      unit_emit(c, P, &ports[u * 4], [&](Pt a, Pt b) { calls++; add_edge(got, a, b); });
      (void)before;
      if (calls != P.slots()) {
        printf("case %d unit %zu: %zu edges plotted, plan says %u\n", cs, u, calls, P.slots());
        bad++;
      }
      slots += P.slots();
    }
    for (const StrokeLink& l : rec.links) {
      if (l.from == kNoUnit) continue;
      const Pt a = ports[l.from], b = ports[l.to];
      if (std::isnan(a.x) || std::isnan(b.x)) {
        printf("case %d: link to a port that was never written (%u -> %u)\n", cs, l.from, l.to);
        bad++;
        continue;
      }
      add_edge(got, a, b);
    }
    std::sort(ref.begin(), ref.end());
    std::sort(got.begin(), got.end());
    if (!(ref == got)) {
      bad++;
      if (bad < 10)
        printf("case %d MISMATCH: %zu vs %zu edges (dashes %d, cap %u, join %u, nodes %zu, ctm %d)\n", cs, ref.size(), got.size(), nd, c.cap,
               c.join, nodes.size(), ct);
    }
    total_edges += ref.size();
    total_units += rec.units.size();
    total_links += rec.links.size();
  }
  // ---- the merged pen search against the two-copy form
  size_t pen_checks = 0;
  for (int t = 0; t < 200000; t++) {
    double m[6] = {1, 0, 0, 1, 0, 0};
    if (t % 3 == 1) { m[0] = -1.5; m[1] = 0.3; m[2] = 0.2; m[3] = 1.1; }
    if (t % 3 == 2) { m[0] = 2.0; m[3] = 0.5; }
    std::vector<PenV> pen = make_pen(uni(0.3, 20.0), 0.1, m);
    for (int k = 0; k < 8; k++) {
      Slope from, to;
      if (pick(3) == 0) {  // exactly along pen edges: the ties the searches are written around
        const PenV& a = pen[pick((int)pen.size())];
        const PenV& b = pen[pick((int)pen.size())];
        from = pick(2) ? Slope{a.cwx, a.cwy} : Slope{a.ccwx, a.ccwy};
        to = pick(2) ? Slope{b.cwx, b.cwy} : Slope{-from.dx, -from.dy};
      } else {
        from = Slope{uni(-1, 1), uni(-1, 1)};
        to = pick(4) == 0 ? Slope{-from.dx, -from.dy} : Slope{uni(-1, 1), uni(-1, 1)};
      }
      if (pick(5) == 0) from.dx = 0;
      if (pick(5) == 0) to.dy = 0;
      for (int cwi = 0; cwi < 2; cwi++) {
        const int2 a = pen_range_of(pen.data(), (int)pen.size(), from, to, cwi != 0);
        const int2 b = pen_range_reference(pen.data(), (int)pen.size(), from, to, cwi != 0);
        pen_checks++;
        if (a.x != b.x || a.y != b.y) {
          if (bad < 20) printf("pen_range mismatch: n %zu cw %d -> (%d,%d) vs (%d,%d)\n", pen.size(), cwi, a.x, a.y, b.x, b.y);
          bad++;
        }
      }
    }
  }
  printf("%zu pen searches compared\n", pen_checks);
  // ---- the edge bound of the single-pass fill flattening against Spline.decompose itself
  size_t curve_checks = 0, bound_sum = 0, seg_sum = 0;
  for (int t = 0; t < 400000; t++) {
    const double span = t % 5 == 0 ? 4.0 : (t % 5 == 1 ? 4000.0 : 300.0);
    Pt q[4];
    for (int k = 0; k < 4; k++) q[k] = Pt{uni(0, span), uni(0, span)};
    if (t % 11 == 0) q[1] = q[0];                       // degenerate control polygons
    if (t % 13 == 0) q[2] = q[3];
    if (t % 17 == 0) q[3] = q[0];                       // closed loop
    if (t % 19 == 0) { q[1] = q[0]; q[2] = q[3]; }      // straight line (Spline.zig:39-42)
    if (t % 23 == 0) { q[2] = Pt{q[0].x + 2 * (q[1].x - q[0].x), q[0].y + 2 * (q[1].y - q[0].y)}; }  // collinear
    const double tol = t % 7 == 0 ? 0.001 : (t % 7 == 1 ? 1.5 : 0.1);
    size_t segs = 0;
    spline_decompose(q[0], q[1], q[2], q[3], tol * tol, [&](Pt) { segs++; });
    const uint32_t bound = curve_edge_bound(q[0], q[1], q[2], q[3], tol);
    curve_checks++;
    bound_sum += bound;
    seg_sum += segs;
    if (segs > bound) {
      if (bad < 20) printf("curve bound violated: %zu segments, bound %u (tol %g)\n", segs, bound, tol);
      bad++;
    }
  }
  printf("%zu curves: bound / segments = %.2f\n", curve_checks, (double)bound_sum / (double)seg_sum);
  printf("%d cases, %zu edges, %zu units, %zu links: %d mismatches\n", n_cases, total_edges, total_units, total_links, bad);
  return bad ? 1 : 0;
}
