// Plotter geometry shared by the flatten kernels (kernels.cu), the sub-path stroker (stroke.cuh) and the unit stroker
// (stroke_units.cuh): points, cubic subdivision (tess/Spline.zig) and the edge sink (tess/Polygon.zig addEdge).  Plain
// arithmetic only, so that tools/stroke_units_host_test.cpp can compile the same code for the host (Z2D_D / Z2D_DN /
// Z2D_LAMBDA are defined by the including translation unit).
#pragma once

namespace z2d {

// (z2d::DevEdge {y0, y1, x_start, x_inc} is declared by the includer: z2d_batch.cuh in the library)

struct Pt {
  double x, y;
};
Z2D_D bool pt_eq(Pt a, Pt b) { return a.x == b.x && a.y == b.y; }

struct Knots {
  Pt a, b, c, d;
};

Z2D_D double knots_error_sq(const Knots& k) {  // tess/Spline.zig:83-123
  double bx = k.b.x - k.a.x, by = k.b.y - k.a.y, cx = k.c.x - k.a.x, cy = k.c.y - k.a.y;
  if (k.a.x != k.d.x || k.a.y != k.d.y) {
    double dx = k.d.x - k.a.x, dy = k.d.y - k.a.y;
    double dd = dx * dx + dy * dy;
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
  double be = bx * bx + by * by, ce = cx * cx + cy * cy;
  return be > ce ? be : ce;
}
Z2D_D Pt lerp_half(Pt a, Pt b) { return {a.x + ((b.x - a.x) / 2), a.y + ((b.y - a.y) / 2)}; }
Z2D_D Knots knots_split(Knots& k) {  // tess/Spline.zig:128-151 (k becomes the first half)
  Pt ab = lerp_half(k.a, k.b), bc = lerp_half(k.b, k.c), cd = lerp_half(k.c, k.d);
  Pt abbc = lerp_half(ab, bc), bccd = lerp_half(bc, cd);
  Pt fin = lerp_half(abbc, bccd);
  Knots r{fin, bccd, cd, k.d};
  k.b = ab;
  k.c = abbc;
  k.d = fin;
  return r;
}

// Edge sink: applies Polygon.addEdge (tess/Polygon.zig:61-109).  EMIT=false counts and tracks extents.
template <bool EMIT>
struct EdgeSink {
  bool unpaired = false;
  double scale;
  uint32_t n = 0;
  double top = 0, bottom = 0, left = 0, right = 0;
  DevEdge* out = nullptr;
  uint32_t* out_draw = nullptr;
  uint32_t draw = 0;
  uint32_t limit = 0xffffffffu;  // emit pass: number of edges the count pass found for this sink
  // The plotters emit a contour's edges as its points arrive.  Where the reference throws an unfinished contour away (its points
  // were only buffered), the edges emitted since a mark are taken back: count and extents in the count pass, the write position
  // in the emit pass.
  struct Mark {
    uint32_t n;
    double top, bottom, left, right;
  };
  Z2D_D Mark mark() const { return Mark{n, top, bottom, left, right}; }
  Z2D_D void rewind(const Mark& m) {
    n = m.n;
    top = m.top;
    bottom = m.bottom;
    left = m.left;
    right = m.right;
  }
  Z2D_D void add(Pt p0, Pt p1) {
    double ax = p0.x * scale, ay = p0.y * scale, bx = p1.x * scale, by = p1.y * scale;
    DevEdge e;
    if (ay < by) {
      e = {ay, by, ax, (bx - ax) / (by - ay)};
    } else if (ay > by) {
      e = {ay, by, bx, (ax - bx) / (ay - by)};
    } else {
      return;
    }
    if (EMIT) {
      // positions at or beyond the count pass's total are always taken back by a later rewind(): never touch the slots of the
      // next sub-path, which another thread is writing
      if (n < limit) {
        out[n] = e;
        out_draw[n] = draw;
      }
    } else {
      double t = ay < by ? ay : by, b = ay < by ? by : ay;
      double l = ax < bx ? ax : bx, r = ax < bx ? bx : ax;
      if (n == 0) {
        top = t; bottom = b; left = l; right = r;
      } else {
        if (t < top) top = t;
        if (b > bottom) bottom = b;
        if (l < left) left = l;
        if (r > right) right = r;
      }
    }
    n++;
  }
};

constexpr int kSplineStack = 48;

// Iterative Spline.decompose (tess/Spline.zig:37-71): depth first, left half first; emits the START point of every accepted
// piece except the very first, then the end point.  Only right halves are stacked (the left half stays in registers), which
// keeps the local-memory traffic of the flatten kernels to one 64-byte store per split and one load per emitted point.
template <class F>
Z2D_D void spline_decompose(Pt a, Pt b, Pt c, Pt d, double tol_sq, F&& line_to) {
  if (pt_eq(a, b) && pt_eq(c, d)) {  // Spline.zig:39-42
    line_to(d);
    return;
  }
  Knots stack[kSplineStack];
  int sp = 0;
  Knots k{a, b, c, d};
  #pragma unroll 1
  for (;;) {
    while (!(knots_error_sq(k) < tol_sq || sp >= kSplineStack - 2)) stack[sp++] = knots_split(k);  // k becomes the left half
    if (!pt_eq(k.a, a)) line_to(k.a);
    if (sp == 0) break;
    k = stack[--sp];
  }
  line_to(d);
}

// Upper bound on the number of line segments Spline.decompose yields for a cubic (the single-pass flattening of kernels.cu
// sizes a curve's edge range with it; tools/stroke_units_host_test.cpp checks it against the subdivision on random curves).
// With M = max |second difference| of the control polygon, Knots.errorSq <= M^2, and a de Casteljau halving divides second
// differences by 4, so no piece is split beyond depth k = min{k : M^2 / 16^k < tol^2}.
Z2D_D uint32_t curve_edge_bound(Pt a, Pt b, Pt c, Pt e, double tol) {
  if (pt_eq(a, b) && pt_eq(c, e)) return 1u;  // Spline.zig:39-42
  const double d0x = a.x - 2.0 * b.x + c.x, d0y = a.y - 2.0 * b.y + c.y;
  const double d1x = b.x - 2.0 * c.x + e.x, d1y = b.y - 2.0 * c.y + e.y;
  const double m0 = d0x * d0x + d0y * d0y, m1 = d1x * d1x + d1y * d1y;
  double ratio = (m0 > m1 ? m0 : m1) / (tol * tol) * 1.01;  // (1 % for the rounding of the halving arithmetic)
  uint32_t k = 0;
  while (ratio >= 1.0 && k < 20u) {
    ratio *= 0.0625;
    k++;
  }
  return 1u << k;
}

}  // namespace z2d
