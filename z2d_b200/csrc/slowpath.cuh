// Order-dependent / globally-coupled draw modes that do not fit the tile pipeline and are
// executed as isolated single-draw launches (included by kernels.cu):
//
//   k_hairline          raster/hairline.zig + tess/polyline_plotter.zig: 1-pixel Bresenham / Wu lines.
//                       Pixels shared by consecutive segments are composited twice, in order, so a
//                       draw is walked sequentially by one thread (hairline draws are tiny).
// (raster/direct.zig with an unbounded operator, once isolated too, now runs in the tile pipeline from the row records of
// k_edge_sim, kernels.cu.)
#pragma once

namespace z2d {

// compositeOpaque / compositeOpacity (raster/shared.zig:9-107) on one pixel of the surface
Z2D_D void px_composite(const DevSurface& S, const DevDraw& d, const GradTables& T, int x, int y, bool use_opacity, int opacity) {
  if (x < 0 || y < 0 || x >= S.w || y >= S.vh) return;  // surface.zig:510,538,566
  if (y < S.y0 || y >= S.y0 + S.h) return;              // band surface: row not held here
  const size_t idx = (size_t)(y - S.y0) * (size_t)S.w + (size_t)x;
  const uint32_t fmt = S.fmt;
  uint32_t raw = load_raw(S.data, fmt, idx);
  if (d.op == Z2D_OP_CLEAR) {
    raw = 0u;
  } else if (d.reduces) {
    if (!use_opacity) {
      raw = d.paint_raw;
    } else {  // surface.zig:557-581
      RGBA16 s = unpack_rgba(d.src.px_rgba);
      if (opacity < 255) s = mask_mul16(s, opacity);
      raw = rgba16_to_raw(fmt, int_op_sw(d.op, raw_to_rgba16(fmt, raw), s));
    }
  } else {
    raw = composite_generic(d, T, d.precision, fmt, raw, opacity, use_opacity, x, y);
  }
  store_raw(S.data, fmt, idx, raw);
}

struct HairCtx {
  const DevSurface& S;
  const DevDraw& d;
  const GradTables& T;
};

Z2D_D void hl_opaque(const HairCtx& c, int x, int y) { px_composite(c.S, c.d, c.T, x, y, false, 255); }
Z2D_D void hl_opacity(const HairCtx& c, int x, int y, int o) { px_composite(c.S, c.d, c.T, x, y, true, o); }

Z2D_D uint32_t hl_err_inc(int a, int b) {  // hairline.zig:383-396
  if (a == b) return 0xffffu;
  return (uint32_t)((((uint32_t)a) << 16) / (uint32_t)b) & 0xffffu;
}

Z2D_D void hl_draw_line(const HairCtx& c, int x0, int y0, int x1, int y1) {  // hairline.zig:94-364
  const int W = c.S.w, H = c.S.vh;
  if ((x0 < 0 || x0 >= W) && (x1 < 0 || x1 >= W)) return;
  if ((y0 < 0 || y0 >= H) && (y1 < 0 || y1 >= H)) return;
  const int adx = abs(x1 - x0), ady = abs(y1 - y0);
  const bool aa = c.d.hair_aa != Z2D_AA_NONE;
  if (adx == 0) {
    const int sy = max(0, min(min(y0, y1), H - 1)), ey = max(0, min(max(y0, y1), H - 1));
    for (int y = sy; y <= ey; y++) hl_opaque(c, x0, y);
  } else if (ady == 0) {
    const int sx = max(0, min(min(x0, x1), W - 1)), ex = max(0, min(max(x0, x1), W - 1));
    for (int x = sx; x <= ex; x++) hl_opaque(c, x, y0);
  } else if (adx < ady) {  // y-major
    if (y0 > y1) {
      int t = x0; x0 = x1; x1 = t;
      t = y0; y0 = y1; y1 = t;
    }
    const int dy = y1 - y0, dx = abs(x1 - x0), sx = x1 < x0 ? -1 : 1;
    if (!aa) {  // Bresenham (hairline.zig:193-222)
      int x = x0, dd = 2 * dx - dy;
      for (int y = y0; y <= y1; y++) {
        hl_opaque(c, x, y);
        if (dd > 0) { x += sx; dd -= 2 * dy; }
        dd += 2 * dx;
      }
    } else {  // Wu (hairline.zig:300-364)
      int x = x0;
      uint32_t err = 0;
      const uint32_t inc = hl_err_inc(dx, dy);
      hl_opaque(c, x, y0);
      for (int y = y0 + 1; y < y1; y++) {
        const uint32_t sum = err + inc;
        err = sum & 0xffffu;
        if (sum > 0xffffu) x += sx;
        const int oc = (int)(err >> 8);
        hl_opacity(c, x, y, oc ^ 0xff);
        hl_opacity(c, x + sx, y, oc);
      }
      hl_opaque(c, x1, y1);
    }
  } else {  // x-major
    if (x0 > x1) {
      int t = x0; x0 = x1; x1 = t;
      t = y0; y0 = y1; y1 = t;
    }
    const int dx = x1 - x0, dy = abs(y1 - y0), sy = y1 < y0 ? -1 : 1;
    if (!aa) {
      int y = y0, dd = 2 * dy - dx;
      for (int x = x0; x <= x1; x++) {
        hl_opaque(c, x, y);
        if (dd > 0) { y += sy; dd -= 2 * dx; }
        dd += 2 * dy;
      }
    } else {
      int y = y0;
      uint32_t err = 0;
      const uint32_t inc = hl_err_inc(dy, dx);
      hl_opaque(c, x0, y);
      for (int x = x0 + 1; x < x1; x++) {
        const uint32_t sum = err + inc;
        err = sum & 0xffffu;
        if (sum > 0xffffu) y += sy;
        const int oc = (int)(err >> 8);
        hl_opacity(c, x, y, oc ^ 0xff);
        hl_opacity(c, x, y + sy, oc);
      }
      hl_opaque(c, x1, y1);
    }
  }
}

// polyline_plotter.plot + hairline.run fused: contours are rasterised as their points arrive
// (same order as building the contour list first and walking it afterwards).
__global__ void k_hairline(const DevSurface* __restrict__ sfcs, const DevDraw* __restrict__ draws, uint32_t draw_index,
                           const z2d_node* __restrict__ nodes, uint32_t node_begin, uint32_t node_end, const double* __restrict__ dashes,
                           GradTables T) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const DevDraw& d = draws[draw_index];
  const DevSurface S = sfcs[d.surface];
  const HairCtx hc{S, d, T};

  // current contour: only its last point and length matter
  Pt c_last{0, 0};
  uint32_t c_len = 0;
  auto contour_plot = [&](Pt p) Z2D_LAMBDA {
    if (c_len >= 1)
      hl_draw_line(hc, (int)round_half_away(c_last.x), (int)round_half_away(c_last.y), (int)round_half_away(p.x), (int)round_half_away(p.y));
    c_last = p;
    c_len++;
  };
  auto contour_end = [&]() Z2D_LAMBDA {  // contour appended to the result list
    if (c_len == 1) hl_opaque(hc, (int)round_half_away(c_last.x), (int)round_half_away(c_last.y));
    c_len = 0;
  };

  // PointBuffer(1, 2)
  Pt p_first{0, 0}, p_last{0, 0};
  int p_len = 0;
  auto pts_add = [&](Pt p) Z2D_LAMBDA {
    if (p_len == 0) p_first = p;
    if (p_len < 2) p_len++;
    p_last = p;
  };

  // Dasher (tess/Dasher.zig)
  const bool dashed = d.dash_count > 0;
  const double* dd = dashes + d.dash_begin;
  const int dn = (int)d.dash_count;
  int d_idx = 0;
  bool d_on = true;
  double d_remain = 0;
  auto dash_reset = [&]() Z2D_LAMBDA {
    d_idx = 0;
    d_on = true;
    d_remain = dd[0];
    d_remain -= d.dash_offset;
    while (d_remain < 0 || d_remain > dd[d_idx]) {
      if (d_remain < 0) {
        d_remain += dd[d_idx];
        d_idx = (d_idx >= dn - 1) ? 0 : d_idx + 1;
      } else {
        d_remain -= dd[d_idx];
        d_idx = (d_idx == 0) ? dn - 1 : d_idx - 1;
      }
      d_on = !d_on;
    }
  };
  auto dash_step = [&](double len) Z2D_LAMBDA -> bool {
    d_remain -= len;
    if (d_remain <= 0) {
      d_on = !d_on;
      d_idx += 1;
      if (d_idx >= dn) d_idx = 0;
      d_remain = dd[d_idx];
      return true;
    }
    return false;
  };
  if (dashed) dash_reset();

  auto dashed_line_to = [&](Pt p0, Pt p1) Z2D_LAMBDA {  // polyline_plotter.zig:183-221
    Slope s{p1.x - p0.x, p1.y - p0.y};
    const double total = slope_normalize(s);
    double remaining = total;
    double step = fmin(d_remain, remaining);
    while (remaining > 0) {
      remaining -= step;
      const Pt dp{p0.x + s.dx * (total - remaining), p0.y + s.dy * (total - remaining)};
      if (!pt_eq(dp, p_last)) pts_add(dp);
      if (d_on) contour_plot(dp);
      if (dash_step(step)) {
        contour_end();
        if (d_on) contour_plot(dp);
      }
      step = fmin(d_remain, remaining);
    }
  };
  auto line_to = [&](Pt p) Z2D_LAMBDA {
    if (p_len == 0) return;
    const Pt last = p_last;
    if (pt_eq(last, p)) return;
    if (dashed) {
      dashed_line_to(last, p);
    } else {
      contour_plot(p);
      pts_add(p);
    }
  };

  const double tol_sq = d.hair_tolerance * d.hair_tolerance;
  for (uint32_t i = node_begin; i < node_end; i++) {
    const z2d_node nd = nodes[i];
    switch (nd.tag) {
      case Z2D_NODE_MOVE_TO:
        if (c_len != 0) contour_end();
        p_len = 0;
        if (dashed) dash_reset();
        if (i == node_end - 1) break;  // trailing auto move_to (polyline_plotter.zig:56-58)
        contour_plot({nd.p[0], nd.p[1]});
        pts_add({nd.p[0], nd.p[1]});
        break;
      case Z2D_NODE_LINE_TO: line_to({nd.p[0], nd.p[1]}); break;
      case Z2D_NODE_CURVE_TO: {
        if (p_len == 0) break;
        const Pt a = p_last, b{nd.p[0], nd.p[1]}, cc{nd.p[2], nd.p[3]}, e{nd.p[4], nd.p[5]};
        spline_decompose(a, b, cc, e, tol_sq, line_to);
        break;
      }
      default:  // close_path (polyline_plotter.zig:104-131)
        if (p_len >= 2) {
          if (pt_eq(p_last, p_first)) break;
          if (dashed) {
            dashed_line_to(p_last, p_first);
          } else {
            contour_plot(p_first);
            pts_add(p_first);
          }
        }
    }
  }
  if (c_len != 0) contour_end();
}

}  // namespace z2d
