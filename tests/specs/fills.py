"""Fill scenes: spec/003-007, 013, 016, 024-026, 028, 031-033, 046-053, 060, 075-079."""
from . import path_scene
from z2d_b200.abi import AntiAliasMode, FillRule, Format, Operator, Precision

WHITE = dict(r=0xFF, g=0xFF, b=0xFF)


def _white_rgb_ctx(z, w, h, aa):
    sfc = z.Surface(Format.rgb, w, h)
    ctx = z.Context(sfc)
    ctx.set_source_to_pixel(z.Pixel.rgb(0xFF, 0xFF, 0xFF))
    ctx.set_anti_aliasing_mode(aa)
    return sfc, ctx


@path_scene("003_fill_triangle")
def s003(z, aa):
    w = h = 300
    sfc, c = _white_rgb_ctx(z, w, h, aa)
    m = 10
    c.move_to(0 + m, 0 + m)
    c.line_to(w - m - 1, 0 + m)
    c.line_to(w // 2 - 1, h - m - 1)
    c.close_path()
    c.fill()
    return sfc


@path_scene("004_fill_square")
def s004(z, aa):
    w = h = 300
    sfc, c = _white_rgb_ctx(z, w, h, aa)
    m = 50
    c.move_to(0 + m, 0 + m)
    c.line_to(w - m - 1, 0 + m)
    c.line_to(w - m - 1, h - m - 1)
    c.line_to(0 + m, h - m - 1)
    c.close_path()
    c.fill()
    return sfc


@path_scene("005_fill_trapezoid")
def s005(z, aa):
    w = h = 300
    sfc, c = _white_rgb_ctx(z, w, h, aa)
    mt, mb, my = 89, 50, 100
    c.move_to(0 + mt, 0 + my)
    c.line_to(w - mt - 1, 0 + my)
    c.line_to(w - mb - 1, h - my - 1)
    c.line_to(0 + mb, h - my - 1)
    c.close_path()
    c.fill()
    return sfc


def _star(c, w, h, m, xs, ys, xo=0, yo=0):
    c.move_to(xo + w // 2, yo + m)
    c.line_to(xo + w - m * xs - 1, yo + h - m - 1)
    c.line_to(xo + m, yo + m * ys)
    c.line_to(xo + w - m - 1, yo + m * ys)
    c.line_to(xo + m * xs, yo + h - m - 1)
    c.close_path()


@path_scene("006_fill_star_even_odd")
def s006(z, aa):
    sfc, c = _white_rgb_ctx(z, 300, 300, aa)
    c.set_fill_rule(FillRule.even_odd)
    _star(c, 300, 300, 20, 3, 5)
    c.fill()
    return sfc


@path_scene("007_fill_bezier")
def s007(z, aa):
    sfc, c = _white_rgb_ctx(z, 300, 300, aa)
    c.move_to(19, 249)
    c.curve_to(89, 49, 209, 49, 279, 249)
    c.close_path()
    c.fill()
    return sfc


@path_scene("013_fill_combined")
def s013(z, aa):
    w, h = 600, 400
    sfc, c = _white_rgb_ctx(z, w, h, aa)
    sw, sh = w // 3, h // 2
    m = 10
    c.move_to(0 + m, 0 + m)
    c.line_to(sw - m - 1, 0 + m)
    c.line_to(sw // 2 - 1, sh - m - 1)
    c.close_path()
    m = 50
    xo = sw
    c.move_to(xo + m, 0 + m)
    c.line_to(xo + sw - m - 1, 0 + m)
    c.line_to(xo + sw - m - 1, sh - m - 1)
    c.line_to(xo + m, sh - m - 1)
    c.close_path()
    tmt, tmb, tmy = 59, 33, 66
    xo = sw * 2
    c.move_to(xo + tmt, 0 + tmy)
    c.line_to(xo + sw - tmt - 1, 0 + tmy)
    c.line_to(xo + sw - tmb - 1, sh - tmy - 1)
    c.line_to(xo + tmb, sh - tmy - 1)
    c.close_path()
    m = 13
    xo = w // 6
    yo = sh
    _star(c, sw, sh, m, 3, 5, xo, yo)
    xo += sw
    c.move_to(xo + 12, yo + 166)
    c.curve_to(xo + 59, yo + 32, xo + 139, yo + 32, xo + 186, yo + 166)
    c.close_path()
    c.fill()
    return sfc


@path_scene("016_fill_star_non_zero")
def s016(z, aa):
    sfc, c = _white_rgb_ctx(z, 300, 300, aa)
    c.set_fill_rule(FillRule.non_zero)
    _star(c, 300, 300, 20, 3, 5)
    c.fill()
    return sfc


@path_scene("024_fill_triangle_direct_cross_format")
def s024(z, aa):
    w = h = 300
    sfc = z.Surface(Format.rgb, w, h)
    c = z.Context(sfc)
    c.set_source_to_pixel(z.Pixel.rgba(0xFF, 0x8A, 0xA5, 0xFF))
    c.set_anti_aliasing_mode(aa)
    m = 10
    c.move_to(0 + m, 0 + m)
    c.line_to(w - m - 1, 0 + m)
    c.line_to(w // 2 - 1, h - m - 1)
    c.close_path()
    c.fill()
    return sfc


@path_scene("025_fill_diamond_clipped")
def s025(z, aa):
    w = h = 300
    sfc, c = _white_rgb_ctx(z, w, h, aa)
    c.move_to(w // 2, 0 - h // 10)
    c.line_to(w + w // 10, h // 2)
    c.line_to(w // 2, h + h // 10)
    c.line_to(0 - w // 10, h // 2)
    c.close_path()
    c.fill()
    return sfc


@path_scene("026_fill_triangle_full")
def s026(z, aa):
    w = h = 300
    sfc, c = _white_rgb_ctx(z, w, h, aa)
    c.move_to(0, 0)
    c.line_to(w, 0)
    c.line_to(w // 2, h)
    c.close_path()
    c.fill()
    return sfc


@path_scene("028_fill_bezier_tolerance")
def s028(z, aa):
    sfc, c = _white_rgb_ctx(z, 900, 300, aa)
    c.set_line_width(5)
    c.move_to(19, 224)
    c.curve_to(89, 49, 209, 49, 279, 224)
    c.close_path()
    c.fill()
    c.set_tolerance(3)
    c.reset_path()
    c.move_to(319, 224)
    c.curve_to(389, 49, 509, 49, 579, 224)
    c.close_path()
    c.fill()
    c.set_tolerance(10)
    c.reset_path()
    c.move_to(619, 224)
    c.curve_to(689, 49, 809, 49, 879, 224)
    c.close_path()
    c.fill()
    return sfc
