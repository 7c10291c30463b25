"""The reference's text scenes (spec/074_text.zig, spec/080_fill_z2d_logo.zig = BASELINE config 1,
spec/085_deja_sans_ignore_invalid_points.zig).  Glyph outlines come from tests/specs/ttf.py (host-side test tooling);
everything downstream of the node list is the hot path under test."""
from z2d_b200.abi import Format
from z2d_b200.host import FillOptions

from . import path_scene, ttf

_FONTS = {}


def _font(name):
    if name not in _FONTS:
        _FONTS[name] = ttf.Font(ttf.font_bytes(name))
    return _FONTS[name]


def _context_show_text(context, font, size, text, x, y):  # Context.showText (Context.zig:651-677)
    ttf.show_text(context.surface, context._wrap_dither(), font, text, x, y, size,
                  FillOptions(context.anti_aliasing_mode, context.fill_rule, context.operator, context.precision, context.tolerance),
                  context.transformation)


@path_scene("074_text")
def s074(z, aa_mode):
    width, height = 900, 100
    sfc = z.Surface(Format.rgb, width, height)
    font = _font("Inter-Regular.ttf")
    text = "The quick brown fox jumps over the lázy dog"
    white = z.Pattern.opaque(z.Pixel.rgb(0xFF, 0xFF, 0xFF))
    ttf.show_text(sfc, white, font, text, 10, 0, 27, z.FillOptions(anti_aliasing_mode=aa_mode))
    context = z.Context(sfc)
    context.set_source_to_pixel(z.Pixel.rgb(0xFF, 0xFF, 0xFF))
    context.set_anti_aliasing_mode(aa_mode)
    _context_show_text(context, font, 27, text, 10, 30)
    g = z.Gradient.linear(10, 60, 900, 90)
    g.add_stop(0, {"rgb": (1, 0, 0)})
    g.add_stop(0.5, {"rgb": (0, 1, 0)})
    g.add_stop(1, {"rgb": (0, 0, 1)})
    context.set_source(z.Pattern.gradient(g))
    context.scale(1.5, 1.0)
    _context_show_text(context, font, 27, text, 10, 60)
    return sfc


@path_scene("080_fill_z2d_logo")
def s080(z, aa_mode):
    sfc = z.Surface(Format.rgba, 601, 172)
    context = z.Context(sfc)
    context.set_source_to_pixel(z.Pixel.rgb(0xF7, 0xA4, 0x1D))
    context.set_anti_aliasing_mode(aa_mode)

    def poly(points):  # note: the scene never resets the path, so earlier polygons are filled again (harmless: opaque source)
        context.move_to(*points[0])
        for p in points[1:]:
            context.line_to(*p)
        context.close_path()
        context.fill()

    context.translate(129, 0)
    poly([(0, 22), (0, 117), (12, 117), (31, 95), (22, 95), (22, 44), (28, 44), (46, 22)])
    context.translate(37, 0)
    poly([(113, 0), (64, 22), (19, 22), (0, 44), (45.728516, 44), (-34, 140), (15, 117), (60, 117), (79, 95), (33.427734, 95)])
    _context_show_text(context, _font("Montserrat-ExtraBold.ttf"), 128, "2d", 86, -11)
    context.translate(253, 0)
    poly([(0, 22), (0, 45), (25, 45), (25, 95), (0, 95), (0, 117), (47, 117), (47, 22)])
    context.set_identity()
    context.translate(0, 135)
    _context_show_text(context, _font("Montserrat-Bold.ttf"), 36, "A PURE ZIG GRAPHICS LIBRARY", 0, 0)
    return sfc


@path_scene("085_deja_sans_ignore_invalid_points")
def s085(z, aa_mode):
    sfc = z.Surface(Format.rgb, 58, 62)
    pattern = z.Pattern.opaque(z.Pixel.from_color({"rgb": (1, 1, 1)}))
    ttf.show_text(sfc, pattern, _font("DejaVuSans.ttf"), "żu", 10, 10, 32, z.FillOptions(anti_aliasing_mode=aa_mode))
    return sfc
