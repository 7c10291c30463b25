"""Hand-ported scenes that use the unmanaged painter / compositor API directly (spec/082)."""
from . import path_scene
from z2d_b200.abi import Format, Operator

_CASES_082 = [(-25, 50, 75, 60), (25, 50, 125, 60), (-25, 50, 75, 50), (25, 50, 125, 50), (25, -50, 75, 50), (25, 50, 75, 150),
              (-25, -50, -75, -150), (125, 150, 175, 250), (50, -25, 60, 75), (50, 25, 60, 125), (50, -25, 50, 75), (50, 25, 50, 125),
              (-50, 25, 50, 75), (50, 25, 150, 75), (-50, -25, -150, -75), (150, 125, 250, 175)]


@path_scene("082_stroke_hairline_clip")
def s082(z, aa_mode):
    width, height = 200, len(_CASES_082) // 2 * 100
    sfc = z.Surface(Format.rgb, width, height)
    for idx, (x0, y0, x1, y1) in enumerate(_CASES_082):
        on = 0x20 * (((idx // 2) & 1) ^ (idx % 2))  # Zig: & and ^ share one precedence level, left to right
        src = z.SurfacePixel(z.Pixel.rgb(on, on, on), 100, 100)
        path = z.Path()  # StaticPath(2)
        path.move_to(float(x0), float(y0))
        path.line_to(float(x1), float(y1))
        z.painter.stroke(src, z.Pattern.opaque(z.Pixel.rgb(0xFF, 0xFF, 0xFF)), path.nodes,
                         z.StrokeOptions(anti_aliasing_mode=aa_mode, hairline=True))
        z.SurfaceCompositor.run(sfc, idx % 2 * 100, idx // 2 * 100, [z.Operation(Operator.src, src=z.Param.surface(src))])
    return sfc
