"""Compositor / gradient / dither / alpha-format scenes:
spec/001, 002, 023, 046-053, 059, 061-072."""
import math

from . import compositor_scene, path_scene
from z2d_b200.abi import (AntiAliasMode, DitherType, FillRule, Format, Interp, JoinMode, Operator, Polar, Precision)

SMILE = """\
                             0000000000000000000000000                             
                        00000000000000000000000000000000000                        
                    0000000000000000000000000000000000000000000                    
                 0000000000000000000000000000000000000000000000000                 
               00000000000000000000000000000000000000000000000000000               
             000000000000000000000000000000000000000000000000000000000             
           0000000000000000000000000000000000000000000000000000000000000           
         00000000000000000000000000000000000000000000000000000000000000000         
        0000000000000000000000000000000000000000000000000000000000000000000        
      00000000000000000000000000000000000000000000000000000000000000000000000      
     0000000000000000000000000000000000000000000000000000000000000000000000000     
    000000000000000000000000000000000000000000000000000000000000000000000000000    
   0000000000000000   000000000000000000000000000000000000000   0000000000000000   
  0000000000000000     0000000000000000000000000000000000000     0000000000000000  
  000000000000000       00000000000000000000000000000000000       000000000000000  
 0000000000000000       00000000000000000000000000000000000       0000000000000000 
 0000000000000000       00000000000000000000000000000000000       0000000000000000 
00000000000000000       00000000000000000000000000000000000       00000000000000000
00000000000000000       00000000000000000000000000000000000       00000000000000000
000000000000000000     0000000000000000000000000000000000000     000000000000000000
00000000000000000000000000000000000000000000000000000000000000000000000000000000000
00000000000000000000000000000000000000000000000000000000000000000000000000000000000
00000000000000000000000000000000000000000000000000000000000000000000000000000000000
00000000000000000000000000000000000000000000000000000000000000000000000000000000000
 000000000000000000000000000000000000000000000000000000000000000000000000000000000 
 0000000000000000                                                 0000000000000000 
  000000000000000                                                 000000000000000  
  0000000000000000                                               0000000000000000  
   00000000000000000                                           00000000000000000   
    00000000000000000                                         00000000000000000    
     000000000000000000                                     000000000000000000     
      0000000000000000000                                 0000000000000000000      
        00000000000000000000                           00000000000000000000        
         00000000000000000000000                   00000000000000000000000         
           0000000000000000000000000000000000000000000000000000000000000           
             000000000000000000000000000000000000000000000000000000000             
               00000000000000000000000000000000000000000000000000000               
                 0000000000000000000000000000000000000000000000000                 
                    0000000000000000000000000000000000000000000                    
                        00000000000000000000000000000000000                        
                             0000000000000000000000000                             """
SMILE_W, SMILE_H = 83, 41
_FG = [(0xC5, 0x0F, 0x1F), (0x88, 0x17, 0x98), (0xFC, 0x7F, 0x11)]
_BG = [(0xC1, 0x9C, 0x10), (0x3A, 0x96, 0xDD), (0x01, 0x24, 0x86)]


def _mul(c, a):  # pixel.RGBA.multiply (pixel.zig:483-491)
    return tuple(v * a // 255 for v in c) + (a,)


def _smile(z, sfc, fg, bg):
    w, h = SMILE_W * 2 + 10, SMILE_H * 2 + 10
    for (x0, y0), f, b in zip([(2, 3), (w // 2 + 3, 3), (w // 4 + 2, h // 2 + 3)], fg, bg):
        for row, line in enumerate(SMILE.split("\n")):
            for col, ch in enumerate(line):
                sfc.put_pixel(x0 + col, y0 + row, f if ch == "0" else b)
    return sfc


@compositor_scene("001_smile_rgb")
def s001(z):
    sfc = z.Surface(Format.rgb, SMILE_W * 2 + 10, SMILE_H * 2 + 10)
    return _smile(z, sfc, [z.Pixel.rgb(*c) for c in _FG], [z.Pixel.rgb(*c) for c in _BG])


@compositor_scene("002_smile_rgba")
def s002(z):
    sfc = z.Surface(Format.rgba, SMILE_W * 2 + 10, SMILE_H * 2 + 10)
    return _smile(z, sfc, [z.Pixel.rgba(*_mul(c, 0xFF)) for c in _FG], [z.Pixel.rgba(*_mul(c, 0x99)) for c in _BG])


@compositor_scene("023_smile_alpha_mask")
def s023(z):
    w, h = SMILE_W * 2 + 10, SMILE_H * 2 + 10
    fg = [z.Pixel.rgba(*_mul(c, 0xFF)) for c in _FG]
    bg = [z.Pixel.rgba(*_mul(c, 0x99)) for c in _BG]
    result = z.Surface(Format.rgba, w, h)
    mask = z.Surface(Format.alpha8, SMILE_W, SMILE_H)
    for row, line in enumerate(SMILE.split("\n")):
        for col, ch in enumerate(line):
            mask.put_pixel(col, row, z.Pixel.alpha8(255 if ch == "0" else 0))
    bsfc = z.SurfacePixel(bg[0], SMILE_W, SMILE_H)
    fsfc = z.SurfacePixel(fg[0], SMILE_W, SMILE_H)
    for i, (x, y) in enumerate([(12, 13), (w // 2 - 7, 13), (w // 4 + 2, h // 2 - 7)]):
        if i > 0:
            bsfc.paint_pixel(bg[i])
            fsfc.paint_pixel(fg[i])
        fsfc.composite(mask, Operator.dst_in, 0, 0)
        bsfc.composite(fsfc, Operator.src_over, 0, 0)
        result.composite(bsfc, Operator.src_over, x, y)
    return result


def _tri(c, w, h, m):
    c.move_to(0 + m, 0 + m)
    c.line_to(w - m - 1, 0 + m)
    c.line_to(w // 2 - 1, h - m - 1)
    c.close_path()


def _alpha_tri(stem, make_sfc, src):
    @path_scene(stem)
    def scene(z, aa):
        sfc = make_sfc(z)
        c = z.Context(sfc)
        c.set_source_to_pixel(src(z))
        c.set_anti_aliasing_mode(aa)
        _tri(c, 300, 300, 10)
        c.fill()
        return sfc
    return scene


_alpha_tri("046_fill_triangle_alpha", lambda z: z.SurfacePixel(z.Pixel.rgb(0xFF, 0xFF, 0xFF), 300, 300), lambda z: z.Pixel.alpha8(255))
_alpha_tri("047_fill_triangle_alpha_gray", lambda z: z.Surface(Format.alpha8, 300, 300), lambda z: z.Pixel.alpha8(119))
_alpha_tri("049_fill_triangle_alpha4_gray", lambda z: z.Surface(Format.alpha4, 300, 300), lambda z: z.Pixel.alpha4(7))
_alpha_tri("050_fill_triangle_alpha2_gray", lambda z: z.Surface(Format.alpha2, 300, 300), lambda z: z.Pixel.alpha2(2))
_alpha_tri("051_fill_triangle_alpha1_gray", lambda z: z.Surface(Format.alpha1, 300, 300), lambda z: z.Pixel.alpha1(1))
_alpha_tri("052_fill_triangle_alpha4_gray_scaledown", lambda z: z.Surface(Format.alpha4, 300, 300), lambda z: z.Pixel.alpha8(119))
_alpha_tri("053_fill_triangle_alpha8_gray_scaleup", lambda z: z.Surface(Format.alpha8, 300, 300), lambda z: z.Pixel.alpha4(7))


@path_scene("048_fill_triangle_static")
def s048(z, aa):
    sfc = z.Surface(Format.rgb, 300, 300)
    p = z.Path()
    _tri(p, 300, 300, 10)  # StaticPath: same node construction (static_path.zig)
    z.painter.fill(sfc, z.Pattern.opaque(z.Pixel.rgb(0xFF, 0xFF, 0xFF)), p.nodes, z.FillOptions(anti_aliasing_mode=aa))
    return sfc


def _rgb3(g):
    g.add_stop(0, {"rgb": (1, 0, 0)})
    g.add_stop(0.5, {"rgb": (0, 1, 0)})
    g.add_stop(1, {"rgb": (0, 0, 1)})
    return g


def _run_gradient(z, scratch, g):
    z.SurfaceCompositor.run(scratch, 0, 0, [z.Operation(Operator.src_over, src=z.Param.gradient(g))])


@compositor_scene("061_linear_gradient")
def s061(z):
    dst = z.Surface(Format.rgb, 200, 400)
    scratch = z.Surface(Format.rgb, 100, 100)
    for sx, sy, x0, y0, x1, y1 in [(0, 0, 0, 49, 99, 49), (100, 0, 99, 49, 0, 49), (0, 100, 49, 0, 49, 99),
                                   (100, 100, 49, 99, 49, 0), (0, 200, 0, 0, 99, 99), (100, 200, 99, 0, 0, 99),
                                   (0, 300, 0, 99, 99, 0), (100, 300, 99, 99, 0, 0)]:
        _run_gradient(z, scratch, _rgb3(z.Gradient.linear(x0, y0, x1, y1)))
        dst.composite(scratch, Operator.src_over, sx, sy)
    return dst


@compositor_scene("062_hsl_gradient")
def s062(z):
    dst = z.Surface(Format.rgb, 100, 200)
    for w, h, sx, sy, x0, y0, x1, y1, c0, c1 in [
        (100, 100, 0, 0, 0, 49, 99, 49, {"hsl": (300, 1, 0.5)}, {"hsl": (60, 1, 0.5)}),
        (50, 100, 0, 100, 0, 49, 49, 49, {"hsl": (0, 1, 0.5)}, {"hsl": (0, 0, 0.5)}),
        (50, 100, 50, 100, 49, 49, 0, 49, {"hsl": (180, 1, 0.5)}, {"hsl": (180, 0, 0.5)}),
        (100, 50, 0, 100, 49, 0, 49, 49, {"hsla": (0, 0, 1, 1)}, {"hsla": (0, 0, 0.5, 0)}),
        (100, 50, 0, 150, 49, 49, 49, 0, {"hsla": (0, 0, 0, 1)}, {"hsla": (0, 0, 0.5, 0)}),
    ]:
        scratch = z.Surface(Format.rgba, w, h)
        g = z.Gradient.linear(x0, y0, x1, y1, Interp.hsl, Polar.shorter)
        g.add_stop(0, c0)
        g.add_stop(1, c1)
        _run_gradient(z, scratch, g)
        dst.composite(scratch, Operator.src_over, sx, sy)
    return dst


@compositor_scene("063_radial_gradient")
def s063(z):
    dst = z.Surface(Format.rgba, 300, 500)
    cases = []
    for y in range(3):
        for x in range(3):
            cases.append((100 * x, 100 * y, 49 + 25 * (float(x) - 1), 49 + 25 * (float(y) - 1), 5, 49, 49, 50))
    cases += [(0, 300, 49, 49, 0, 49, 49, 50), (100, 300, 49, 49, 50, 49, 49, 0), (200, 300, 49, 49, 50, 49, 49, 50),
              (0, 400, 49, 49, 0, 49, 49, 0), (100, 400, 10, 49, 0, 49, 49, 25), (200, 400, 49, 49, 49, 49, 49, 50)]
    for sx, sy, ix, iy, ir, ox, oy, orad in cases:
        scratch = z.Surface(Format.rgba, 100, 100)
        _run_gradient(z, scratch, _rgb3(z.Gradient.radial(ix, iy, ir, ox, oy, orad)))
        dst.composite(scratch, Operator.src_over, sx, sy)
    return dst


@compositor_scene("065_conic_gradient")
def s065(z):
    dst = z.Surface(Format.rgba, 200, 200)
    for y in range(2):
        for x in range(2):
            scratch = z.Surface(Format.rgba, 100, 100)
            g = z.Gradient.conic(49, 49, math.pi / 2.0 * float(y * 2 + x), Interp.hsl, Polar.increasing)
            g.add_stop(0, {"hsl": (0, 1, 0.5)})
            g.add_stop(1, {"hsl": (360, 1, 0.5)})
            _run_gradient(z, scratch, g)
            dst.composite(scratch, Operator.src_over, 100 * x, 100 * y)
    return dst


@compositor_scene("068_gradient_deband")
def s068(z):
    dst = z.SurfacePixel(z.Pixel.from_color({"rgb": (1, 1, 1)}).to_rgb(), 500, 500)
    for sx, sy, gray, bits, dither in [(0, 0, False, 8, DitherType.none), (0, 100, False, 8, DitherType.bayer),
                                       (0, 200, True, 8, DitherType.none), (0, 300, True, 4, DitherType.none),
                                       (0, 400, True, 4, DitherType.bayer)]:
        fmt = (Format.alpha4 if bits == 4 else Format.alpha8) if gray else Format.rgb
        scratch = z.Surface(fmt, 500, 100)
        g = z.Gradient.linear(0, 50, 500, 50)
        if gray:
            g.add_stop(0, {"rgba": (1, 1, 1, 0)})
            g.add_stop(1, {"rgba": (0, 0, 0, 1)})
        else:
            g.add_stop(0, {"rgb": (27.0 / 255.0, 93.0 / 255.0, 124.0 / 255.0)})
            g.add_stop(1, {"rgb": (38.0 / 255.0, 32.0 / 255.0, 16.0 / 255.0)})
        z.SurfaceCompositor.run(scratch, 0, 0, [z.Operation(Operator.src_over, src=z.Param.dither(z.Dither(dither, g, bits)))])
        dst.composite(scratch, Operator.src_over, sx, sy)
    return dst


def _gamma_scene(z):
    width, height = 400, 300
    sfc = z.Surface(Format.rgb, width, height)
    c = z.Context(sfc)
    c.set_anti_aliasing_mode(AntiAliasMode.none)

    def rect(w, h):
        c.move_to(0, 0)
        c.line_to(w, 0)
        c.line_to(w, h)
        c.line_to(0, h)
        c.close_path()
        c.set_identity()
        c.fill()
        c.reset_path()

    for pos, prof in [(0.0, "rgb"), (1.0, "srgb")]:
        c.set_source_to_pixel(z.Pixel.from_color({prof: (0.3, 0.3, 0.3)}))
        c.translate(0, pos * height / 2)
        rect(width, height // 2)
    for pos, prof, method in [(0.0, "rgb", Interp.linear_rgb), (1.0, "srgb", Interp.linear_rgb),
                              (2.0, "rgb", Interp.srgb), (3.0, "srgb", Interp.srgb)]:
        offset = 30.0
        gw = width - offset * 2
        gh = (height - offset * 2) / 4
        g = z.Gradient.linear(0, gh / 2, gw, gh / 2, method)
        g.add_stop(0, {prof: (0.80, 0, 0)})
        g.add_stop(0.5, {prof: (0, 0.80, 0)})
        g.add_stop(1, {prof: (0, 0, 0.80)})
        c.translate(offset, offset + gh * pos)
        c.set_source(z.Pattern.gradient(g))
        rect(gw, gh)
        c.set_source_to_pixel(z.Pixel.rgb(0, 0, 0))
    return sfc


compositor_scene("071_gamma_linear")(_gamma_scene)
compositor_scene("072_gamma_srgb", profile="srgb")(_gamma_scene)


@path_scene("059_stroke_star_gradient")
def s059(z, aa):
    w = h = 300
    sfc = z.Surface(Format.rgb, w, h)
    c = z.Context(sfc)
    c.set_anti_aliasing_mode(aa)
    c.set_line_width(10)
    c.set_line_join_mode(JoinMode.round)
    m, xs, ys = 20, 3, 5
    g = _rgb3(z.Gradient.linear(0 + m * 3, h // 2, w - m * 3, h // 2))
    c.set_source(z.Pattern.gradient(g))
    c.move_to(w // 2, 0 + m)
    c.line_to(w - m * xs - 1, h - m - 1)
    c.line_to(0 + m, 0 + m * ys)
    c.line_to(w - m - 1, 0 + m * ys)
    c.line_to(0 + m * xs, h - m - 1)
    c.close_path()
    c.stroke()
    return sfc


def _rect(c, w, h):
    c.move_to(0, 0)
    c.line_to(w, 0)
    c.line_to(w, h)
    c.line_to(0, h)
    c.close_path()


@path_scene("064_radial_source")
def s064(z, aa):
    sfc = z.Surface(Format.rgb, 100, 100)
    g = _rgb3(z.Gradient.radial(49, 49, 0, 49, 49, 50))
    c = z.Context(sfc)
    c.set_anti_aliasing_mode(aa)
    c.set_source(z.Pattern.gradient(g))
    _rect(c, 100, 100)
    c.fill()
    return sfc


@path_scene("066_conic_pie_gradient")
def s066(z, aa):
    sfc = z.Surface(Format.rgb, 300, 300)
    g = z.Gradient.conic(149, 149, 0)
    g.add_stop(0, {"rgb": (1, 0, 0)})
    g.add_stop(1.0 / 3.0, {"rgb": (1, 0, 0)})
    g.add_stop(1.0 / 3.0 + 0.005, {"rgb": (0, 1, 0)})
    g.add_stop(2.0 / 3.0, {"rgb": (0, 1, 0)})
    g.add_stop(2.0 / 3.0 + 0.005, {"rgb": (0, 0, 1)})
    g.add_stop(1, {"rgb": (0, 0, 1)})
    c = z.Context(sfc)
    c.set_anti_aliasing_mode(aa)
    c.set_source(z.Pattern.gradient(g))
    c.arc(149, 149, 100, 0, math.pi * 2)
    c.close_path()
    c.fill()
    return sfc


@path_scene("067_gradient_transforms")
def s067(z, aa):
    sfc = z.Surface(Format.rgb, 200, 200)
    linear = _rgb3(z.Gradient.linear(0, 0, 50, 50))
    c = z.Context(sfc)
    c.scale(2, 2)
    c.set_anti_aliasing_mode(aa)
    c.set_source(z.Pattern.gradient(linear))
    _rect(c, 50, 50)
    c.fill()
    radial = _rgb3(z.Gradient.radial(25, 50, 0, 25, 50, 50))
    c.set_identity()
    skew = z.Transformation()
    skew.by = 0.5
    c.mul(skew)
    c.translate(100, 0)
    c.set_source(z.Pattern.gradient(radial))
    c.set_identity()
    c.translate(100, 0)
    c.reset_path()
    _rect(c, 100, 100)
    c.fill()
    conic = z.Gradient.conic(50, 50, 0, Interp.hsl, Polar.increasing)
    conic.add_stop(0, {"hsl": (0, 1, 0.5)})
    conic.add_stop(1, {"hsl": (360, 1, 0.5)})
    c.set_identity()
    c.scale(2, 1)
    c.translate(0, 100)
    c.set_source(z.Pattern.gradient(conic))
    c.reset_path()
    _rect(c, 100, 100)
    c.fill()
    return sfc


@path_scene("069_gradient_dither_context")
def s069(z, aa):
    dst = z.Surface(Format.alpha4, 400, 150)
    g = z.Gradient.linear(0, 25, 400, 25)
    g.add_stop(0, {"rgba": (0, 0, 0, 0)})
    g.add_stop(1, {"rgba": (1, 1, 1, 1)})
    c = z.Context(dst)
    c.set_anti_aliasing_mode(aa)
    c.set_source(z.Pattern.gradient(g))
    _rect(c, 400, 50)
    c.fill()
    c.reset_path()
    c.translate(0, 50)
    c.set_source(z.Pattern.gradient(g))
    c.set_dither(DitherType.bayer)
    _rect(c, 400, 50)
    c.fill()
    c.reset_path()
    c.set_identity()
    c.translate(0, 100)
    c.set_source(z.Pattern.gradient(g))
    c.set_dither(DitherType.blue_noise)
    _rect(c, 400, 50)
    c.fill()
    return dst


@path_scene("070_compositor_ops")
def s070(z, aa):
    width, height = 460, 3090
    sfc = z.SurfacePixel(z.Pixel.from_color({"rgb": (1, 1, 1)}).to_rgb(), width, height)

    def draw(sx, sy, op, precision, transparent):
        bg = z.Pixel.from_color({"rgba": (0.69, 0.23, 0.21, 0.9)} if transparent else {"rgb": (0.69, 0.23, 0.21)})
        fg = z.Pixel.from_color({"rgba": (0.56, 0.50, 0.89, 0.8)} if transparent else {"rgb": (0.56, 0.50, 0.89)})
        scratch = z.Surface(Format.rgba, 100, 100)
        c = z.Context(scratch)
        c.set_anti_aliasing_mode(aa)
        c.set_precision(precision)
        c.set_source_to_pixel(bg)
        _rect(c, 75, 75)
        c.fill()
        c.set_source_to_pixel(fg)
        c.set_operator(op)
        c.reset_path()
        c.translate(25, 25)
        _rect(c, 75, 75)
        c.fill()
        sfc.composite(scratch, Operator.src_over, sx, sy)

    for i, op in enumerate(Operator):
        for j in range(2):
            draw(10 + 110 * j, 10 + 110 * i, op, Precision.integer, bool(j))
        for j in range(2, 4):
            draw(10 + 110 * j, 10 + 110 * i, op, Precision.float, bool(j % 2))
    return sfc
