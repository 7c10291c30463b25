"""Host-side mirror of z2d's API for the fill/stroke/composite path.

Same names, argument meaning and error behaviour as the reference's Zig API
(`Context`, `Path`, `Transformation`, `Pixel`, `Gradient`, `Pattern`,
`Surface`, `painter.fill/stroke`, `compositor.SurfaceCompositor.run`), so the
parity tests read like the reference's own spec scenes.  Everything here is
host logic only (path construction, option marshalling into the C-ABI PODs of
include/z2d_cuda.h); pixels live behind a *backend*:

* the product backend is `z2d_b200.cuda_backend.CudaBackend` (libz2d_cuda.so,
  device-resident surfaces) -- the default, and it raises if the library is
  missing;
* tests inject `tests/oracle_backend.OracleBackend` (the CPU restatement) to
  run the very same scene code against the oracle.

Reference citations are file:line in the z2d tree.
"""
import ctypes as C
import math

import numpy as np

from . import abi
from .abi import (AntiAliasMode, CapMode, ColorSpace, DitherSource, DitherType, FillRule, Format, GradientType, Interp,
                  JoinMode, NodeTag, Operator, ParamKind, PatternKind, Polar, Precision)

f32 = np.float32
default_tolerance = 0.1  # options.zig:10


# --------------------------------------------------------------------------
# Transformation.zig
class Transformation:
    __slots__ = ("ax", "by", "cx", "dy", "tx", "ty")

    def __init__(self, ax=1.0, by=0.0, cx=0.0, dy=1.0, tx=0.0, ty=0.0):
        self.ax, self.by, self.cx, self.dy, self.tx, self.ty = float(ax), float(by), float(cx), float(dy), float(tx), float(ty)

    @staticmethod
    def identity():
        return Transformation()

    def as_tuple(self):
        return (self.ax, self.by, self.cx, self.dy, self.tx, self.ty)

    def equal(self, o):
        return self.as_tuple() == o.as_tuple()

    def determinant(self):
        return self.ax * self.dy - self.by * self.cx

    def mul(a, b):  # Transformation.zig:58-78
        return Transformation(a.ax * b.ax + a.by * b.cx, a.ax * b.by + a.by * b.dy, a.cx * b.ax + a.dy * b.cx,
                              a.cx * b.by + a.dy * b.dy, a.ax * b.tx + a.by * b.ty + a.tx,
                              a.cx * b.tx + a.dy * b.ty + a.ty)

    def inverse(a):  # Transformation.zig:103-162
        if a.by == 0 and a.cx == 0:
            if a.ax == 0 or a.dy == 0:
                raise abi.InvalidMatrix()
            if a.ax != 1 or a.dy != 1:
                return Transformation(1 / a.ax, 0, 0, 1 / a.dy, -a.tx / a.ax, -a.ty / a.dy)
            return Transformation(1, 0, 0, 1, -a.tx, -a.ty)
        det = a.determinant()
        if det == 0:
            raise abi.InvalidMatrix()
        k = 1 / det
        return Transformation(a.dy * k, -a.by * k, -a.cx * k, a.ax * k, (a.by * a.ty - a.dy * a.tx) * k,
                              (a.cx * a.tx - a.ax * a.ty) * k)

    def translate(a, tx, ty):
        return a.mul(Transformation(tx=tx, ty=ty))

    def scale(a, sx, sy):
        return a.mul(Transformation(ax=sx, dy=sy))

    def rotate(a, angle):
        s, c = math.sin(angle), math.cos(angle)
        return a.mul(Transformation(c, -s, s, c, 0, 0))

    def user_to_device_distance(a, x, y):
        return a.ax * x + a.by * y, a.cx * x + a.dy * y

    def user_to_device(a, x, y):
        dx, dy = a.user_to_device_distance(x, y)
        return dx + a.tx, dy + a.ty

    def device_to_user(a, x, y):
        return a.inverse().user_to_device(x, y)

    def device_to_user_distance(a, x, y):
        return a.inverse().user_to_device_distance(x, y)


# --------------------------------------------------------------------------
# internal/arc.zig
def _hypot(a, b):
    return math.hypot(a, b)


def transformed_circle_major_axis(m, radius):  # arc.zig:92-267
    eps = 0.00390625
    det = m.ax * m.dy - m.by * m.cx
    if abs(det * det - 1.0) < eps:
        if abs(m.by) < eps and abs(m.cx) < eps:
            return radius
        if abs(m.ax) < eps and abs(m.dy) < eps:
            return radius
    i = m.ax * m.ax + m.by * m.by
    j = m.cx * m.cx + m.dy * m.dy
    f = 0.5 * (i + j)
    g = 0.5 * (i - j)
    h = m.ax * m.cx + m.by * m.dy
    return radius * math.sqrt(f + _hypot(g, h))


_ARC_TABLE = [(math.pi / 1.0, 0.0185185185185185036127), (math.pi / 2.0, 0.000272567143730179811158),
              (math.pi / 3.0, 2.38647043651461047433e-05), (math.pi / 4.0, 4.2455377443222443279e-06),
              (math.pi / 5.0, 1.11281001494389081528e-06), (math.pi / 6.0, 3.72662000942734705475e-07),
              (math.pi / 7.0, 1.47783685574284411325e-07), (math.pi / 8.0, 6.63240432022601149057e-08),
              (math.pi / 9.0, 3.2715520137536980553e-08), (math.pi / 10.0, 1.73863223499021216974e-08),
              (math.pi / 11.0, 9.81410988043554039085e-09)]


def _arc_error_normalized(angle):  # arc.zig:40-42
    return 2.0 / 27.0 * math.pow(math.sin(angle / 4), 6) / math.pow(math.cos(angle / 4), 2)


def _arc_max_angle(tolerance):  # arc.zig:44-83
    for ang, err in _ARC_TABLE:
        if err < tolerance:
            return ang
    angle = None
    for i in range(len(_ARC_TABLE), 1000):
        angle = math.pi / float(i)
        if _arc_error_normalized(angle) <= tolerance:
            break
    return angle


def _arc_segments_needed(angle, radius, ctm, tolerance):  # arc.zig:85-91
    major = transformed_circle_major_axis(ctm, radius)
    return int(math.ceil(abs(angle) / _arc_max_angle(tolerance / major)))


def _arc_segment(path, xc, yc, radius, a, b):  # arc.zig:294-318
    rsa, rca = radius * math.sin(a), radius * math.cos(a)
    rsb, rcb = radius * math.sin(b), radius * math.cos(b)
    h = 4.0 / 3.0 * math.tan((b - a) / 4.0)
    path.curve_to(xc + rca - h * rsa, yc + rsa + h * rca, xc + rcb + h * rsb, yc + rsb - h * rcb, xc + rcb, yc + rsb)


def _arc_in_direction(path, xc, yc, radius, amin, amax, forward, ctm, tolerance):  # arc.zig:330-392
    if not (amax * amax >= 0.0) or not (amin * amin >= 0.0):
        return
    max_full = 65536
    if amax - amin > 2 * math.pi * max_full:
        amax = math.fmod(amax - amin, 2 * math.pi)
        amin = math.fmod(amin, 2 * math.pi)
        if amin < 0:
            amin += 2 * math.pi
        amax += amin + 2 * math.pi * max_full
    if amax - amin > math.pi:
        amid = amin + (amax - amin) / 2.0
        if forward:
            _arc_in_direction(path, xc, yc, radius, amin, amid, forward, ctm, tolerance)
            _arc_in_direction(path, xc, yc, radius, amid, amax, forward, ctm, tolerance)
        else:
            _arc_in_direction(path, xc, yc, radius, amid, amax, forward, ctm, tolerance)
            _arc_in_direction(path, xc, yc, radius, amin, amid, forward, ctm, tolerance)
    elif amax != amin:
        segments = _arc_segments_needed(amax - amin, radius, ctm, tolerance)
        step = (amax - amin) / float(segments)
        segments -= 1
        if not forward:
            amin, amax = amax, amin
            step = -step
        path._arc_line_to(xc + radius * math.cos(amin), yc + radius * math.sin(amin))
        for _ in range(max(0, segments)):
            _arc_segment(path, xc, yc, radius, amin, amin + step)
            amin += step
        _arc_segment(path, xc, yc, radius, amin, amax)
    else:
        path._arc_line_to(xc + radius * math.cos(amin), yc + radius * math.sin(amin))


# --------------------------------------------------------------------------
# Path.zig -- nodes are stored in DEVICE space as (tag, p0..p5) tuples
_I24_MIN, _I24_MAX = float(-(1 << 23)), float((1 << 23) - 1)


def _clamp_i24(x):
    return max(_I24_MIN, min(float(x), _I24_MAX))


class Path:
    def __init__(self):
        self.nodes = []
        self.initial_point = None
        self.current_point = None
        self.tolerance = default_tolerance
        self.transformation = Transformation()

    def reset(self):
        self.nodes = []
        self.initial_point = None
        self.current_point = None

    def _pt(self, x, y):
        return self.transformation.user_to_device(_clamp_i24(x), _clamp_i24(y))

    def move_to(self, x, y):  # Path.zig:124-145
        p = self._pt(x, y)
        if self.nodes and self.nodes[-1][0] == NodeTag.move_to and self.nodes[-1][1:3] == p:
            return
        self.nodes.append((NodeTag.move_to, p[0], p[1], 0.0, 0.0, 0.0, 0.0))
        self.initial_point = p
        self.current_point = p

    def rel_move_to(self, x, y):
        if self.current_point is None:
            raise ValueError("NoCurrentPoint")
        ux, uy = self.transformation.device_to_user(*self.current_point)
        self.move_to(ux + x, uy + y)

    def line_to(self, x, y):  # Path.zig:175-183
        if self.current_point is None:
            return self.move_to(x, y)
        p = self._pt(x, y)
        self.nodes.append((NodeTag.line_to, p[0], p[1], 0.0, 0.0, 0.0, 0.0))
        self.current_point = p

    def rel_line_to(self, x, y):
        if self.current_point is None:
            raise ValueError("NoCurrentPoint")
        ux, uy = self.transformation.device_to_user(*self.current_point)
        self.line_to(ux + x, uy + y)

    def curve_to(self, x1, y1, x2, y2, x3, y3):  # Path.zig:237-260
        if self.current_point is None:
            raise ValueError("NoCurrentPoint")
        p1, p2, p3 = self._pt(x1, y1), self._pt(x2, y2), self._pt(x3, y3)
        self.nodes.append((NodeTag.curve_to, p1[0], p1[1], p2[0], p2[1], p3[0], p3[1]))
        self.current_point = p3

    def rel_curve_to(self, x1, y1, x2, y2, x3, y3):
        if self.current_point is None:
            raise ValueError("NoCurrentPoint")
        ux, uy = self.transformation.device_to_user(*self.current_point)
        self.curve_to(ux + x1, uy + y1, ux + x2, uy + y2, ux + x3, uy + y3)

    def _arc_line_to(self, x, y):  # Path.zig:365-384 (compares untransformed x,y with the device-space current point)
        if self.current_point is not None and self.current_point[0] == x and self.current_point[1] == y:
            return
        self.line_to(x, y)

    def arc(self, xc, yc, radius, angle1, angle2):  # Path.zig:304-332
        a2 = angle2
        while a2 < angle1:
            a2 += math.pi * 2
        _arc_in_direction(self, xc, yc, radius, angle1, a2, True, self.transformation, max(self.tolerance, 0.001))

    def arc_negative(self, xc, yc, radius, angle1, angle2):  # Path.zig:334-362
        a2 = angle2
        while a2 > angle1:
            a2 -= math.pi * 2
        _arc_in_direction(self, xc, yc, radius, a2, angle1, False, self.transformation, max(self.tolerance, 0.001))

    def close(self):  # Path.zig:453-476: close_path + explicit move_to(initial point); current point is NOT updated
        if self.current_point is None:
            return
        self.nodes.append((NodeTag.close_path, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0))
        ip = self.initial_point
        self.nodes.append((NodeTag.move_to, ip[0], ip[1], 0.0, 0.0, 0.0, 0.0))

    close_path = close

    def is_closed(self):
        return is_closed_node_set(self.nodes)


def is_closed_node_set(nodes):  # path_nodes.zig:23-37
    if not nodes:
        return False
    closed = False
    for i, n in enumerate(nodes):
        if n[0] == NodeTag.move_to:
            if not closed and i != 0:
                break
        elif n[0] == NodeTag.close_path:
            closed = True
        else:
            closed = False
    return closed


def nodes_to_array(nodes):
    """[]PathNode -> z2d_node[] (numpy structured view + ctypes pointer)."""
    arr = (abi.Node * max(1, len(nodes)))()
    for i, n in enumerate(nodes):
        arr[i].tag = int(n[0])
        for k in range(6):
            arr[i].p[k] = n[1 + k]
    return arr


# --------------------------------------------------------------------------
# color.zig / pixel.zig (host side only: Color.init + Pixel.fromColor)
def _clamp01(v):
    return f32(max(0.0, min(float(f32(v)), 1.0)))


class Color:
    """color.Color: de-multiplied colour in linear RGB, sRGB (gamma 2.2) or HSL."""

    def __init__(self, space, c):
        self.space = ColorSpace(space)
        self.c = tuple(f32(v) for v in c)

    @staticmethod
    def init(args):  # color.zig:58-69: {"rgb": (..)} etc.
        if isinstance(args, Color):
            return args
        (kind, v), = args.items()
        v = tuple(v)
        if kind in ("rgb", "rgba"):
            a = v[3] if kind == "rgba" else 1
            return Color(ColorSpace.linear_rgb, (_clamp01(v[0]), _clamp01(v[1]), _clamp01(v[2]), _clamp01(a)))
        if kind in ("srgb", "srgba"):
            a = v[3] if kind == "srgba" else 1
            return Color(ColorSpace.srgb, (_clamp01(v[0]), _clamp01(v[1]), _clamp01(v[2]), _clamp01(a)))
        if kind in ("hsl", "hsla"):
            a = v[3] if kind == "hsla" else 1
            h = f32(v[0])
            if h < 0 or h > 360:  # color.zig:392-394, Zig @mod (floored)
                h = f32(math.fmod(float(h), 360.0))
                if h < 0:
                    h = f32(h + f32(360))
            return Color(ColorSpace.hsl, (h, _clamp01(v[1]), _clamp01(v[2]), _clamp01(a)))
        raise ValueError(kind)

    def to_linear(self):  # LinearRGB.fromColor (color.zig:166-195), all in f32
        c = self.c
        if self.space == ColorSpace.linear_rgb:
            return c
        if self.space == ColorSpace.srgb:
            g = f32(2.2)
            return (np.power(c[0], g), np.power(c[1], g), np.power(c[2], g), c[3])
        return _hsl_to_rgb(c)

    def pod(self):
        p = abi.ColorPOD()
        p.space = int(self.space)
        for i in range(4):
            p.c[i] = float(self.c[i])
        return p


def _hsl_to_rgb(c):  # color.zig:459-478
    h, s, l, a = c
    hue = f32(math.fmod(float(h), 360.0))
    if hue < 0:
        hue = f32(hue + f32(360))

    def ch(n):
        k = f32(math.fmod(float(f32(n) + hue / f32(30)), 12.0))
        aa = s * min(l, f32(1) - l)
        return l - aa * max(f32(-1), min(k - f32(3), f32(9) - k, f32(1)))

    return (ch(0), ch(8), ch(4), a)


def _zround(v):  # Zig @round: half away from zero
    v = float(v)
    return math.floor(v + 0.5) if v >= 0 else -math.floor(-v + 0.5)


class Pixel:
    """pixel.Pixel: a tagged pixel value (channel values as stored by the format)."""
    __slots__ = ("format", "r", "g", "b", "a")

    def __init__(self, format, r=0, g=0, b=0, a=0):
        self.format = Format(format)
        self.r, self.g, self.b, self.a = int(r), int(g), int(b), int(a)

    @staticmethod
    def rgb(r, g, b):
        return Pixel(Format.rgb, r, g, b, 0)

    @staticmethod
    def xrgb(r, g, b):
        return Pixel(Format.xrgb, r, g, b, 0)

    @staticmethod
    def rgba(r, g, b, a):
        return Pixel(Format.rgba, r, g, b, a)

    @staticmethod
    def argb(r, g, b, a):
        return Pixel(Format.argb, r, g, b, a)

    @staticmethod
    def alpha8(a):
        return Pixel(Format.alpha8, 0, 0, 0, a)

    @staticmethod
    def alpha4(a):
        return Pixel(Format.alpha4, 0, 0, 0, a)

    @staticmethod
    def alpha2(a):
        return Pixel(Format.alpha2, 0, 0, 0, a)

    @staticmethod
    def alpha1(a):
        return Pixel(Format.alpha1, 0, 0, 0, a)

    @staticmethod
    def from_color(args):  # pixel.zig:115-117 + color.zig:214-232 (encodeRGBA: round, then integer multiply)
        lin = Color.init(args).to_linear()
        r, g, b, a = (int(_zround(f32(255.0) * v)) for v in lin)
        return Pixel(Format.rgba, r * a // 255, g * a // 255, b * a // 255, a)

    def to_rgb(self):
        """pixel.RGB.fromPixel (pixel.zig:399-433)."""
        if self.format in (Format.alpha8, Format.alpha4, Format.alpha2, Format.alpha1):
            return Pixel(Format.rgb, 0, 0, 0, 0)
        return Pixel(Format.rgb, self.r, self.g, self.b, 0)

    def pod(self):
        return abi.PixelPOD(int(self.format), self.r, self.g, self.b, self.a)


# --------------------------------------------------------------------------
# gradient.zig
class Gradient:
    def __init__(self, type, geom, method=Interp.linear_rgb, polar=Polar.shorter):
        self.type = GradientType(type)
        self.geom = tuple(float(v) for v in geom) + (0.0,) * (6 - len(geom))
        self.method = Interp(method)
        self.polar = Polar(polar)
        self.stops = []  # (offset f32, Color, idx)
        self._idx = 0
        self.transformation = Transformation()  # stored INVERTED (gradient.zig:201-203)
        self._keep = None

    @staticmethod
    def linear(x0, y0, x1, y1, method=Interp.linear_rgb, polar=Polar.shorter):
        return Gradient(GradientType.linear, (x0, y0, x1, y1), method, polar)

    @staticmethod
    def radial(ix, iy, ir, ox, oy, orad, method=Interp.linear_rgb, polar=Polar.shorter):
        return Gradient(GradientType.radial, (ix, iy, ir, ox, oy, orad), method, polar)

    @staticmethod
    def conic(x, y, angle, method=Interp.linear_rgb, polar=Polar.shorter):
        return Gradient(GradientType.conic, (x, y, angle), method, polar)

    def add_stop(self, offset, color):  # gradient.zig:797-811
        off = f32(max(0.0, min(float(f32(offset)), 1.0)))
        self.stops.append((off, Color.init(color), self._idx))
        self.stops.sort(key=lambda s: (float(s[0]), s[2]))
        self._idx += 1

    def set_transformation(self, tr):
        self.transformation = tr.inverse()

    def pod(self):
        g = abi.GradientPOD()
        g.type, g.method, g.polar, g.n_stops = int(self.type), int(self.method), int(self.polar), len(self.stops)
        for i in range(6):
            g.geom[i] = self.geom[i]
        for i, v in enumerate(self.transformation.as_tuple()):
            g.inv_ctm[i] = v
        stops = (abi.StopPOD * max(1, len(self.stops)))()
        for i, (off, col, _) in enumerate(self.stops):
            stops[i].offset = float(off)
            stops[i].color = col.pod()
        g.stops = C.cast(stops, C.POINTER(abi.StopPOD))
        self._keep = (g, stops)
        return g


class Dither:  # Dither.zig:28-58
    def __init__(self, type, source, scale):
        self.type = DitherType(type)
        self.source = source  # Pixel | Color/InitArgs dict | Gradient
        self.scale = int(scale)


class Pattern:
    """pattern.Pattern (pattern.zig:32-44): opaque pixel | gradient | dither."""

    def __init__(self, kind, value):
        self.kind = PatternKind(kind)
        self.value = value
        self._keep = None

    @staticmethod
    def opaque(px):
        return Pattern(PatternKind.opaque, px)

    @staticmethod
    def gradient(g):
        return Pattern(PatternKind.gradient, g)

    @staticmethod
    def dither(d):
        return Pattern(PatternKind.dither, d)

    def pod(self):
        p = abi.PatternPOD()
        p.kind = int(self.kind)
        keep = []
        if self.kind == PatternKind.opaque:
            p.pixel = self.value.pod()
        elif self.kind == PatternKind.gradient:
            g = self.value.pod()
            keep.append(g)
            p.gradient = C.pointer(g)
        else:
            d = self.value
            p.dither_type, p.dither_scale = int(d.type), d.scale
            if isinstance(d.source, Pixel):
                p.dither_source = int(DitherSource.pixel)
                p.pixel = d.source.pod()
            elif isinstance(d.source, Gradient):
                p.dither_source = int(DitherSource.gradient)
                g = d.source.pod()
                keep.append(g)
                p.gradient = C.pointer(g)
            else:
                p.dither_source = int(DitherSource.color)
                p.dither_color = Color.init(d.source).pod()
        self._keep = keep
        return p


# --------------------------------------------------------------------------
# surface.zig -- pixels live in the backend (device memory for the product)
_default_backend = None


def default_backend():
    """The product backend: libz2d_cuda.so.  Raises if it cannot be loaded."""
    global _default_backend
    if _default_backend is None:
        from .cuda_backend import CudaBackend
        _default_backend = CudaBackend()
    return _default_backend


class Surface:
    def __init__(self, format, width, height, initial_px=None, backend=None, band=None):
        """Surface.init / initPixel (surface.zig:97-157).  band=(y0, rows): hold only rows [y0, y0 + rows) of a canvas
        `height` rows high (z2d_surface_create_band); `self.height` is then the number of rows held."""
        self.backend = backend or default_backend()
        self.format = Format(format)
        if width < 1:
            raise abi.InvalidWidth()
        if height < 1:
            raise abi.InvalidHeight()
        self.width, self.canvas_height = int(width), int(height)
        if band is None:
            self.band_y0, self.height = 0, int(height)
            self.handle = self.backend.surface_create(self.format, self.width, self.height, initial_px)
        else:
            self.band_y0, self.height = int(band[0]), int(band[1])
            self.handle = self.backend.surface_create_band(self.format, self.width, self.canvas_height, self.band_y0, self.height, initial_px)

    @classmethod
    def _wrap(cls, backend, handle, format, width, canvas_height, y0, rows):
        s = cls.__new__(cls)
        s.backend, s.format, s.width, s.canvas_height = backend, Format(format), int(width), int(canvas_height)
        s.band_y0, s.height, s.handle = int(y0), int(rows), handle
        return s

    def band_view(self, y0, rows):
        """Rows [y0, y0 + rows) of this canvas as a band surface over the SAME memory (z2d_surface_band_view)."""
        return Surface._wrap(self.backend, self.backend.surface_band_view(self.handle, int(y0), int(rows)), self.format, self.width, self.height, y0, rows)

    def ipc_export(self):
        """64-byte handle another process of the node opens with Surface.open_peer_band (z2d_surface_ipc_export)."""
        return self.backend.surface_ipc_export(self.handle)

    @staticmethod
    def open_peer_band(handle, format, width, canvas_height, y0, rows, backend=None):
        """Band over a canvas that lives on another process / GPU of the node: writes go over NVLink into that canvas."""
        backend = backend or default_backend()
        return Surface._wrap(backend, backend.surface_open_peer_band(handle, format, width, canvas_height, y0, rows), format, width, canvas_height, y0, rows)

    @staticmethod
    def init(format, width, height, backend=None):
        return Surface(format, width, height, None, backend)

    @staticmethod
    def init_pixel(px, width, height, backend=None):  # surface.zig:128: surface type follows the pixel's format
        return Surface(px.format, width, height, px, backend)

    def deinit(self):
        if self.handle is not None:
            self.backend.surface_destroy(self.handle)
            self.handle = None

    def get_width(self):
        return self.width

    def get_height(self):
        return self.height

    def get_format(self):
        return self.format

    def byte_len(self):
        return abi.surface_byte_len(self.format, self.width, self.height)

    def download(self):
        """Reference `buf` bytes (tightly packed; sub-byte formats bit-contiguous)."""
        return self.backend.surface_download(self.handle, self.byte_len())

    def export(self, profile=None, filter_byte=False):
        """The scanline bytes export_png.writeToPNGFile compresses (export_png.zig:150-373), produced on the device:
        (rows, row_bytes) uint8.  profile: None / "linear" / "srgb" (WriteToPNGFileOptions.color_profile)."""
        return self.backend.surface_export(self.handle, srgb=(profile == "srgb"), filter_byte=filter_byte)

    def write_png(self, filename, profile=None):
        """export_png.writeToPNGFile (export_png.zig:36-60): magic, IHDR, optional gAMA, zlib'd scanlines, IEND.  Only the
        pixel transform runs on the device; chunk framing and deflate stay on the host, as in the reference."""
        import struct
        import zlib

        def chunk(tag, data):
            return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)

        fmt = Format(self.format)
        depth = {Format.alpha4: 4, Format.alpha2: 2, Format.alpha1: 1}.get(fmt, 8)
        color_type = 6 if fmt in (Format.argb, Format.rgba) else 2 if fmt in (Format.xrgb, Format.rgb) else 0
        out = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", self.width, self.height, depth, color_type, 0, 0, 0))
        if profile is not None:  # export_png.zig:118-133
            gamma = np.float32(1) / np.float32(2.2 if profile == "srgb" else 1.0)
            out += chunk(b"gAMA", struct.pack(">I", int(np.float32(gamma * np.float32(100000)))))
        out += chunk(b"IDAT", zlib.compress(self.export(profile, filter_byte=True).tobytes()))
        out += chunk(b"IEND", b"")
        with open(filename, "wb") as f:
            f.write(out)

    def upload(self, data):
        data = np.ascontiguousarray(np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data, dtype=np.uint8)
        assert data.size == self.byte_len()
        self.backend.surface_upload(self.handle, data)

    def paint_pixel(self, px):  # surface.zig:295
        self.backend.surface_paint_pixel(self.handle, px)

    def downsample(self):  # surface.zig:447-490: 4x box average in place, dimensions / 4
        self.width, self.height = self.backend.surface_downsample(self.handle)
        self.canvas_height = self.height

    def put_pixel(self, x, y, px):  # surface.zig:288
        self.backend.surface_put_pixel(self.handle, int(x), int(y), px)

    def get_pixel(self, x, y):  # surface.zig:280: None when (x, y) is outside the surface
        got = self.backend.surface_get_pixel(self.handle, int(x), int(y))
        return None if got is None else Pixel(Format(got[0]), *got[1:])

    def composite(self, src, operator, dst_x, dst_y, precision=Precision.integer):  # surface.zig:225-241
        SurfaceCompositor.run(self, dst_x, dst_y, [Operation(operator, src=Param.surface(src))], precision)

    def pixels(self):
        """Decoded (h, w, 4) uint8 RGBA view of the surface content (alpha-only
        formats decode to a=raw sample, rgb=0) -- convenience for tests."""
        raw = self.download()
        return decode_pixels(raw, self.format, self.width, self.height)


def decode_pixels(raw, fmt, w, h):
    raw = np.asarray(raw, dtype=np.uint8)
    out = np.zeros((h, w, 4), dtype=np.uint8)
    fmt = Format(fmt)
    if fmt in (Format.argb, Format.xrgb, Format.rgb, Format.rgba):
        px = raw.reshape(h, w, 4)
        if fmt == Format.rgba:
            out[:] = px
        elif fmt == Format.rgb:
            out[..., :3] = px[..., :3]
            out[..., 3] = 255
        elif fmt == Format.argb:
            out[..., 0], out[..., 1], out[..., 2], out[..., 3] = px[..., 2], px[..., 1], px[..., 0], px[..., 3]
        else:
            out[..., 0], out[..., 1], out[..., 2] = px[..., 2], px[..., 1], px[..., 0]
            out[..., 3] = 255
        return out
    bits = abi.FORMAT_BITS[fmt]
    if bits == 8:
        out[..., 3] = raw[:w * h].reshape(h, w)
        return out
    allbits = np.unpackbits(raw, bitorder="little")[:w * h * bits].reshape(h * w, bits)
    vals = np.zeros(h * w, dtype=np.uint8)
    for k in range(bits):
        vals |= (allbits[:, k] << k).astype(np.uint8)
    out[..., 3] = vals.reshape(h, w)
    return out


# --------------------------------------------------------------------------
# painter.zig
class FillOptions:  # painter.zig:28-48
    def __init__(self, anti_aliasing_mode=AntiAliasMode.default, fill_rule=FillRule.non_zero, operator=Operator.src_over,
                 precision=Precision.integer, tolerance=default_tolerance):
        self.anti_aliasing_mode, self.fill_rule, self.operator = anti_aliasing_mode, fill_rule, operator
        self.precision, self.tolerance = precision, tolerance

    def pod(self):
        return abi.FillOptsPOD(int(self.anti_aliasing_mode), int(self.fill_rule), int(self.operator), int(self.precision),
                               float(self.tolerance))


class StrokeOptions:  # painter.zig:145-198
    def __init__(self, anti_aliasing_mode=AntiAliasMode.default, dashes=(), dash_offset=0.0, line_cap_mode=CapMode.butt,
                 line_join_mode=JoinMode.miter, line_width=2.0, miter_limit=10.0, operator=Operator.src_over,
                 precision=Precision.integer, tolerance=default_tolerance, transformation=None, hairline=False):
        self.anti_aliasing_mode, self.dashes, self.dash_offset = anti_aliasing_mode, tuple(dashes), dash_offset
        self.line_cap_mode, self.line_join_mode, self.line_width = line_cap_mode, line_join_mode, line_width
        self.miter_limit, self.operator, self.precision, self.tolerance = miter_limit, operator, precision, tolerance
        self.transformation = transformation or Transformation()
        self.hairline = hairline
        self._keep = None

    def pod(self):
        o = abi.StrokeOptsPOD()
        o.anti_aliasing_mode, o.line_cap_mode, o.line_join_mode = int(self.anti_aliasing_mode), int(self.line_cap_mode), int(self.line_join_mode)
        o.op, o.precision, o.hairline = int(self.operator), int(self.precision), 1 if self.hairline else 0
        o.line_width, o.miter_limit, o.tolerance, o.dash_offset = float(self.line_width), float(self.miter_limit), float(self.tolerance), float(self.dash_offset)
        d = (C.c_double * max(1, len(self.dashes)))(*[float(v) for v in self.dashes])
        self._keep = d
        o.dashes = C.cast(d, C.POINTER(C.c_double))
        o.n_dashes = len(self.dashes)
        for i, v in enumerate(self.transformation.as_tuple()):
            o.ctm[i] = v
        return o


class painter:
    """painter.fill / painter.stroke (painter.zig:66, 214): the drop-in boundary."""

    @staticmethod
    def fill(surface, pattern, nodes, opts=None):
        opts = opts or FillOptions()
        arr = nodes_to_array(nodes)
        pat = pattern.pod()
        abi.check(surface.backend.fill(surface.handle, pat, arr, len(nodes), opts.pod()))

    @staticmethod
    def fill_glyphs(surface, pattern, instances, opts=None):
        """One painter.fill of a text run whose glyph outlines live in the backend's glyph cache: instances = [(glyph id,
        Transformation), ...] in run order (text.show, text.zig:73-195; z2d_fill_glyphs)."""
        opts = opts or FillOptions()
        arr = (abi.GlyphInstancePOD * max(1, len(instances)))()
        for i, (gid, tr) in enumerate(instances):
            arr[i].glyph = int(gid)
            for k, v in enumerate(tr.as_tuple()):
                arr[i].m[k] = v
        pat = pattern.pod()
        abi.check(surface.backend.fill_glyphs(surface.handle, pat, arr, len(instances), opts.pod()))

    @staticmethod
    def stroke(surface, pattern, nodes, opts=None):
        opts = opts or StrokeOptions()
        arr = nodes_to_array(nodes)
        pat = pattern.pod()
        abi.check(surface.backend.stroke(surface.handle, pat, arr, len(nodes), opts.pod()))


# --------------------------------------------------------------------------
# compositor.zig (surface-level API)
class Param:  # SurfaceCompositor.Operation.Param (compositor.zig:232-283)
    def __init__(self, kind, value=None):
        self.kind, self.value = ParamKind(kind), value

    @staticmethod
    def none():
        return Param(ParamKind.none)

    @staticmethod
    def pixel(px):
        return Param(ParamKind.pixel, px)

    @staticmethod
    def gradient(g):
        return Param(ParamKind.gradient, g)

    @staticmethod
    def dither(d):
        return Param(ParamKind.dither, d)

    @staticmethod
    def surface(s):
        return Param(ParamKind.surface, s)


class Operation:  # compositor.zig:220-230
    def __init__(self, operator, dst=None, src=None):
        self.operator, self.dst, self.src = Operator(operator), dst or Param.none(), src or Param.none()


class SurfaceCompositor:
    @staticmethod
    def run(dst, dst_x, dst_y, operations, precision=Precision.integer):  # compositor.zig:302-309
        n = len(operations)
        ops = (abi.CompOpPOD * max(1, n))()
        keep = []

        def fill_param(pod, prm):
            pod.kind = int(prm.kind)
            if prm.kind == ParamKind.surface:
                pod.surface = dst.backend.surface_param(prm.value.handle, keep)
            elif prm.kind == ParamKind.pixel:
                pat = Pattern.opaque(prm.value)
                keep.append(pat)
                pod.pattern = pat.pod()
            elif prm.kind == ParamKind.gradient:
                pat = Pattern.gradient(prm.value)
                keep.append(pat)
                pod.pattern = pat.pod()
            elif prm.kind == ParamKind.dither:
                pat = Pattern.dither(prm.value)
                keep.append(pat)
                pod.pattern = pat.pod()

        for i, o in enumerate(operations):
            ops[i].op = int(o.operator)
            fill_param(ops[i].dst, o.dst)
            fill_param(ops[i].src, o.src)
        abi.check(dst.backend.composite(dst.handle, int(dst_x), int(dst_y), ops, n, int(precision)))


# --------------------------------------------------------------------------
# Context.zig
class Context:
    def __init__(self, surface):
        self.surface = surface
        self.path = Path()
        self.pattern = Pattern.opaque(Pixel.rgba(0, 0, 0, 255))
        self.anti_aliasing_mode = AntiAliasMode.default
        self.dashes = ()
        self.dash_offset = 0.0
        self.dither = DitherType.none
        self.fill_rule = FillRule.non_zero
        self.hairline = False
        self.line_cap_mode = CapMode.butt
        self.line_join_mode = JoinMode.miter
        self.line_width = 2.0
        self.miter_limit = 10.0
        self.operator = Operator.src_over
        self.precision = Precision.integer
        self.tolerance = default_tolerance
        self.transformation = Transformation()

    def deinit(self):
        pass

    def set_source(self, source):  # Context.zig:115-121
        if source.kind == PatternKind.gradient:
            try:
                source.value.set_transformation(self.transformation)
            except abi.InvalidMatrix:
                return
        self.pattern = source

    def set_source_to_pixel(self, px):
        self.pattern = Pattern.opaque(px)

    def get_source(self):
        return self.pattern

    def get_transformation(self):
        return self.transformation

    def get_line_width(self):
        return self.line_width

    def device_to_user_distance(self, x, y):
        return self.transformation.device_to_user_distance(x, y)

    def user_to_device_distance(self, x, y):
        return self.transformation.user_to_device_distance(x, y)

    def set_dither(self, d):
        self.dither = DitherType(d)

    def set_anti_aliasing_mode(self, m):
        self.anti_aliasing_mode = AntiAliasMode(m)

    def set_fill_rule(self, r):
        self.fill_rule = FillRule(r)

    def set_line_cap_mode(self, m):
        self.line_cap_mode = CapMode(m)

    def set_line_join_mode(self, m):
        self.line_join_mode = JoinMode(m)

    def set_line_width(self, w):
        self.line_width = float(w)

    def set_dashes(self, d):
        self.dashes = tuple(float(v) for v in d)

    def set_dash_offset(self, o):
        self.dash_offset = float(o)

    def set_miter_limit(self, m):
        self.miter_limit = float(m)

    def set_operator(self, op):
        self.operator = Operator(op)

    def set_precision(self, p):
        self.precision = Precision(p)

    def set_tolerance(self, t):  # Context.zig:303-307
        t = max(float(t), 0.001)
        self.tolerance = t
        self.path.tolerance = t

    def set_hairline(self, h):
        self.hairline = bool(h)

    def set_transformation(self, t):  # Context.zig:345-348
        self.transformation = t
        self.path.transformation = t

    def set_identity(self):
        self.set_transformation(Transformation())

    def mul(self, a):
        self.set_transformation(self.transformation.mul(a))

    def translate(self, tx, ty):
        self.set_transformation(self.transformation.translate(tx, ty))

    def rotate(self, angle):
        self.set_transformation(self.transformation.rotate(angle))

    def scale(self, sx, sy):
        self.set_transformation(self.transformation.scale(sx, sy))

    def reset_path(self):
        self.path.reset()

    def move_to(self, x, y):
        self.path.move_to(x, y)

    def rel_move_to(self, x, y):
        self.path.rel_move_to(x, y)

    def line_to(self, x, y):
        self.path.line_to(x, y)

    def rel_line_to(self, x, y):
        self.path.rel_line_to(x, y)

    def curve_to(self, *a):
        self.path.curve_to(*a)

    def rel_curve_to(self, *a):
        self.path.rel_curve_to(*a)

    def arc(self, *a):
        self.path.arc(*a)

    def arc_negative(self, *a):
        self.path.arc_negative(*a)

    def close_path(self):
        self.path.close()

    def _wrap_dither(self):  # Context.zig:679-699
        if self.dither == DitherType.none:
            return self.pattern
        if self.pattern.kind == PatternKind.opaque:
            src = self.pattern.value
        elif self.pattern.kind == PatternKind.gradient:
            src = self.pattern.value
        else:
            return self.pattern
        scale = {Format.alpha1: 1, Format.alpha2: 2, Format.alpha4: 4}.get(self.surface.format, 8)
        return Pattern.dither(Dither(self.dither, src, scale))

    def fill(self):  # Context.zig:592-607
        painter.fill(self.surface, self._wrap_dither(), self.path.nodes,
                     FillOptions(self.anti_aliasing_mode, self.fill_rule, self.operator, self.precision, self.tolerance))

    def stroke(self):  # Context.zig:619-641
        painter.stroke(self.surface, self._wrap_dither(), self.path.nodes,
                       StrokeOptions(self.anti_aliasing_mode, self.dashes, self.dash_offset, self.line_cap_mode,
                                     self.line_join_mode, self.line_width, self.miter_limit, self.operator, self.precision,
                                     self.tolerance, self.transformation, self.hairline))
