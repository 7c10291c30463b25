"""ctypes mirror of include/z2d_cuda.h (the C-ABI PODs and enums).

Enum values follow the declaration order of the reference's Zig enums
(pixel.zig:47-56, compositor.zig:46-155, options.zig:23-91,
internal/path_nodes.zig:9-21), so this file is what a Zig/cgo/JNI binding of
the library would look like on the Python side.
"""
import ctypes as C
from enum import IntEnum

# status codes --------------------------------------------------------------
OK = 0
E_PATH_NOT_CLOSED = -1
E_PIXEL_SOURCE_NOT_PREMULTIPLIED = -2
E_INVALID_WIDTH = -3
E_INVALID_HEIGHT = -4
E_INVALID_STATE = -5
E_OUT_OF_MEMORY = -6
E_INVALID_MATRIX = -7
E_DEVICE = -8
E_INVALID_ARG = -9


class Z2DError(Exception):
    """Base of the error set the reference returns as Zig error unions."""

    code = None


class PathNotClosed(Z2DError):  # painter.zig:57
    code = E_PATH_NOT_CLOSED


class PixelSourceNotPreMultiplied(Z2DError):  # painter.zig:62
    code = E_PIXEL_SOURCE_NOT_PREMULTIPLIED


class InvalidWidth(Z2DError):  # surface.zig:87
    code = E_INVALID_WIDTH


class InvalidHeight(Z2DError):  # surface.zig:90
    code = E_INVALID_HEIGHT


class InvalidState(Z2DError):  # internal/InternalError.zig
    code = E_INVALID_STATE


class OutOfMemory(Z2DError):
    code = E_OUT_OF_MEMORY


class InvalidMatrix(Z2DError):  # Transformation.zig:27
    code = E_INVALID_MATRIX


class DeviceError(Z2DError):
    code = E_DEVICE


class InvalidArg(Z2DError):
    code = E_INVALID_ARG


_ERRORS = {c.code: c for c in (PathNotClosed, PixelSourceNotPreMultiplied, InvalidWidth, InvalidHeight, InvalidState,
                               OutOfMemory, InvalidMatrix, DeviceError, InvalidArg)}


def check(rc, detail=""):
    if rc == OK:
        return
    raise _ERRORS.get(rc, Z2DError)(f"status {rc} {detail}".strip())


class Format(IntEnum):
    argb = 0
    xrgb = 1
    rgb = 2
    rgba = 3
    alpha8 = 4
    alpha4 = 5
    alpha2 = 6
    alpha1 = 7


EXPORT_SRGB = 1         # z2d_surface_export flags (include/z2d_cuda.h)
EXPORT_FILTER_BYTE = 2

FORMAT_BITS = {Format.argb: 32, Format.xrgb: 32, Format.rgb: 32, Format.rgba: 32, Format.alpha8: 8, Format.alpha4: 4,
               Format.alpha2: 2, Format.alpha1: 1}


class Operator(IntEnum):
    clear = 0
    src = 1
    dst = 2
    src_over = 3
    dst_over = 4
    src_in = 5
    dst_in = 6
    src_out = 7
    dst_out = 8
    src_atop = 9
    dst_atop = 10
    xor = 11
    plus = 12
    multiply = 13
    screen = 14
    overlay = 15
    darken = 16
    lighten = 17
    color_dodge = 18
    color_burn = 19
    hard_light = 20
    soft_light = 21
    difference = 22
    exclusion = 23
    hue = 24
    saturation = 25
    color = 26
    luminosity = 27

    def requires_float(self):  # compositor.zig:165-177
        return self in (Operator.color_dodge, Operator.color_burn, Operator.soft_light, Operator.hue,
                        Operator.saturation, Operator.color, Operator.luminosity)

    def is_bounded(self):  # compositor.zig:187-196
        return self not in (Operator.src_in, Operator.dst_in, Operator.src_out, Operator.dst_atop)


class Precision(IntEnum):
    integer = 0
    float = 1


class FillRule(IntEnum):
    non_zero = 0
    even_odd = 1


class JoinMode(IntEnum):
    miter = 0
    round = 1
    bevel = 2


class CapMode(IntEnum):
    butt = 0
    round = 1
    square = 2


class AntiAliasMode(IntEnum):
    none = 0
    default = 1
    multisample_4x = 2
    supersample_4x = 3


class NodeTag(IntEnum):
    move_to = 0
    line_to = 1
    curve_to = 2
    close_path = 3


class ColorSpace(IntEnum):
    linear_rgb = 0
    srgb = 1
    hsl = 2


class GradientType(IntEnum):
    linear = 0
    radial = 1
    conic = 2


class Interp(IntEnum):
    linear_rgb = 0
    srgb = 1
    hsl = 2


class Polar(IntEnum):
    shorter = 0
    longer = 1
    increasing = 2
    decreasing = 3


class DitherType(IntEnum):
    none = 0
    bayer = 1
    blue_noise = 2


class DitherSource(IntEnum):
    pixel = 0
    color = 1
    gradient = 2


class PatternKind(IntEnum):
    opaque = 0
    gradient = 1
    dither = 2


class ParamKind(IntEnum):
    none = 0
    dither = 1
    gradient = 2
    pixel = 3
    surface = 4


# PODs ------------------------------------------------------------------------
class Node(C.Structure):
    _fields_ = [("tag", C.c_uint32), ("_pad", C.c_uint32), ("p", C.c_double * 6)]


class PixelPOD(C.Structure):
    _fields_ = [("format", C.c_uint32), ("r", C.c_uint8), ("g", C.c_uint8), ("b", C.c_uint8), ("a", C.c_uint8)]


class ColorPOD(C.Structure):
    _fields_ = [("space", C.c_uint32), ("c", C.c_float * 4)]


class StopPOD(C.Structure):
    _fields_ = [("offset", C.c_float), ("color", ColorPOD)]


class GradientPOD(C.Structure):
    _fields_ = [("type", C.c_uint32), ("method", C.c_uint32), ("polar", C.c_uint32), ("n_stops", C.c_uint32),
                ("geom", C.c_double * 6), ("inv_ctm", C.c_double * 6), ("stops", C.POINTER(StopPOD))]


class PatternPOD(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("pixel", PixelPOD), ("gradient", C.POINTER(GradientPOD)),
                ("dither_type", C.c_uint32), ("dither_source", C.c_uint32), ("dither_scale", C.c_uint32),
                ("dither_color", ColorPOD)]


class FillOptsPOD(C.Structure):
    _fields_ = [("anti_aliasing_mode", C.c_uint32), ("fill_rule", C.c_uint32), ("op", C.c_uint32),
                ("precision", C.c_uint32), ("tolerance", C.c_double)]


class StrokeOptsPOD(C.Structure):
    _fields_ = [("anti_aliasing_mode", C.c_uint32), ("line_cap_mode", C.c_uint32), ("line_join_mode", C.c_uint32),
                ("op", C.c_uint32), ("precision", C.c_uint32), ("hairline", C.c_uint32), ("line_width", C.c_double),
                ("miter_limit", C.c_double), ("tolerance", C.c_double), ("dash_offset", C.c_double),
                ("dashes", C.POINTER(C.c_double)), ("n_dashes", C.c_size_t), ("ctm", C.c_double * 6)]


class CompParamPOD(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("pattern", PatternPOD), ("surface", C.c_void_p)]


class CompOpPOD(C.Structure):
    _fields_ = [("op", C.c_uint32), ("dst", CompParamPOD), ("src", CompParamPOD)]


class GlyphInstancePOD(C.Structure):
    _fields_ = [("glyph", C.c_uint32), ("_pad", C.c_uint32), ("m", C.c_double * 6)]


class StatsPOD(C.Structure):
    _fields_ = [("draws", C.c_uint64), ("nodes", C.c_uint64), ("edges", C.c_uint64), ("band_edges", C.c_uint64),
                ("tile_items", C.c_uint64), ("tiles", C.c_uint64), ("covered_px", C.c_uint64), ("region_px", C.c_uint64),
                ("kernel_launches", C.c_uint64), ("h2d_bytes", C.c_uint64), ("tile_pairs", C.c_uint64), ("crossings", C.c_uint64),
                ("ms_flatten", C.c_float),
                ("ms_bin", C.c_float), ("ms_lists", C.c_float), ("ms_raster", C.c_float), ("ms_total", C.c_float),
                ("_pad", C.c_float)]


class DrawCmdPOD(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("_pad", C.c_uint32), ("surface", C.c_void_p), ("pattern", C.POINTER(PatternPOD)),
                ("nodes", C.POINTER(Node)), ("n_nodes", C.c_size_t), ("fill", C.POINTER(FillOptsPOD)),
                ("stroke", C.POINTER(StrokeOptsPOD))]


assert C.sizeof(Node) == 56
assert C.sizeof(GlyphInstancePOD) == 56


def surface_byte_len(fmt, w, h):
    """surface.zig:391-394,632: tightly packed, sub-byte formats bit-contiguous."""
    return (w * h * FORMAT_BITS[Format(fmt)] + 7) // 8
