"""Python ports of the reference's acceptance scenes (spec/NNN_*.zig).

Each scene is a function `render(z, aa_mode) -> Surface` (path scenes) or
`render(z) -> Surface` (compositor scenes), registered with the golden file
stem it reproduces.  `z` is a small namespace bound to one backend (the CUDA
library or the CPU oracle), so the same code renders through either.  Integer
arithmetic in the Zig scenes (comptime ints, `/` == truncating division) is
kept as `//` here.
"""
import types

from z2d_b200 import host
from z2d_b200.abi import (AntiAliasMode, CapMode, DitherType, FillRule, Format, Interp, JoinMode, Operator, Polar,
                          Precision)

PATH_SCENES = {}        # stem -> render(z, aa)
COMPOSITOR_SCENES = {}  # stem -> render(z)
COLOR_PROFILE = {}      # stem -> "srgb" when the export re-encodes with gamma


def path_scene(stem):
    def deco(fn):
        PATH_SCENES[stem] = fn
        return fn
    return deco


def compositor_scene(stem, profile=None):
    def deco(fn):
        COMPOSITOR_SCENES[stem] = fn
        if profile:
            COLOR_PROFILE[stem] = profile
        return fn
    return deco


def bind(backend):
    """Namespace with the host API bound to `backend`."""
    z = types.SimpleNamespace()
    z.backend = backend
    z.Surface = lambda fmt, w, h: host.Surface(fmt, w, h, None, backend)
    z.SurfacePixel = lambda px, w, h: host.Surface(px.format, w, h, px, backend)
    z.Context = host.Context
    z.Pixel = host.Pixel
    z.Gradient = host.Gradient
    z.Pattern = host.Pattern
    z.Dither = host.Dither
    z.Transformation = host.Transformation
    z.Path = host.Path
    z.painter = host.painter
    z.FillOptions = host.FillOptions
    z.StrokeOptions = host.StrokeOptions
    z.SurfaceCompositor = host.SurfaceCompositor
    z.Operation = host.Operation
    z.Param = host.Param
    return z


from . import fills, strokes, compositing, extra, text  # noqa: E402,F401  (register scenes)
