"""Compare rendered surfaces with the reference's golden PNGs, per decoded pixel.

Mirrors what src/export_png.zig does to a surface before writing it
(export_png.zig:200-360): RGB/XRGB drop the padding byte, RGBA/ARGB are
de-multiplied in integer space (pixel_vector.zig:27-49), alpha8 is written as
8-bit grey, alpha4/2/1 as 4/2/1-bit grey, and an optional sRGB profile
re-encodes with gamma 2.2.
"""
import os

import numpy as np
from PIL import Image

from z2d_b200.abi import AntiAliasMode, Format

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "spec_files")

AA_SUFFIX = [(AntiAliasMode.none, "_pixelated"), (AntiAliasMode.supersample_4x, "_smooth"),
             (AntiAliasMode.multisample_4x, "_smooth_multisample")]


def golden_path(stem, aa=None):
    """Golden file for a scene (+AA mode); MSAA falls back to _smooth (main_spec.zig:787-805)."""
    if aa is None:
        return os.path.join(GOLDEN_DIR, stem + ".png")
    suffix = dict(AA_SUFFIX)[aa]
    p = os.path.join(GOLDEN_DIR, stem + suffix + ".png")
    if aa == AntiAliasMode.multisample_4x and not os.path.exists(p):
        p = os.path.join(GOLDEN_DIR, stem + "_smooth.png")
    return p


def _srgb_encode(c8):
    x = c8.astype(np.float32) / np.float32(255.0)
    y = np.power(x, np.float32(1 / np.float32(2.2)), dtype=np.float32)
    v = np.float32(255.0) * y
    return np.floor(v + np.float32(0.5)).astype(np.int32)


def export_view(surface, profile=None):
    """Surface -> (array, kind) in the representation the PNG holds."""
    px = surface.pixels().astype(np.int32)  # h,w,4 RGBA (decoded)
    fmt = surface.format
    if fmt in (Format.rgb, Format.xrgb):
        rgb = px[..., :3]
        if profile == "srgb":
            rgb = _srgb_encode(rgb)
        return rgb.astype(np.uint8), "RGB"
    if fmt in (Format.rgba, Format.argb):
        a = px[..., 3:4]
        rgb = np.where(a == 0, 0, px[..., :3] * 255 // np.maximum(1, a))
        out = np.concatenate([rgb, a], axis=-1)
        if profile == "srgb":
            out = np.concatenate([_srgb_encode(out[..., :3]), a], axis=-1)
        return out.astype(np.uint8), "RGBA"
    return px[..., 3].astype(np.uint8), {Format.alpha8: "L8", Format.alpha4: "L4", Format.alpha2: "L2", Format.alpha1: "L1"}[fmt]


def load_golden(path, kind):
    im = Image.open(path)
    if kind == "RGB":
        return np.asarray(im.convert("RGB"))
    if kind == "RGBA":
        assert im.mode == "RGBA", im.mode
        return np.asarray(im)
    arr = np.asarray(im.convert("L")).astype(np.int32)
    # Pillow expands sub-8-bit greys to 0..255; bring them back to raw samples
    if kind == "L4":
        return (arr // 17).astype(np.uint8)
    if kind == "L2":
        return (arr // 85).astype(np.uint8)
    if kind == "L1":
        return (arr // 255).astype(np.uint8)
    return arr.astype(np.uint8)


def diff_count(surface, path, profile=None):
    got, kind = export_view(surface, profile)
    exp = load_golden(path, kind)
    if got.shape != exp.shape:
        return -1, got, exp
    d = got != exp
    if d.ndim == 3:
        d = d.any(axis=-1)
    return int(d.sum()), got, exp
