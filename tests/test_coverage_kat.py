"""Coverage known-answer tests lifted from the reference's unit tests:

* `addSpan` (src/internal/raster/multisample.zig:281-428): sub-scanline spans [x0 + j, x1 - j), j = 0..k-1, accumulate to per-pixel
  coverage 4 | 7,8,7 | 9,12,9 | 10,16,10 for k = 1..4 -- at the offset / length of the first test (pixels 50..149) and at the
  capacity boundaries of the second (255, 256, 65535, 65536, 131072 pixels wide).  There is no sparse coverage buffer here; the
  same spans are produced by filling the trapezoid whose edges cross sub-scanline j at exactly x0 + j and x1 - j, and the expected
  numbers are checked through the coverage -> alpha map of multisample.zig:203-224 (16 -> opaque, else 16 * cov - 1).
* `Polygon.inBox` (src/internal/tess/Polygon.zig:452-784, tests/golden/inbox_kat.json): on the oracle directly; on the device
  through a rectangle with those extents, which is rasterised at all (non-empty region) exactly when the reference says true.
"""
import ctypes as C
import json
import os

import numpy as np
import pytest

from tests import specs
from z2d_b200 import host
from z2d_b200.abi import AntiAliasMode, Format

INBOX = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "inbox_kat.json")))["cases"]
BACKENDS = ["oracle", pytest.param("cuda", marks=pytest.mark.gpu)]
EXPECTED = {1: (4, 4, 4), 2: (7, 8, 7), 3: (9, 12, 9), 4: (10, 16, 10)}  # (first pixel, middle, last pixel) after k spans


def _alpha(cov):
    return 255 if cov == 16 else 16 * cov - 1


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("width,first,count", [(200, 50, 100), (255, 0, 255), (256, 0, 256), (65535, 0, 65535), (65536, 0, 65536), (131072, 0, 131072)])
@pytest.mark.parametrize("k", [1, 2, 3, 4])
def test_add_span_accumulation(request, backend, width, first, count, k):
    z = specs.bind(request.getfixturevalue(backend))
    sfc = z.Surface(Format.alpha8, width, 1)
    x0, x1 = 4 * first, 4 * (first + count)  # supersampled span of sub-scanline 0: [x0, x1)
    p = z.Path()
    p.move_to((x0 - 0.5) / 4, 0.0)
    p.line_to((x1 + 0.5) / 4, 0.0)
    p.line_to((x1 + 0.5 - k) / 4, k / 4)
    p.line_to((x0 - 0.5 + k) / 4, k / 4)
    p.close()
    z.painter.fill(sfc, host.Pattern.opaque(host.Pixel.alpha8(255)), p.nodes, z.FillOptions(anti_aliasing_mode=AntiAliasMode.multisample_4x))
    got = sfc.download().astype(int)
    e_first, e_mid, e_last = (_alpha(c) for c in EXPECTED[k])
    assert got[first] == e_first and got[first + count - 1] == e_last
    assert (got[first + 1:first + count - 1] == e_mid).all()
    assert not got[:first].any() and not got[first + count:].any()


@pytest.mark.parametrize("case", INBOX, ids=[c["name"].replace(" ", "_").replace(",", "") for c in INBOX])
def test_in_box_oracle(case):
    from tests.oracle_backend import load_oracle
    lib = load_oracle()
    lib.z2d_ref_in_box.restype = C.c_int32
    lib.z2d_ref_in_box.argtypes = [C.c_double] * 5 + [C.c_int32] * 2
    got = lib.z2d_ref_in_box(case["left"], case["top"], case["right"], case["bottom"], case["scale"], case["arg_width"], case["arg_height"])
    assert bool(got) == case["expected"]


@pytest.mark.gpu
@pytest.mark.parametrize("case", INBOX, ids=[c["name"].replace(" ", "_").replace(",", "") for c in INBOX])
def test_in_box_device(cuda, case):
    z = specs.bind(cuda)
    s = case["scale"]
    sfc = z.Surface(Format.rgba, case["arg_width"], case["arg_height"])
    p = z.Path()
    l, t, r, b = (case[k] / s for k in ("left", "top", "right", "bottom"))
    p.move_to(l, t); p.line_to(r, t); p.line_to(r, b); p.line_to(l, b); p.close()
    aa = AntiAliasMode.none if s == 1.0 else AntiAliasMode.multisample_4x
    z.painter.fill(sfc, host.Pattern.opaque(host.Pixel.rgba(255, 255, 255, 255)), p.nodes, z.FillOptions(anti_aliasing_mode=aa))
    cuda.sync()
    st = cuda.stats()
    assert (st["region_px"] > 0) == case["expected"], (case, st["region_px"], st["edges"])
