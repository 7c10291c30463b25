"""Operator known-answer tests lifted from the reference (src/compositor.zig:3078-3860, 102 cases; fixture
tests/golden/compositor_kat.json written by tests/golden/extract_compositor_kat.py).

CPU: the oracle's runPixel must return the reference's expected pixel exactly (this pins the oracle's
IntegerOps / FloatOps independently of the golden images).  GPU: the same cases through z2d_composite on a
1x1 surface (SurfaceCompositor upgrades float-only operators to float precision, compositor.zig:317-322, so
integer cases of float-only operators are exercised with the float expectation)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from tests.oracle_backend import load_oracle
from z2d_b200 import host
from z2d_b200.abi import Format, Operator, Precision

KAT = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "compositor_kat.json")))["cases"]
IDS = [f"{c['precision']}-{c['name'].replace(' ', '_')}" for c in KAT]
FLOAT_ONLY = {"color_dodge", "color_burn", "soft_light", "hue", "saturation", "color", "luminosity"}


def _pixel(args):
    (kind, vals), = args.items()
    return host.Pixel.from_color({kind: tuple(vals)})


def test_fixture_is_complete():
    assert len(KAT) == 102
    assert {c["operator"] for c in KAT} == {o.name for o in Operator}


@pytest.mark.parametrize("case", KAT, ids=IDS)
def test_oracle_run_pixel(case):
    lib = load_oracle()
    bg, fg = _pixel(case["bg"]), _pixel(case["fg"])
    d = (C.c_uint8 * 4)(bg.r, bg.g, bg.b, bg.a)
    s = (C.c_uint8 * 4)(fg.r, fg.g, fg.b, fg.a)
    out = (C.c_uint8 * 4)()
    lib.z2d_ref_run_pixel(int(Precision[case["precision"]]), d, s, int(Operator[case["operator"]]), out)
    assert list(out) == case["expected"], case["name"]


@pytest.mark.gpu
@pytest.mark.parametrize("case", KAT, ids=IDS)
def test_device_composite(cuda, case):
    if case["precision"] == "integer" and case["operator"] in FLOAT_ONLY:
        pytest.skip("SurfaceCompositor runs float-only operators in float precision")
    from tests import specs
    z = specs.bind(cuda)
    sfc = z.Surface(Format.rgba, 1, 1)
    bg, fg = _pixel(case["bg"]), _pixel(case["fg"])
    z.SurfaceCompositor.run(sfc, 0, 0, [z.Operation(Operator[case["operator"]], dst=z.Param.pixel(bg), src=z.Param.pixel(fg))],
                            precision=Precision[case["precision"]])
    got = sfc.pixels()[0, 0].astype(int).tolist()
    exp = case["expected"]
    if case["precision"] == "float":
        assert max(abs(a - b) for a, b in zip(got, exp)) <= 1, f"{case['name']}: {got} vs {exp}"  # +-1 LSB (north star)
    else:
        assert got == exp, f"{case['name']}: {got} vs {exp}"
