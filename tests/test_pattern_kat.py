"""Pattern known-answer tests lifted from the reference: gradient getPixel (src/gradient.zig:1252-1372, 1608-1619, 1768-1779)
and Dither.getPixel (src/Dither.zig:164-262); fixture tests/golden/pattern_kat.json (hand-transcribed).

CPU: the oracle's per-pixel pattern evaluation must return the reference's expected premultiplied RGBA8 exactly.
GPU: the same pixels through z2d_composite (`src` operator onto an RGBA surface large enough to hold the coordinate;
negative coordinates exist only on the CPU side), +-1 LSB as the north star allows for floating-point sources."""
import ctypes as C
import json
import os

import pytest

from tests import specs
from tests.oracle_backend import load_oracle
from z2d_b200 import abi, host
from z2d_b200.abi import DitherType, Format, Interp, Operator

KAT = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pattern_kat.json")))


def _gradient(name):
    g = KAT["gradients"][name]
    ctor = {"linear": host.Gradient.linear, "radial": host.Gradient.radial, "conic": host.Gradient.conic}[g["type"]]
    out = ctor(*g["geom"], method=Interp[g["method"]])
    for off, col in g["stops"]:
        (k, v), = col.items()
        out.add_stop(off, {k: tuple(v)})
    return out


def _dither_pattern(case):
    (kind, v), = case["source"].items()
    if kind == "pixel":
        src = host.Pixel.rgba(*v)
    elif kind == "color":
        (k, c), = v.items()
        src = {k: tuple(c)}
    else:
        src = _gradient(v)
    return host.Pattern.dither(host.Dither(DitherType[case["type"]], src, case["scale"]))


def _oracle_pixel(pattern, x, y):
    lib = load_oracle()
    pod = pattern.pod()
    out = (C.c_uint8 * 4)()
    lib.z2d_ref_pattern_pixel(C.byref(pod), x, y, out)
    return list(out)


GRAD_IDS = [f"{n}@{x},{y}" for n, x, y, _ in KAT["gradient_cases"]]


@pytest.mark.parametrize("case", KAT["gradient_cases"], ids=GRAD_IDS)
def test_oracle_gradient_pixel(case):
    name, x, y, expected = case
    assert _oracle_pixel(host.Pattern.gradient(_gradient(name)), x, y) == expected


@pytest.mark.parametrize("case", KAT["dither_cases"], ids=[c["name"].replace(" ", "_") for c in KAT["dither_cases"]])
def test_oracle_dither_pixel(case):
    assert _oracle_pixel(_dither_pattern(case), case["x"], case["y"]) == case["expected"]


def _device_pixel(cuda, param, x, y):
    z = specs.bind(cuda)
    sfc = z.Surface(Format.rgba, x + 1, y + 1)
    z.SurfaceCompositor.run(sfc, 0, 0, [z.Operation(Operator.src, src=param)])
    return sfc.pixels()[y, x].astype(int).tolist()


@pytest.mark.gpu
@pytest.mark.parametrize("case", [c for c in KAT["gradient_cases"] if c[1] >= 0 and c[2] >= 0],
                         ids=[i for i, c in zip(GRAD_IDS, KAT["gradient_cases"]) if c[1] >= 0 and c[2] >= 0])
def test_device_gradient_pixel(cuda, case):
    name, x, y, expected = case
    got = _device_pixel(cuda, host.Param.gradient(_gradient(name)), x, y)
    assert max(abs(a - b) for a, b in zip(got, expected)) <= 1, f"{got} vs {expected}"


@pytest.mark.gpu
@pytest.mark.parametrize("case", KAT["dither_cases"], ids=[c["name"].replace(" ", "_") for c in KAT["dither_cases"]])
def test_device_dither_pixel(cuda, case):
    pat = _dither_pattern(case)
    got = _device_pixel(cuda, host.Param.dither(pat.value), case["x"], case["y"])
    assert max(abs(a - b) for a, b in zip(got, case["expected"])) <= 1, f"{got} vs {case['expected']}"
