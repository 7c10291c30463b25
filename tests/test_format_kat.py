"""Per-format known-answer tests lifted from the reference (src/compositor.zig:2452-3076: `src_over` and `dst_in` on every
destination format with sources of every format; fixture tests/golden/format_kat.json written by
tests/golden/extract_format_kat.py).  CPU: the oracle; GPU: the same cases through z2d_composite on a 1x1 surface of the
destination's format.  All integer precision: exact."""
import json
import os

import pytest

from tests import specs
from z2d_b200 import host
from z2d_b200.abi import Format, Operator, Precision

KAT = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "format_kat.json")))["cases"]
IDS = [f"{i}-{c['operator']}-{c['dst']['format']}-from-{c['src']['format']}" for i, c in enumerate(KAT)]


def _px(p):
    return host.Pixel(Format[p["format"]], p["r"], p["g"], p["b"], p["a"])


def test_fixture_is_complete():
    assert len(KAT) == 76
    assert {c["dst"]["format"] for c in KAT} == {"rgb", "rgba", "alpha8", "alpha4", "alpha2", "alpha1"}


def _run(z, case):
    dst = _px(case["dst"])
    sfc = z.Surface(dst.format, 1, 1)
    sfc.paint_pixel(dst)
    z.SurfaceCompositor.run(sfc, 0, 0, [z.Operation(Operator[case["operator"]], src=z.Param.pixel(_px(case["src"])))],
                            precision=Precision[case["precision"]])
    r, g, b, a = (int(v) for v in sfc.pixels()[0, 0])
    exp = case["expected"]
    if exp["format"] in ("rgb", "xrgb"):
        assert (r, g, b) == (exp["r"], exp["g"], exp["b"]), (case, (r, g, b, a))
    elif exp["format"].startswith("alpha"):
        assert a == exp["a"], (case, a)
    else:
        assert (r, g, b, a) == (exp["r"], exp["g"], exp["b"], exp["a"]), (case, (r, g, b, a))


@pytest.mark.parametrize("case", KAT, ids=IDS)
def test_oracle(oracle, case):
    _run(specs.bind(oracle), case)


@pytest.mark.gpu
@pytest.mark.parametrize("case", KAT, ids=IDS)
def test_device(cuda, case):
    _run(specs.bind(cuda), case)
