"""The unit stroker (z2d_b200/csrc/stroke_units.cuh: walker -> units -> links) against the sub-path stroker (stroke.cuh), both
compiled FOR THE HOST from the files the library compiles for the device (tools/stroke_units_host_test.cpp): the same multiset of
edges for every sub-path of a random corpus, and the merged Pen vertex search against the reference's two-copy form
(Pen.zig:138-232), and the upper bound on a cubic's segment count that sizes the edge ranges of the single-pass fill
flattening (geom.cuh curve_edge_bound) against Spline.decompose on 400 000 random and degenerate curves.  The sub-path stroker is the one every oracle / golden comparison of round 1 pinned; on the GPU both run under
the same parity tests (Z2D_NO_STROKE_UNITS=1 selects the old one)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs a host C++ compiler")
def test_unit_stroker_equals_subpath_stroker_on_host(tmp_path):
    exe = str(tmp_path / "stroke_units_host_test")
    inc = [d for d in ("/usr/local/cuda/include", os.path.join(os.environ.get("CUDA_HOME", ""), "include")) if os.path.exists(os.path.join(d, "vector_types.h"))]
    if not inc:
        pytest.skip("CUDA headers (vector_types.h) not found")
    cmd = ["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-Wno-unknown-pragmas", "-I" + inc[0], "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tools", "stroke_units_host_test.cpp"), "-o", exe]
    res = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    run = subprocess.run([exe, "4000"], capture_output=True, text=True, timeout=600)
    assert run.returncode == 0, run.stdout[-2000:]
    assert " 0 mismatches" in run.stdout
    assert "curve bound violated" not in run.stdout
