"""The drop-in boundary without a GPU: the C header is plain C, the shared library loads, exports
every entry point include/z2d_cuda.h declares, the ctypes mirror has the C layout, and the product
refuses to run (loudly) when no CUDA device is present -- there is no CPU fallback."""
import ctypes as C
import os
import re
import subprocess

import pytest

from z2d_b200 import abi, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "z2d_cuda.h")


def _declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(z2d_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    return C.CDLL(build.build())


def test_header_declares_the_expected_entry_points():
    names = _declared_functions()
    for must in ("z2d_ctx_create", "z2d_fill", "z2d_stroke", "z2d_composite", "z2d_submit", "z2d_replay", "z2d_surface_download"):
        assert must in names
    assert len(names) >= 23


def test_library_exports_every_declared_symbol(lib):
    missing = [n for n in _declared_functions() if not hasattr(lib, n)]
    assert not missing, f"libz2d_cuda.so does not export: {missing}"


def test_library_does_not_link_the_oracle():
    """The oracle is test infrastructure; the product must not depend on it."""
    out = subprocess.run(["ldd", build.SO], capture_output=True, text=True).stdout
    assert "oracle" not in out
    syms = subprocess.run(["nm", "-D", "--defined-only", build.SO], capture_output=True, text=True).stdout
    assert "z2d_ref_" not in syms


def test_version(lib):
    lib.z2d_version.restype = C.c_int32
    assert lib.z2d_version() == 1


_PODS = {"z2d_node": abi.Node, "z2d_pixel": abi.PixelPOD, "z2d_color": abi.ColorPOD, "z2d_stop": abi.StopPOD,
         "z2d_gradient": abi.GradientPOD, "z2d_pattern": abi.PatternPOD, "z2d_fill_opts": abi.FillOptsPOD,
         "z2d_stroke_opts": abi.StrokeOptsPOD, "z2d_comp_param": abi.CompParamPOD, "z2d_comp_op": abi.CompOpPOD,
         "z2d_stats": abi.StatsPOD, "z2d_draw_cmd": abi.DrawCmdPOD}


def test_header_is_plain_c_and_layouts_match_ctypes(tmp_path):
    src = tmp_path / "layout.c"
    body = "\n".join(f'  printf("{n} %zu\\n", sizeof({n}));' for n in _PODS)
    src.write_text(f'#include <stdio.h>\n#include "z2d_cuda.h"\nint main(void) {{\n{body}\n  return 0;\n}}\n')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.dirname(HEADER), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout
    sizes = dict(line.split() for line in out.strip().splitlines())
    for name, pod in _PODS.items():
        assert int(sizes[name]) == C.sizeof(pod), f"{name}: C {sizes[name]} bytes, ctypes {C.sizeof(pod)}"


def test_status_codes_match_header():
    text = open(HEADER).read()
    codes = {k: int(v) for k, v in re.findall(r"\b(Z2D_(?:OK|E_[A-Z_]+))\s*=\s*(-?\d+)", text)}
    assert codes["Z2D_OK"] == abi.OK == 0
    mirror = {"Z2D_" + k: getattr(abi, k) for k in dir(abi) if k.startswith("E_")}
    assert mirror == {k: v for k, v in codes.items() if k != "Z2D_OK"}
    assert set(abi._ERRORS) == set(mirror.values())


def test_no_cpu_fallback_without_a_device(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    ctx = C.c_void_p()
    lib.z2d_ctx_create.restype = C.c_int32
    rc = lib.z2d_ctx_create(0, None, C.byref(ctx))
    assert rc != 0 and not ctx.value
    from z2d_b200.cuda_backend import CudaBackend
    with pytest.raises(Exception):
        CudaBackend()
