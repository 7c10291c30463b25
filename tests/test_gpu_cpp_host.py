"""The compiled-language host side: include/z2d.hpp (C++17 mirror of z2d's API over the C ABI).

CPU (-m "not gpu"): the header and tests/cpp/scenes.cpp compile and link against libz2d_cuda.so.
GPU: the program renders five scenes; each has a line-for-line Python twin below, rendered by the CPU oracle, and the raw
surface bytes must match (exactly for the integer pipeline, +-1 LSB where the reference computes in floating point)."""
import math
import os
import subprocess

import numpy as np
import pytest

from tests import specs
from z2d_b200 import build
from z2d_b200.abi import AntiAliasMode, CapMode, FillRule, Format, Interp, JoinMode, Operator, Precision

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def scenes_exe(tmp_path_factory):
    so = build.build()
    exe = tmp_path_factory.mktemp("cpp") / "scenes"
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "scenes.cpp"), "-o", str(exe), so, f"-Wl,-rpath,{os.path.dirname(so)}"], check=True)
    return str(exe)


def test_cpp_host_compiles_and_links(scenes_exe):
    assert os.path.exists(scenes_exe)


# ---- Python twins of tests/cpp/scenes.cpp ------------------------------------------------------------------------------
def bezier_fill_rgba(z):
    sfc = z.Surface(Format.rgba, 300, 300)
    c = z.Context(sfc)
    c.set_source_to_pixel(z.Pixel.rgba(90, 40, 10, 128))
    c.set_fill_rule(FillRule.even_odd)
    c.move_to(19, 249)
    c.curve_to(89, 49, 209, 49, 279, 249)
    c.curve_to(209, 149, 89, 149, 19, 20.5)
    c.close_path()
    c.move_to(100, 100)
    c.line_to(250.25, 120)
    c.line_to(140, 280.75)
    c.close_path()
    c.fill()
    c.reset_path()
    c.set_source_to_pixel(z.Pixel.rgba(0, 100, 200, 200))
    c.set_operator(Operator.multiply)
    c.set_precision(Precision.float)
    c.set_fill_rule(FillRule.non_zero)
    c.set_anti_aliasing_mode(AntiAliasMode.supersample_4x)
    c.move_to(10, 150)
    c.line_to(290, 130)
    c.line_to(290, 190)
    c.line_to(10, 170)
    c.close_path()
    c.fill()
    return sfc


def dashed_stroke_arc(z):
    sfc = z.Surface(Format.rgb, 400, 400)
    c = z.Context(sfc)
    c.set_source_to_pixel(z.Pixel.rgb(0xFF, 0xFF, 0xFF))
    c.set_line_width(6)
    c.set_line_join_mode(JoinMode.round)
    c.set_line_cap_mode(CapMode.round)
    c.set_dashes((25, 10, 5, 10))
    c.set_dash_offset(7.5)
    c.translate(200, 200)
    c.scale(150, 100)
    c.arc(0, 0, 1, 0, 2 * math.pi)
    c.close_path()
    c.stroke()
    c.reset_path()
    c.set_identity()
    c.set_dashes(())
    c.set_line_join_mode(JoinMode.miter)
    c.set_line_cap_mode(CapMode.square)
    c.set_source_to_pixel(z.Pixel.rgb(0x20, 0xC0, 0x40))
    c.rotate(0.25)
    c.move_to(120, 20)
    c.line_to(300, 60)
    c.rel_line_to(-60, 120)
    c.rel_curve_to(-30, 40, -90, 40, -120, 0)
    c.stroke()
    return sfc


def conic_gradient_alpha4(z):
    sfc = z.Surface(Format.alpha4, 301, 299)
    g = z.Gradient.conic(149, 149, 0.5)
    g.add_stop(0, {"rgba": (1, 0, 0, 1)})
    g.add_stop(0.5, {"rgba": (0, 1, 0, 0.25)})
    g.add_stop(1, {"rgba": (0, 0, 1, 1)})
    c = z.Context(sfc)
    c.set_source(z.Pattern.gradient(g))
    c.arc(149, 149, 120, 0, math.pi * 2)
    c.close_path()
    c.fill()
    return sfc


def compositor_ops(z):
    dst = z.SurfacePixel(z.Pixel.rgba(40, 80, 120, 160), 128, 96)
    g = z.Gradient.linear(0, 0, 127, 95, method=Interp.srgb)
    g.add_stop(0, {"rgb": (1, 0, 0)})
    g.add_stop(0.5, {"rgba": (0, 1, 0, 0.5)})
    g.add_stop(1, {"rgb": (0, 0, 1)})
    z.SurfaceCompositor.run(dst, 0, 0, [z.Operation(Operator.xor, src=z.Param.gradient(g))], precision=Precision.float)
    stamp = z.SurfacePixel(z.Pixel.rgba(100, 0, 50, 100), 64, 64)
    stamp.put_pixel(3, 3, z.Pixel.rgba(255, 255, 255, 255))
    z.SurfaceCompositor.run(dst, -10, 5, [z.Operation(Operator.src_over, src=z.Param.surface(stamp))])
    z.SurfaceCompositor.run(dst, 100, 70, [z.Operation(Operator.plus, src=z.Param.surface(stamp))])
    return dst


def hairline_unbounded(z):
    sfc = z.SurfacePixel(z.Pixel.rgba(10, 20, 30, 255), 200, 150)
    c = z.Context(sfc)
    c.set_source_to_pixel(z.Pixel.rgba(200, 100, 50, 200))
    c.set_anti_aliasing_mode(AntiAliasMode.none)
    c.set_operator(Operator.dst_in)
    c.move_to(30, 20)
    c.line_to(170, 40)
    c.line_to(100, 130)
    c.close_path()
    c.fill()
    c.reset_path()
    c.set_operator(Operator.src_over)
    c.set_anti_aliasing_mode(AntiAliasMode.default)
    c.set_hairline(True)
    c.set_source_to_pixel(z.Pixel.rgba(255, 255, 255, 255))
    c.move_to(5, 5)
    c.line_to(190, 140)
    c.line_to(190, 10)
    c.curve_to(150, 60, 60, 60, 10, 140)
    c.stroke()
    return sfc


TWINS = {"bezier_fill_rgba": (bezier_fill_rgba, 1), "dashed_stroke_arc": (dashed_stroke_arc, 0),
         "conic_gradient_alpha4": (conic_gradient_alpha4, None), "compositor_ops": (compositor_ops, 1),
         "hairline_unbounded": (hairline_unbounded, 0)}


@pytest.fixture(scope="module")
def cpp_outputs(scenes_exe, tmp_path_factory):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    out = tmp_path_factory.mktemp("cpp_out")
    res = subprocess.run([scenes_exe, str(out)], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "errors_ok=5" in res.stdout
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(TWINS))
def test_cpp_scene_matches_oracle(cpp_outputs, oracle, name):
    twin, tol = TWINS[name]
    ref_sfc = twin(specs.bind(oracle))
    got = np.fromfile(os.path.join(cpp_outputs, name + ".bin"), dtype=np.uint8)
    ref = np.asarray(ref_sfc.download(), dtype=np.uint8)
    assert got.shape == ref.shape
    if tol == 0:
        assert np.array_equal(got, ref), f"{int((got != ref).sum())} bytes differ"
    elif tol is None:  # packed 4-bit samples through a float gradient: compare decoded samples with +-1 level
        g = np.stack([got & 15, got >> 4], -1).astype(np.int32)
        r = np.stack([ref & 15, ref >> 4], -1).astype(np.int32)
        assert np.abs(g - r).max() <= 1
    else:
        assert np.abs(got.astype(np.int32) - ref.astype(np.int32)).max() <= tol
