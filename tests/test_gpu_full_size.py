"""Parity at BASELINE.json's full sizes, through the C ABI on the GPU.

The oracle cannot render these sizes in seconds, so each configuration is checked by (a) an exact comparison of a
bounded part of the same workload at the full canvas size and (b) size-independent properties of the full run.
  C2  4096^2, 100k cubic paths      : the WHOLE scene == oracle; full run deterministic; alpha never decreases (src_over)
  C3  2048^2, 50k strokes           : the WHOLE scene == oracle; full run chunked == single batch
  C4  8192^2 composites             : all 28 operators (integer + float) x {pixel, linear, radial, conic, dither} sources on
                                      RGBA, plus RGB / alpha8 / alpha4 / alpha2 / alpha1 destinations: three 2-row strips of the
                                      full surface == oracle (gradients evaluated through a translated transformation)
  C5  batch of 1024^2 mixed scenes  : sampled scenes == oracle; sharding scenes over 2 ranks gives the same checksums
"""
import ctypes as C

import numpy as np
import pytest

from tests import specs
from tests.oracle_backend import load_oracle, render_scene
from z2d_b200 import abi, host, sharding, workloads
from z2d_b200.abi import Format, Operator, Precision
from z2d_b200.host import Surface

pytestmark = pytest.mark.gpu


def _submit(cuda, scene, sfc, lo=0, hi=None):
    cmds = scene.draw_cmds(sfc.handle, lo, hi)
    cuda.submit(cmds.ctypes.data_as(C.POINTER(abi.DrawCmdPOD)), len(cmds))


# ------------------------------------------------------------------------------------------------ C2
@pytest.fixture(scope="module")
def c2_scene():
    return workloads.cubic_paths_scene(100_000, 4096)


def test_c2_whole_scene_matches_oracle(cuda, c2_scene):
    """All 100 000 ordered fills of BASELINE config 2, byte for byte (the CPU restatement needs ~20 s for the scene)."""
    sfc = Surface(Format.rgba, 4096, 4096, None, cuda)
    _submit(cuda, c2_scene, sfc)
    got = sfc.download()
    ref = render_scene(load_oracle(fast=True), c2_scene)
    sfc.deinit()
    assert np.array_equal(got, ref), f"{int((got != ref).sum())} bytes differ"


def test_c2_full_run_properties(cuda, c2_scene):
    a = Surface(Format.rgba, 4096, 4096, None, cuda)
    _submit(cuda, c2_scene, a)
    full = a.download().copy()
    a.paint_pixel(host.Pixel.rgba(0, 0, 0, 0))
    cuda.set_chunk(7777)  # different batch boundaries, same result
    _submit(cuda, c2_scene, a)
    again = a.download().copy()
    cuda.set_chunk(32768)
    assert np.array_equal(full, again), "the full run is not deterministic / depends on batch boundaries"
    a.paint_pixel(host.Pixel.rgba(0, 0, 0, 0))
    _submit(cuda, c2_scene, a, 0, 50_000)
    half = a.download()
    a.deinit()
    assert (full.reshape(-1, 4)[:, 3] >= half.reshape(-1, 4)[:, 3]).all(), "src_over lowered an alpha value"
    px = full.reshape(-1, 4).astype(np.int32)
    assert (px[:, :3] <= px[:, 3:4]).all(), "result is not premultiplied"


# ------------------------------------------------------------------------------------------------ C3
@pytest.fixture(scope="module")
def c3_scene():
    return workloads.stroke_paths_scene(50_000, 2048)


def test_c3_whole_scene_matches_oracle(cuda, c3_scene):
    """All 50 000 strokes of BASELINE config 3, byte for byte (about a minute of CPU for the restatement)."""
    sfc = Surface(Format.rgba, 2048, 2048, None, cuda)
    _submit(cuda, c3_scene, sfc)
    got = sfc.download()
    undefined = []
    ref = render_scene(load_oracle(fast=True), c3_scene, asserting=undefined)
    sfc.deinit()
    assert not undefined, "the workload must stay inside the reference's defined inputs"
    assert np.array_equal(got, ref), f"{int((got != ref).sum())} bytes differ"


def test_c3_full_run_chunked_equals_single_batch(cuda, c3_scene):
    sfc = Surface(Format.rgba, 2048, 2048, None, cuda)
    cuda.set_chunk(0)
    _submit(cuda, c3_scene, sfc)
    whole = sfc.download().copy()
    sfc.paint_pixel(host.Pixel.rgba(0, 0, 0, 0))
    cuda.set_chunk(5000)
    _submit(cuda, c3_scene, sfc)
    parts = sfc.download().copy()
    cuda.set_chunk(32768)
    sfc.deinit()
    assert np.array_equal(whole, parts)
    assert int((whole.reshape(-1, 4)[:, 3] > 0).sum()) > 2048 * 2048 // 2


# ------------------------------------------------------------------------------------------------ C4
W4 = 8192
STRIPS = (0, 4096, 8128)  # multiples of 64: the dither matrices line up between a strip and the full surface
STRIP_H = 2
BITS = {Format.rgba: 32, Format.rgb: 32, Format.alpha8: 8, Format.alpha4: 4, Format.alpha2: 2, Format.alpha1: 1}


def _prefill(fmt):
    """hash32(x, y) content for a W4 x W4 surface in the raw layout of `fmt` (RGBA premultiplied)."""
    bits = BITS[fmt]
    n_words = W4 * W4 * bits // 32
    i = np.arange(n_words, dtype=np.uint64)
    h = (i * np.uint64(0x9E3779B1)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(15)
    h = (h * np.uint64(0x85EBCA77)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(13)
    raw = h.astype(np.uint32)
    if fmt == Format.rgba:
        a = (raw >> 24).astype(np.uint32)
        out = (a << 24)
        for sh in (0, 8, 16):
            out |= ((((raw >> sh) & 255) * a) // 255) << sh
        raw = out
    elif fmt == Format.rgb:
        raw &= np.uint32(0x00FFFFFF)
    return raw.view(np.uint8)


def _strip(raw, fmt, y0):
    row = W4 * BITS[fmt] // 8
    return raw[y0 * row:(y0 + STRIP_H) * row]


def _stops(g):
    g.add_stop(0.0, {"rgba": (1, 0, 0, 1)})
    g.add_stop(0.5, {"rgba": (0, 1, 0, 0.5)})
    g.add_stop(1.0, {"rgba": (0, 0, 1, 1)})
    return g


def _sources(z, y0):
    """name -> Param evaluated at rows y0.. (gradients carry the translation, pixels and dither-over-pixel do not need it)."""
    def shifted(g):
        if y0:
            g.set_transformation(z.Transformation().translate(0.0, -float(y0)))
        return g
    lin = lambda **kw: shifted(_stops(z.Gradient.linear(0, 0, W4, W4, **kw)))  # noqa: E731
    rad = shifted(_stops(z.Gradient.radial(W4 / 2, W4 / 2, 0, W4 / 2, W4 / 2, W4 / 2)))
    con = shifted(_stops(z.Gradient.conic(W4 / 2, W4 / 2, 0)))
    return {
        "pixel": z.Param.pixel(z.Pixel.rgba(90, 40, 10, 128)),
        "linear": z.Param.gradient(lin()),
        "linear_srgb": z.Param.gradient(lin(method=abi.Interp.srgb)),
        "linear_hsl": z.Param.gradient(lin(method=abi.Interp.hsl)),
        "radial": z.Param.gradient(rad),
        "conic": z.Param.gradient(con),
        "bayer": z.Param.dither(z.Dither(abi.DitherType.bayer, lin(), 8)),
        "blue_noise": z.Param.dither(z.Dither(abi.DitherType.blue_noise, lin(), 4)),
    }


SRC_NAMES = ["pixel", "linear", "radial", "conic", "bayer", "blue_noise", "linear_srgb", "linear_hsl"]
FLOAT_ONLY = {Operator.color_dodge, Operator.color_burn, Operator.soft_light, Operator.hue, Operator.saturation, Operator.color,
              Operator.luminosity}


def _samples(raw_bytes, fmt):
    """Per-pixel samples of a strip: (n, 4) bytes for the 32-bit formats, else the n-bit alpha samples (LSB-first packing)."""
    bits = BITS[fmt]
    if bits == 32:
        return raw_bytes.reshape(-1, 4).astype(np.int32)
    if bits == 8:
        return raw_bytes.astype(np.int32)
    b = np.unpackbits(raw_bytes, bitorder="little").reshape(-1, bits).astype(np.int32)
    return (b << np.arange(bits)).sum(axis=1)


def _check_composite(cuda, oracle, fmt, prefill, big, op, precision, src_name, tol):
    """tol: allowed difference per stored sample, in units of the destination's own quantisation (a +-1 LSB difference of the
    8-bit float result moves an n-bit sample by at most one level)."""
    zc, zo = specs.bind(cuda), specs.bind(oracle)
    big.upload(prefill)
    zc.SurfaceCompositor.run(big, 0, 0, [zc.Operation(op, src=_sources(zc, 0)[src_name])], precision=precision)
    got = big.download()
    for y0 in STRIPS:
        small = zo.Surface(fmt, W4, STRIP_H)
        small.upload(_strip(prefill, fmt, y0).copy())
        zo.SurfaceCompositor.run(small, 0, 0, [zo.Operation(op, src=_sources(zo, y0)[src_name])], precision=precision)
        ref = small.download()
        g = _strip(got, fmt, y0)
        if tol == 0:
            assert np.array_equal(g, ref), f"{fmt.name} {op.name} {precision.name} {src_name} rows {y0}: {int((g != ref).sum())} bytes differ"
        else:
            d = np.abs(_samples(g, fmt) - _samples(ref, fmt))
            if BITS[fmt] == 32 and fmt == Format.rgb:
                d = d[:, :3]  # the padding byte is undefined after compositing (surface.zig:1732)
            assert d.max() <= tol, f"{fmt.name} {op.name} {precision.name} {src_name} rows {y0}: max diff {d.max()}"


def test_c4_rgba_all_operators(cuda, oracle):
    prefill = _prefill(Format.rgba)
    big = Surface(Format.rgba, W4, W4, None, cuda)
    k = 0
    for op in Operator:
        for precision in (Precision.integer, Precision.float):
            if precision == Precision.integer and op in FLOAT_ONLY:
                continue
            src_name = SRC_NAMES[k % len(SRC_NAMES)]
            k += 1
            # integer pipeline with a single-pixel source is exact; anything through f32 is allowed +-1 LSB
            tol = 0 if (precision == Precision.integer and src_name == "pixel") else 1
            _check_composite(cuda, oracle, Format.rgba, prefill, big, op, precision, src_name, tol)
    big.deinit()


C4_OPS = [(Operator.src_over, Precision.integer), (Operator.src, Precision.integer), (Operator.clear, Precision.integer),
          (Operator.dst_in, Precision.integer), (Operator.xor, Precision.integer), (Operator.multiply, Precision.integer),
          (Operator.plus, Precision.float), (Operator.src_over, Precision.float), (Operator.soft_light, Precision.float),
          (Operator.hue, Precision.float)]


@pytest.mark.parametrize("fmt", [Format.rgba, Format.rgb, Format.alpha8, Format.alpha4, Format.alpha2, Format.alpha1])
def test_c4_every_format_source_and_operator_class(cuda, oracle, fmt):
    """Every destination format x every source kind (pixel, three gradient types, two dither matrices, sRGB and HSL
    interpolation) x write-only / Porter-Duff / separable / float-only / non-separable operators in both precisions."""
    prefill = _prefill(fmt)
    big = Surface(fmt, W4, W4, None, cuda)
    for src_name in SRC_NAMES:
        for op, precision in C4_OPS:
            tol = 0 if (precision == Precision.integer and src_name == "pixel") else 1
            _check_composite(cuda, oracle, fmt, prefill, big, op, precision, src_name, tol)
    big.deinit()


def test_c4_partial_rows_and_offsets(cuda, oracle):
    """Row ranges that do not start on a byte / 16-byte boundary (odd widths, packed formats) and surface sources at offsets."""
    zc, zo = specs.bind(cuda), specs.bind(oracle)
    rng = np.random.default_rng(7)
    for fmt in (Format.alpha1, Format.alpha2, Format.alpha4, Format.alpha8, Format.rgba):
        for w, h in ((601, 37), (17, 5), (130, 3)):
            data = rng.integers(0, 256, abi.surface_byte_len(fmt, w, h), dtype=np.uint8)
            if fmt == Format.rgba:
                px = data.reshape(-1, 4).astype(np.int32)
                px[:, :3] = px[:, :3] * px[:, 3:4] // 255
                data = px.astype(np.uint8).reshape(-1)
            for src_name in ("pixel", "linear", "bayer"):
                for op, precision in ((Operator.src_over, Precision.integer), (Operator.dst_out, Precision.float), (Operator.src, Precision.integer)):
                    res = []
                    for z in (zc, zo):
                        s = z.Surface(fmt, w, h)
                        s.upload(data.copy())
                        g = z.Gradient.linear(0, 0, w, h)
                        _stops(g)
                        prm = {"pixel": z.Param.pixel(z.Pixel.rgba(90, 40, 10, 128)), "linear": z.Param.gradient(g),
                               "bayer": z.Param.dither(z.Dither(abi.DitherType.bayer, g, 8 if BITS[fmt] >= 8 else BITS[fmt]))}[src_name]
                        z.SurfaceCompositor.run(s, 0, 0, [z.Operation(op, src=prm)], precision=precision)
                        res.append(s.download().copy())
                    n_px = w * h
                    d = np.abs(_samples(res[0], fmt)[:n_px] - _samples(res[1], fmt)[:n_px])
                    tol = 0 if (precision == Precision.integer and src_name == "pixel") else 1
                    assert d.max() <= tol, f"{fmt.name} {w}x{h} {src_name} {op.name} {precision.name}: max diff {d.max()}"


# ------------------------------------------------------------------------------------------------ C5
def test_c5_scene_batch_sharding(cuda):
    n_scenes, size = 16, 1024
    scenes = [workloads.mixed_scene(s, size) for s in range(n_scenes)]
    lib = load_oracle(fast=True)

    def render(indices):
        sfcs = {s: Surface(Format.rgba, size, size, None, cuda) for s in indices}
        for s in indices:  # one batch spanning many surfaces
            _submit(cuda, scenes[s], sfcs[s])
        out = {s: sfcs[s].download().copy() for s in indices}
        for sfc in sfcs.values():
            sfc.deinit()
        return out

    whole = render(list(range(n_scenes)))
    for s in range(n_scenes):  # every scene against the oracle (fills, strokes through the unit stroker, gradient fills)
        assert np.array_equal(whole[s], render_scene(lib, scenes[s])), f"scene {s} differs from the oracle"
    all_sums = [sharding.surface_checksum(whole[s]) for s in range(n_scenes)]
    for world in (2, 4):
        merged = {}
        for rank in range(world):
            part = render(sharding.scenes_of_rank(n_scenes, world, rank))
            merged.update({s: sharding.surface_checksum(b) for s, b in part.items()})
        assert [merged[s] for s in range(n_scenes)] == all_sums


def test_c5_more_scenes_match_the_oracle(cuda):
    """64 further scenes of the config-5 generator (indices 1000 ...), rendered as ONE batch over 64 surfaces: integer work is
    compared byte for byte; the gradient fills go through floating point, where the north star allows 1 LSB (none is used)."""
    size, first, n = 1024, 1000, 64
    scenes = [workloads.mixed_scene(first + k, size) for k in range(n)]
    lib = load_oracle(fast=True)
    sfcs = [Surface(Format.rgba, size, size, None, cuda) for _ in range(n)]
    for sc, sfc in zip(scenes, sfcs):
        _submit(cuda, sc, sfc)
    got = [sfc.download().copy() for sfc in sfcs]
    for sfc in sfcs:
        sfc.deinit()
    for k in range(n):
        ref = render_scene(lib, scenes[k])
        diff = np.abs(got[k].astype(np.int16) - ref.astype(np.int16))
        assert int(diff.max()) <= 1, f"scene {first + k}: {int((diff > 1).sum())} bytes differ by more than 1 LSB"
        assert int((diff != 0).sum()) == 0, f"scene {first + k}: {int((diff != 0).sum())} bytes differ by 1 LSB"
