"""Batching must not change results: the recorder hands chunks of draws to a worker thread (z2d_ctx_set_chunk) and the
tile pipeline composites a whole batch per tile, yet the output has to equal executing every call in order."""
import ctypes as C

import numpy as np
import pytest

from z2d_b200 import abi, host, workloads
from z2d_b200.abi import Format
from z2d_b200.host import Pixel, Surface

pytestmark = pytest.mark.gpu


def _render(cuda, scene, chunk, size):
    cuda.set_chunk(chunk)
    sfc = Surface(abi.Format.rgba, size, size, None, cuda)
    cmds = scene.draw_cmds(sfc.handle)
    cuda.submit(cmds.ctypes.data_as(C.POINTER(abi.DrawCmdPOD)), scene.n)
    out = sfc.download().copy()
    sfc.deinit()
    cuda.set_chunk(32768)
    return out


def _oracle(scene, size, n):
    from tests.oracle_backend import load_oracle
    lib = load_oracle(fast=True)
    buf = np.zeros(size * size * 4, dtype=np.uint8)
    cmds = scene.draw_cmds(0, 0, n)
    P = C.POINTER
    for i in range(n):
        rc = lib.z2d_ref_fill(buf.ctypes.data_as(C.c_void_p), int(abi.Format.rgba), size, size,
                              C.cast(C.c_void_p(int(cmds["pattern"][i])), P(abi.PatternPOD)),
                              C.cast(C.c_void_p(int(cmds["nodes"][i])), P(abi.Node)), int(cmds["n_nodes"][i]),
                              C.cast(C.c_void_p(int(cmds["fill"][i])), P(abi.FillOptsPOD)))
        assert rc == 0
    return buf


@pytest.mark.parametrize("chunk", [1, 7, 64, 1000])
def test_chunked_submission_equals_single_batch(cuda, chunk):
    size = 384
    scene = workloads.cubic_paths_scene(1500, size, seed=0x5EED0001, r_log2=(3.0, 6.0))
    whole = _render(cuda, scene, 0, size)
    parts = _render(cuda, scene, chunk, size)
    assert np.array_equal(whole, parts)


def test_ordered_translucent_scene_matches_oracle(cuda):
    size = 512
    scene = workloads.cubic_paths_scene(3000, size, seed=0x5EED0002, r_log2=(3.0, 7.0))
    got = _render(cuda, scene, 256, size)
    ref = _oracle(scene, size, scene.n)
    assert np.array_equal(got, ref), f"{int((got != ref).sum())} bytes differ"


def test_interleaved_surfaces_keep_per_surface_order(cuda):
    """Draws alternate between two surfaces; each surface must look as if only its own draws were issued, in order."""
    size = 256
    scene = workloads.cubic_paths_scene(600, size, seed=0x5EED0003, r_log2=(3.0, 6.0))
    a = Surface(abi.Format.rgba, size, size, None, cuda)
    b = Surface(abi.Format.rgba, size, size, None, cuda)
    cmds = scene.draw_cmds(a.handle)
    cmds["surface"][1::2] = b.handle.value if hasattr(b.handle, "value") else int(b.handle)
    cuda.set_chunk(50)
    cuda.submit(cmds.ctypes.data_as(C.POINTER(abi.DrawCmdPOD)), scene.n)
    got_a, got_b = a.download().copy(), b.download().copy()
    cuda.set_chunk(32768)
    # references: each half alone on a fresh surface
    for got, sl in ((got_a, slice(0, None, 2)), (got_b, slice(1, None, 2))):
        ref_sfc = Surface(abi.Format.rgba, size, size, None, cuda)
        sub = scene.draw_cmds(ref_sfc.handle)[sl].copy()
        cuda.submit(sub.ctypes.data_as(C.POINTER(abi.DrawCmdPOD)), len(sub))
        assert np.array_equal(got, ref_sfc.download())
        ref_sfc.deinit()
    a.deinit()
    b.deinit()


def test_parallel_recorder_equals_one_by_one_calls(cuda):
    """z2d_submit records long runs of plain fills with several host threads; statuses and pixels must equal issuing the same
    calls one by one through z2d_fill -- including calls that fail (unclosed path, non-premultiplied source) or record
    nothing (empty node list)."""
    size, n = 320, 5000
    scene = workloads.cubic_paths_scene(n, size, seed=0x5EED0004, r_log2=(2.0, 5.0))
    a = Surface(abi.Format.rgba, size, size, None, cuda)
    cmds = scene.draw_cmds(a.handle)
    # break some calls
    scene.nodes["tag"][scene.node_off[7] + 5] = int(abi.NodeTag.line_to)     # close_path -> line_to: PathNotClosed
    scene.patterns["pixel"]["r"][11] = 255                                   # r > a: not premultiplied
    scene.patterns["pixel"]["a"][11] = 10
    cmds["n_nodes"][13] = 0                                                  # empty node list: silent no-op
    scene.nodes["tag"][scene.node_off[17]] = int(abi.NodeTag.line_to)        # line_to without a current point: InvalidState
    statuses = np.zeros(n, dtype=np.int32)
    rc = cuda.lib.z2d_submit(cuda.ctx, cmds.ctypes.data_as(C.POINTER(abi.DrawCmdPOD)), n, statuses.ctypes.data_as(C.POINTER(C.c_int32)))
    got = a.download().copy()
    assert rc == abi.E_PATH_NOT_CLOSED  # the first failing call
    assert statuses[7] == abi.E_PATH_NOT_CLOSED and statuses[11] == abi.E_PIXEL_SOURCE_NOT_PREMULTIPLIED
    assert statuses[13] == abi.OK and statuses[17] == abi.E_INVALID_STATE
    assert int((statuses != 0).sum()) == 3
    # the same calls one by one
    b = Surface(abi.Format.rgba, size, size, None, cuda)
    c2 = scene.draw_cmds(b.handle)
    c2["n_nodes"][13] = 0
    P = C.POINTER
    ref_status = []
    for i in range(n):
        ref_status.append(cuda.lib.z2d_fill(cuda.ctx, b.handle, C.cast(C.c_void_p(int(c2["pattern"][i])), P(abi.PatternPOD)),
                                            C.cast(C.c_void_p(int(c2["nodes"][i])), P(abi.Node)), int(c2["n_nodes"][i]),
                                            C.cast(C.c_void_p(int(c2["fill"][i])), P(abi.FillOptsPOD))))
    ref = b.download()
    assert list(statuses) == ref_status
    assert np.array_equal(got, ref)
    a.deinit()
    b.deinit()


def test_composite_surface_params_must_cover_the_rectangle(cuda):
    """z2d_composite clips against ops[0].src only (compositor.zig:347-374); any other surface parameter that is smaller than the
    composited rectangle, and the destination as a parameter of itself at an offset, is refused instead of read out of bounds."""
    import pytest as _pytest
    from tests import specs
    z = specs.bind(cuda)
    dst = z.Surface(Format.rgba, 32, 32)
    big = z.SurfacePixel(host.Pixel.rgba(10, 20, 30, 40), 32, 32)
    small = z.SurfacePixel(host.Pixel.rgba(10, 20, 30, 40), 16, 16)
    with _pytest.raises(abi.InvalidArg):  # second operation reads a smaller source
        z.SurfaceCompositor.run(dst, 0, 0, [z.Operation(abi.Operator.src_over, src=z.Param.surface(big)),
                                            z.Operation(abi.Operator.src_over, src=z.Param.surface(small))])
    with _pytest.raises(abi.InvalidArg):  # dst override smaller than the destination region
        z.SurfaceCompositor.run(dst, 0, 0, [z.Operation(abi.Operator.dst_in, dst=z.Param.surface(small), src=z.Param.surface(big))])
    with _pytest.raises(abi.InvalidArg):  # the destination shifted onto itself
        z.SurfaceCompositor.run(dst, 3, 0, [z.Operation(abi.Operator.src_over, src=z.Param.surface(dst))])
    z.SurfaceCompositor.run(dst, 0, 0, [z.Operation(abi.Operator.plus, src=z.Param.surface(dst))])  # pixel for pixel is fine
    assert not dst.download().any()


@pytest.mark.parametrize("dst_fmt,src_fmt", [(Format.rgba, Format.rgba), (Format.rgba, Format.alpha8), (Format.alpha8, Format.rgba),
                                             (Format.alpha4, Format.alpha8), (Format.rgb, Format.alpha2)])
def test_surface_composite_at_negative_and_positive_offsets_matches_oracle(cuda, oracle, dst_fmt, src_fmt):
    """Surface.composite / SurfaceCompositor.run with a surface source placed partly outside the destination on every side
    (compositor.zig:347-374: negative offsets start the source at -offset), byte for byte."""
    from tests import specs
    rng = np.random.default_rng(11)

    def content(fmt, w, h):
        data = rng.integers(0, 256, abi.surface_byte_len(fmt, w, h), dtype=np.uint8)
        if fmt == Format.rgba:
            px = data.reshape(-1, 4).astype(np.int32)
            px[:, :3] = px[:, :3] * px[:, 3:4] // 255
            data = px.astype(np.uint8).reshape(-1)
        return data

    d0, s0 = content(dst_fmt, 37, 29), content(src_fmt, 23, 31)
    for dx, dy in ((-5, -3), (20, -7), (-11, 9), (30, 25), (0, 0), (3, 2)):
        for op in (abi.Operator.src_over, abi.Operator.xor, abi.Operator.dst_in):
            res = []
            for z in (specs.bind(cuda), specs.bind(oracle)):
                dst, src = z.Surface(dst_fmt, 37, 29), z.Surface(src_fmt, 23, 31)
                dst.upload(d0.copy())
                src.upload(s0.copy())
                z.SurfaceCompositor.run(dst, dx, dy, [z.Operation(op, src=z.Param.surface(src))])
                res.append(dst.download().copy())
            if dst_fmt == Format.rgb:
                res = [r.reshape(-1, 4)[:, :3] for r in res]
            assert np.array_equal(res[0], res[1]), f"{dst_fmt.name} <- {src_fmt.name} at ({dx},{dy}) {op.name}"


def test_stroke_pool_and_counted_edges_share_the_edge_array():
    """The unit stroker writes stroke edges into a pool at the front of the edge array before the counted total of the batch is
    known (z2d_lib.cu run_pipeline).  On a FRESH context: a small stroke batch sizes the pool; a batch with a few strokes and
    thousands of fills then needs a larger array after the pool was written (it must survive the reallocation); a batch with
    many more strokes than the capacities allow is redone with larger ones.  Every call is compared with the oracle."""
    from tests.fuzz_util import assert_scene_matches
    from z2d_b200.cuda_backend import CudaBackend
    cb = CudaBackend()
    cb.set_chunk(0)  # one batch per scene
    size = 512
    assert_scene_matches(cb, workloads.mixed_scene(0, size, n_fills=4, n_strokes=40, n_gradients=0), max_undefined=3)
    assert_scene_matches(cb, workloads.mixed_scene(1, size, n_fills=3000, n_strokes=30, n_gradients=0), max_undefined=3)
    assert_scene_matches(cb, workloads.mixed_scene(2, size, n_fills=4, n_strokes=900, n_gradients=0), max_undefined=3)
    assert_scene_matches(cb, workloads.mixed_scene(3, size, n_fills=200, n_strokes=200, n_gradients=4), max_undefined=3)
