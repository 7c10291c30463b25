"""Batching must not change results: the recorder hands chunks of draws to a worker thread (z2d_ctx_set_chunk) and the
tile pipeline composites a whole batch per tile, yet the output has to equal executing every call in order."""
import ctypes as C

import numpy as np
import pytest

from z2d_b200 import abi, workloads
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
