"""Shared by the randomised parity tests: render a workloads.Scene through the C ABI and through the oracle and compare.

On a mismatch every draw call is rendered ALONE by both sides, so that the failure names the calls that differ instead of
hiding a whole seed behind one number (a call that only differs on top of earlier content is reported as "order dependent")."""
import ctypes as C

import numpy as np

from tests.oracle_backend import load_oracle, render_scene
from z2d_b200 import abi
from z2d_b200.abi import Format
from z2d_b200.host import Pixel, Surface


def render_cuda(cuda, scene, lo=0, hi=None, sfc=None, keep=None):
    own = sfc is None
    if own:
        sfc = Surface(Format.rgba, scene.width, scene.height, None, cuda)
    else:
        sfc.paint_pixel(Pixel.rgba(0, 0, 0, 0))
    cmds = scene.draw_cmds(sfc.handle, lo, hi)
    if keep is not None:
        cmds = np.ascontiguousarray(cmds[np.asarray(keep, dtype=bool)])
    statuses = np.zeros(len(cmds), dtype=np.int32)
    cuda._check(cuda.lib.z2d_submit(cuda.ctx, cmds.ctypes.data_as(C.POINTER(abi.DrawCmdPOD)), len(cmds),
                                    statuses.ctypes.data_as(C.POINTER(C.c_int32))))
    assert (statuses == 0).all(), f"statuses {np.unique(statuses)}"
    out = sfc.download().copy()
    if own:
        sfc.deinit()
    return out


def assert_scene_matches(cuda, scene, min_covered=None, max_undefined=0):
    """Byte-for-byte comparison of every call of the scene.  max_undefined: how many calls may be left out because the REFERENCE
    has no defined result for them (one of its own debug.asserts fires: it panics in safe builds) -- they are identified by the
    oracle, dropped from both renders, and everything else is still compared."""
    lib = load_oracle(fast=True)
    undefined = []
    ref = render_scene(lib, scene, asserting=undefined)
    keep = None
    if undefined:
        assert len(undefined) <= max_undefined, f"{len(undefined)} calls trip a reference assertion: {undefined[:10]}"
        keep = np.ones(scene.n, dtype=bool)
        keep[undefined] = False
        ref = render_scene(lib, scene, keep=keep)
    got = render_cuda(cuda, scene, keep=keep)
    if min_covered is not None:
        assert int((ref.reshape(-1, 4)[:, 3] > 0).sum()) > min_covered, "the scene covers too little to mean anything"
    if np.array_equal(got, ref):
        return
    bad_px = int((got.reshape(-1, 4) != ref.reshape(-1, 4)).any(axis=1).sum())
    differing = []
    sfc = Surface(Format.rgba, scene.width, scene.height, None, cuda)
    for i in range(scene.n):
        if keep is not None and not keep[i]:
            continue
        g = render_cuda(cuda, scene, i, i + 1, sfc)
        r = render_scene(lib, scene, i, i + 1)
        if not np.array_equal(g, r):
            differing.append((i, int((g.reshape(-1, 4) != r.reshape(-1, 4)).any(axis=1).sum())))
    sfc.deinit()
    what = f"draw calls that differ on their own (index, pixels): {differing[:20]}" if differing else "no single call differs: order dependent"
    raise AssertionError(f"{bad_px} pixels differ from the oracle; {what}")
