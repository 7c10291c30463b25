"""Band views (second row): a band whose rows live inside a full canvas -- of the same context, or of ANOTHER process
through a CUDA IPC handle (the multi-GPU form: tile write-backs go over NVLink into rank 0's canvas).  The second process runs on
the same GPU when the box has only one (IPC works across processes on one device); with two GPUs it uses the second."""
import ctypes as C
import multiprocessing as mp

import numpy as np
import pytest

from z2d_b200 import abi, workloads
from z2d_b200.abi import Format
from z2d_b200.host import Surface

pytestmark = pytest.mark.gpu
SIZE, N = 512, 400


def _scene():
    return workloads.cubic_paths_scene(N, SIZE, seed=77, r_log2=(3.0, 7.0))


def _submit(cb, scene, sfc):
    cmds = scene.draw_cmds(sfc.handle)
    cb.submit(cmds.ctypes.data_as(C.POINTER(abi.DrawCmdPOD)), scene.n)


def test_band_views_of_one_canvas_equal_the_full_render(cuda):
    scene = _scene()
    full = Surface(Format.rgba, SIZE, SIZE, None, cuda)
    _submit(cuda, scene, full)
    ref = full.download().copy()
    canvas = Surface(Format.rgba, SIZE, SIZE, None, cuda)
    views = [canvas.band_view(y0, rows) for y0, rows in ((0, 128), (128, 256), (384, 128))]
    for v in views:
        _submit(cuda, scene, v)
    cuda.sync()
    assert np.array_equal(canvas.download(), ref)
    with pytest.raises(abi.InvalidArg):
        canvas.band_view(8, 64)  # not on a tile row
    with pytest.raises(abi.InvalidArg):
        views[0].band_view(0, 16)  # a view of a view
    for v in views:
        v.deinit()
    canvas.deinit()
    full.deinit()


def _peer(handle, device, y0, rows, done):
    from z2d_b200.cuda_backend import CudaBackend
    cb = CudaBackend(device)
    band = Surface.open_peer_band(handle, Format.rgba, SIZE, SIZE, y0, rows, cb)
    _submit(cb, _scene(), band)
    cb.sync()
    band.deinit()
    cb.close()
    done.put(True)


def test_peer_band_written_from_another_process(cuda):
    import torch
    scene = _scene()
    full = Surface(Format.rgba, SIZE, SIZE, None, cuda)
    _submit(cuda, scene, full)
    ref = full.download().copy()
    canvas = Surface(Format.rgba, SIZE, SIZE, None, cuda)
    handle = canvas.ipc_export()
    ctx = mp.get_context("spawn")
    done = ctx.Queue()
    device = 1 if torch.cuda.device_count() > 1 else 0
    proc = ctx.Process(target=_peer, args=(handle, device, 256, 256, done))
    proc.start()
    own = canvas.band_view(0, 256)
    _submit(cuda, scene, own)
    cuda.sync()
    assert done.get(timeout=120)
    proc.join(timeout=60)
    assert proc.exitcode == 0
    assert np.array_equal(canvas.download(), ref)
    own.deinit()
    canvas.deinit()
    full.deinit()
