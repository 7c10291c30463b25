"""Fixed cost of one batch: the first n fills of config 2 as ONE batch, wall time (submit + sync) and device stage times."""
import ctypes as C
import sys
import time

import torch

sys.path.insert(0, ".")
from z2d_b200 import abi, workloads  # noqa: E402
from z2d_b200.cuda_backend import CudaBackend  # noqa: E402
from z2d_b200.host import Pixel, Surface  # noqa: E402

stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
cb = CudaBackend(0, stream=stream.cuda_stream)
cb.set_chunk(0)
scene = workloads.cubic_paths_scene(100_000, 4096)
sfc = Surface(abi.Format.rgba, 4096, 4096, None, cb)
cmds = scene.draw_cmds(sfc.handle)
ptr = cmds.ctypes.data_as(C.POINTER(abi.DrawCmdPOD))
for n in (1024, 4096, 8192, 16384, 32768, 100000):
    best = None
    for it in range(6):
        cb.sync()
        t0 = time.perf_counter()
        cb.submit(ptr, n)
        t1 = time.perf_counter()
        cb.sync()
        t2 = time.perf_counter()
        st = cb.stats()
        rec = (n, round((t1 - t0) * 1e3, 3), round((t2 - t1) * 1e3, 3), round(st["ms_total"], 3), round(st["ms_flatten"], 3), round(st["ms_bin"], 3),
               round(st["ms_lists"], 3), round(st["ms_raster"], 3))
        if it >= 2 and (best is None or rec[1] + rec[2] < best[1] + best[2]):
            best = rec
    print("n %6d record %.3f flush+sync %.3f | device total %.3f flatten %.3f bin %.3f lists %.3f raster %.3f" % best, flush=True)
