"""Fixed cost of one batch: device stage timings and host wall time for batches of different sizes."""
import ctypes as C
import sys
import time

import torch

sys.path.insert(0, ".")
from z2d_b200 import abi, workloads  # noqa: E402
from z2d_b200.cuda_backend import CudaBackend  # noqa: E402
from z2d_b200.host import Surface  # noqa: E402

stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
cb = CudaBackend(0, stream=stream.cuda_stream)
cb.set_chunk(0)
scene = workloads.cubic_paths_scene(100_000, 4096)
sfc = Surface(abi.Format.rgba, 4096, 4096, None, cb)
for n in (1, 64, 1024, 4096, 16384, 100000):
    cmds = scene.draw_cmds(sfc.handle, 0, n)
    p = cmds.ctypes.data_as(C.POINTER(abi.DrawCmdPOD))
    best = None
    for it in range(5):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        cb.submit(p, n)
        t1 = time.perf_counter()
        cb.sync()
        t2 = time.perf_counter()
        st = cb.stats()
        rec = ((t2 - t1) * 1e3, (t1 - t0) * 1e3, st["ms_total"], st["ms_flatten"], st["ms_bin"], st["ms_lists"], st["ms_raster"])
        if it >= 1 and (best is None or rec[0] < best[0]):
            best = rec
    print(f"n={n:6d}: flush+sync wall {best[0]:7.3f} ms, record {best[1]:6.3f} ms | device total {best[2]:7.3f} "
          f"(flatten {best[3]:.3f} bin {best[4]:.3f} lists {best[5]:.3f} raster {best[6]:.3f})", flush=True)
