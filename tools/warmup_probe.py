"""How does the device time of one step evolve from a cold start? (fresh box, first process)"""
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
cb.submit(cmds.ctypes.data_as(C.POINTER(abi.DrawCmdPOD)), scene.n)
cb.sync()
zero = Pixel.rgba(0, 0, 0, 0)
t0 = time.perf_counter()
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 80):
    sfc.paint_pixel(zero)
    cb.replay()
    st = cb.stats()
    if it < 12 or it % 10 == 0:
        print(f"step {it:3d} t={time.perf_counter() - t0:6.2f}s raster {st['ms_raster']:.3f} total {st['ms_total']:.3f}", flush=True)
