"""End-to-end time of the headline workload for several recorder chunk sizes (run on the GPU box)."""
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
scene = workloads.cubic_paths_scene(100_000, 4096)
sfc = Surface(abi.Format.rgba, 4096, 4096, None, cb)
cmds = scene.draw_cmds(sfc.handle)
cmds_p = cmds.ctypes.data_as(C.POINTER(abi.DrawCmdPOD))
nbytes = sfc.byte_len()
host_out = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
host_ptr = C.c_void_p(host_out.data_ptr())
zero = Pixel.rgba(0, 0, 0, 0)
for chunk in (0, 4096, 8192, 16384, 25000, 32768, 50000):
    cb.set_chunk(chunk)
    ts = []
    for it in range(6):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        sfc.paint_pixel(zero)
        cb.submit(cmds_p, scene.n)
        t1 = time.perf_counter()
        cb._check(cb.lib.z2d_surface_download(sfc.handle, host_ptr, nbytes))
        t2 = time.perf_counter()
        if it >= 2:
            ts.append(((t2 - t0) * 1e3, (t1 - t0) * 1e3))
    print(f"chunk {chunk:6d}: e2e {min(t[0] for t in ts):7.2f} ms (record {min(t[1] for t in ts):6.2f} ms)", flush=True)
