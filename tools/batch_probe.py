import ctypes as C, sys
import torch
sys.path.insert(0, ".")
from z2d_b200 import abi, workloads
from z2d_b200.cuda_backend import CudaBackend
from z2d_b200.host import Surface
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
cb = CudaBackend(0, stream=stream.cuda_stream); cb.set_chunk(0)
scene = workloads.cubic_paths_scene(100_000, 4096)
sfc = Surface(abi.Format.rgba, 4096, 4096, None, cb)
cmds = scene.draw_cmds(sfc.handle)
ptr = cmds.ctypes.data_as(C.POINTER(abi.DrawCmdPOD))
for it in range(4):
    cb.submit(ptr, 2048); cb.sync()
