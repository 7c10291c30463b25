"""Compositor kernel A/B: 8192^2 RGBA8 src_over with a single-pixel source (BASELINE config 4 shape), CUDA-event timed."""
import os
import sys

import torch

sys.path.insert(0, ".")
from z2d_b200 import abi  # noqa: E402
from z2d_b200.cuda_backend import CudaBackend  # noqa: E402
from z2d_b200.host import Operation, Param, Pixel, Surface, SurfaceCompositor  # noqa: E402

stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
cb = CudaBackend(0, stream=stream.cuda_stream)
W = 8192
dst = Surface(abi.Format.rgba, W, W, None, cb)
src = Surface(abi.Format.rgba, W, W, Pixel.rgba(10, 20, 30, 128), cb)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
res = {}
for name, ops, prec in (("px_src_over", [Operation(abi.Operator.src_over, src=Param.pixel(Pixel.rgba(90, 40, 10, 128)))], abi.Precision.integer),
                        ("sfc_src_over", [Operation(abi.Operator.src_over, src=Param.surface(src))], abi.Precision.integer),
                        ("px_multiply_f", [Operation(abi.Operator.multiply, src=Param.pixel(Pixel.rgba(90, 40, 10, 128)))], abi.Precision.float)):
    ts = []
    for it in range(12):
        flush.zero_()  # evict the surface from L2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        SurfaceCompositor.run(dst, 0, 0, ops, precision=prec)
        e1.record()
        torch.cuda.synchronize()
        if it >= 2:
            ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    by = W * W * (8 if name != "sfc_src_over" else 12)
    res[name] = (ms, by / ms / 1e6)
print(os.environ.get("Z2D_CUDA_LIB", "default").split("/")[-1], " ".join(f"{k}: {v[0]:.4f} ms {v[1]:.0f} GB/s" for k, v in res.items()), flush=True)
