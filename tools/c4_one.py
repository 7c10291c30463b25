"""One config-4 cell, a few launches (for ncu): python tools/c4_one.py rgba linear none src_over integer"""
import sys

import torch

sys.path.insert(0, ".")
import bench_extra  # noqa: E402
from z2d_b200.abi import Format, Operator, Precision  # noqa: E402
from z2d_b200.cuda_backend import CudaBackend  # noqa: E402
from z2d_b200.host import Operation, Surface, SurfaceCompositor  # noqa: E402

fmt, sname, dname, op, prec = Format[sys.argv[1]], sys.argv[2], sys.argv[3], Operator[sys.argv[4]], Precision[sys.argv[5]]
n = int(sys.argv[6]) if len(sys.argv) > 6 else 8192
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
cb = CudaBackend(0, stream=stream.cuda_stream)
sfc = Surface(fmt, n, n, None, cb)
sfc.upload(bench_extra._prefill(fmt, n))
bpc = 8 if bench_extra.BITS[fmt] >= 8 else bench_extra.BITS[fmt]
prm = bench_extra.c4_sources(n, bpc)[(sname, dname)]
for _ in range(4):
    SurfaceCompositor.run(sfc, 0, 0, [Operation(op, src=prm)], precision=prec)
cb.sync()
