import sys
import torch
sys.path.insert(0, ".")
import bench_extra
from z2d_b200.abi import Format, Operator, Precision
from z2d_b200.cuda_backend import CudaBackend
from z2d_b200.host import Operation, Surface, SurfaceCompositor
fmt, sname, dname, op, prec = Format[sys.argv[1]], sys.argv[2], sys.argv[3], Operator[sys.argv[4]], Precision[sys.argv[5]]
n = 8192
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
cb = CudaBackend(0, stream=stream.cuda_stream)
sfc = Surface(fmt, n, n, None, cb); sfc.upload(bench_extra._prefill(fmt, n))
bpc = 8 if bench_extra.BITS[fmt] >= 8 else bench_extra.BITS[fmt]
prm = bench_extra.c4_sources(n, bpc)[(sname, dname)]
run = lambda: SurfaceCompositor.run(sfc, 0, 0, [Operation(op, src=prm)], precision=prec)
for _ in range(2): run()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(5): run()
e1.record(); torch.cuda.synchronize()
print(round(e0.elapsed_time(e1) / 5, 4), "ms")
