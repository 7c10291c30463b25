"""BASELINE config 3 shape (2048^2, 50k stroked polylines / Beziers with joins, caps and dashes): device stage timings."""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from z2d_b200 import abi, workloads  # noqa: E402
from z2d_b200.cuda_backend import CudaBackend  # noqa: E402
from z2d_b200.host import Pixel, Surface  # noqa: E402

stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
cb = CudaBackend(0, stream=stream.cuda_stream)
cb.set_chunk(0)
scene = workloads.stroke_paths_scene(50_000, 2048)
sfc = Surface(abi.Format.rgba, 2048, 2048, None, cb)
cmds = scene.draw_cmds(sfc.handle)
cb.submit(cmds.ctypes.data_as(C.POINTER(abi.DrawCmdPOD)), scene.n)
cb.sync()
zero = Pixel.rgba(0, 0, 0, 0)
best = None
for it in range(8):
    sfc.paint_pixel(zero)
    cb.replay()
    st = cb.stats()
    if it >= 2 and (best is None or st["ms_total"] < best["ms_total"]):
        best = st
print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in best.items() if k in ("draws", "edges", "band_edges", "tile_pairs", "crossings", "covered_px", "ms_flatten", "ms_bin", "ms_lists", "ms_raster", "ms_total")})
print(best["ms_raster"], best["ms_flatten"], f"{best['draws'] / best['ms_total'] / 1e3:.2f} M strokes/s, {best['covered_px'] / best['ms_total'] / 1e6:.2f} Gpix/s")
