"""Where does config 5's device time go?  Renders N scenes with / without strokes and gradient fills (device-resident replay)."""
import ctypes as C
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from z2d_b200 import abi, workloads  # noqa: E402
from z2d_b200.abi import Format  # noqa: E402
from z2d_b200.cuda_backend import CudaBackend  # noqa: E402
from z2d_b200.host import Surface  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
cb = CudaBackend(0, stream=stream.cuda_stream)
cb.set_chunk(0)
for label, kw in (("fills only", dict(n_strokes=0, n_gradients=0)), ("fills+strokes", dict(n_gradients=0)), ("fills+gradients", dict(n_strokes=0)),
                  ("all", dict())):
    scenes = [workloads.mixed_scene(s, 1024, **kw) for s in range(n)]
    sfcs = [Surface(Format.rgba, 1024, 1024, None, cb) for _ in range(n)]
    cmds = np.concatenate([sc.draw_cmds(sf.handle) for sc, sf in zip(scenes, sfcs)])
    cb.submit(cmds.ctypes.data_as(C.POINTER(abi.DrawCmdPOD)), len(cmds))
    cb.sync()
    for _ in range(3):
        cb.replay()
    st = cb.stats()
    print(label, len(cmds), {k: round(st[k], 3) for k in ("ms_flatten", "ms_bin", "ms_lists", "ms_raster", "ms_total")},
          {k: st[k] for k in ("edges", "band_edges", "tile_pairs", "covered_px", "crossings")}, flush=True)
    for sf in sfcs:
        sf.deinit()
