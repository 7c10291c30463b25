"""Debug aid: render every draw of tests/test_gpu_stroke_fuzz.py's scene on its own surface, GPU vs oracle; print the style of those that differ."""
import ctypes as C
import sys

import numpy as np

sys.path.insert(0, ".")
from tests.oracle_backend import load_oracle, render_scene  # noqa: E402
from tests.test_gpu_stroke_fuzz import SIZE, fuzz_scene  # noqa: E402
from z2d_b200 import abi  # noqa: E402
from z2d_b200.abi import AntiAliasMode, Format  # noqa: E402
from z2d_b200.cuda_backend import CudaBackend  # noqa: E402
from z2d_b200.host import Pixel, Surface  # noqa: E402

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 11
cb = CudaBackend(0)
scene = fuzz_scene(seed, 300, AntiAliasMode.default)
orc = load_oracle(fast=True)
sfc = Surface(Format.rgba, SIZE, SIZE, None, cb)
nbad = 0
for i in range(scene.n):
    sfc.paint_pixel(Pixel.rgba(0, 0, 0, 0))
    cmds = scene.draw_cmds(sfc.handle, i, i + 1)
    cb.submit(cmds.ctypes.data_as(C.POINTER(abi.DrawCmdPOD)), 1)
    got = sfc.download()
    ref = render_scene(orc, scene, i, i + 1)
    bad = int((got.reshape(-1, 4) != ref.reshape(-1, 4)).any(axis=1).sum())
    if bad:
        nbad += 1
        so = scene.stroke_opts[i]
        nd = scene.nodes[scene.node_off[i]:scene.node_off[i + 1]]
        if nbad <= 12:
            print(f"draw {i}: {bad} px  cap={so['line_cap_mode']} join={so['line_join_mode']} w={so['line_width']} ml={so['miter_limit']} tol={so['tolerance']} "
                  f"nd={so['n_dashes']} off={so['dash_offset']} ctm={[round(float(v), 3) for v in so['ctm'][:4]]} tags={''.join('MLCZ'[t] for t in nd['tag'])}")
            if so['n_dashes']:
                d = (C.c_double * int(so['n_dashes'])).from_address(int(so['dashes']))
                print("    dashes", list(d))
print("draws that differ:", nbad, "of", scene.n)
