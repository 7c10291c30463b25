"""Debug aid: every draw of tests/test_gpu_fill_fuzz.py's scene on its own surface, GPU vs oracle; prints those that differ."""
import ctypes as C
import sys

import numpy as np

sys.path.insert(0, ".")
from tests.oracle_backend import load_oracle, render_scene  # noqa: E402
from tests.test_gpu_fill_fuzz import SIZE, fuzz_scene  # noqa: E402
from z2d_b200 import abi, workloads  # noqa: E402
from z2d_b200.abi import AntiAliasMode, Format  # noqa: E402
from z2d_b200.cuda_backend import CudaBackend  # noqa: E402
from z2d_b200.host import Pixel, Surface  # noqa: E402

seed = int(sys.argv[1])
aa = AntiAliasMode[sys.argv[2]] if len(sys.argv) > 2 else AntiAliasMode.default
cb = CudaBackend(0)
scene = fuzz_scene(seed, 300, aa)
orc = load_oracle(fast=True)
sfc = Surface(Format.rgba, SIZE, SIZE, None, cb)
nbad = 0
for i in range(scene.n):
    for bg in (0, 1):  # empty and half-covered destination (operators that read dst)
        px = Pixel.rgba(0, 0, 0, 0) if bg == 0 else Pixel.rgba(40, 80, 20, 128)
        sfc.paint_pixel(px)
        cmds = scene.draw_cmds(sfc.handle, i, i + 1)
        cb.submit(cmds.ctypes.data_as(C.POINTER(abi.DrawCmdPOD)), 1)
        got = sfc.download()
        buf = np.zeros(SIZE * SIZE * 4, dtype=np.uint8)
        if bg:
            buf.reshape(-1, 4)[:] = (40, 80, 20, 128)
        # oracle on the same background
        P = C.POINTER
        pat = C.cast(C.c_void_p(int(cmds["pattern"][0])), P(abi.PatternPOD))
        nodes = C.cast(C.c_void_p(int(cmds["nodes"][0])), P(abi.Node))
        rc = orc.z2d_ref_fill(buf.ctypes.data_as(C.c_void_p), int(Format.rgba), SIZE, SIZE, pat, nodes, int(cmds["n_nodes"][0]),
                              C.cast(C.c_void_p(int(cmds["fill"][0])), P(abi.FillOptsPOD)))
        assert rc == 0
        bad = int((got.reshape(-1, 4) != buf.reshape(-1, 4)).any(axis=1).sum())
        if bad:
            nbad += 1
            fo = scene.fill_opts[i]
            nd = scene.nodes[scene.node_off[i]:scene.node_off[i + 1]]
            if nbad <= 10:
                print(f"draw {i} bg={bg}: {bad} px  rule={fo['fill_rule']} op={fo['op']} tol={fo['tolerance']} tags={''.join('MLCZ'[t] for t in nd['tag'])}")
print("draw/background pairs that differ:", nbad, "of", 2 * scene.n)
