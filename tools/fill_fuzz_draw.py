"""Debug aid: one draw of the fill fuzz scene, sub-path by sub-path and pair by pair, GPU vs oracle."""
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

seed, aa, di = int(sys.argv[1]), AntiAliasMode[sys.argv[2]], int(sys.argv[3])
cb = CudaBackend(0)
scene = fuzz_scene(seed, 300, aa)
orc = load_oracle(fast=True)
sfc = Surface(Format.rgba, SIZE, SIZE, None, cb)
nd = scene.nodes[scene.node_off[di]:scene.node_off[di + 1]].copy()
fo = scene.fill_opts[di:di + 1].copy()
fo["op"] = int(abi.Operator.src_over)
for n in nd:
    print("  ", "MLCZ"[n["tag"]], [float(v) for v in n["p"]])
print(fo)


def run(nodes, label):
    sc = workloads.Scene(SIZE, SIZE, nodes, np.array([0, len(nodes)], dtype=np.int64), scene.patterns[di:di + 1].copy(), np.zeros(1, np.uint32),
                         fo, np.zeros(0, dtype=workloads.STROKEOPTS_DT), np.zeros(1, np.int64))
    sfc.paint_pixel(Pixel.rgba(0, 0, 0, 0))
    cmds = sc.draw_cmds(sfc.handle)
    cb.submit(cmds.ctypes.data_as(C.POINTER(abi.DrawCmdPOD)), 1)
    got = sfc.download().reshape(SIZE, SIZE, 4)
    ref = render_scene(orc, sc).reshape(SIZE, SIZE, 4)
    d = (got != ref).any(axis=-1)
    rows = np.nonzero(d.any(axis=1))[0]
    print(f"{label}: {int(d.sum())} px differ, rows {rows[:6].tolist()}{'...' if len(rows) > 6 else ''}", end="")
    if len(rows):
        r = rows[0]
        xs = np.nonzero(d[r])[0]
        print(f"  row {r}: x {xs[0]}..{xs[-1]} gpu a={got[r, xs[0], 3]} oracle a={ref[r, xs[0], 3]}", end="")
    print()


run(nd, "whole")
starts = [i for i, n in enumerate(nd) if n["tag"] == 0] + [len(nd)]
for a in range(len(starts) - 1):
    for b in range(a + 1, len(starts)):
        run(nd[starts[a]:starts[b]].copy(), f"sub-paths {a}..{b - 1} {''.join('MLCZ'[t] for t in nd['tag'][starts[a]:starts[b]])}")
