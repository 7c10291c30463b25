"""Debug aid: one draw of the stroke fuzz scene, sub-path by sub-path and with style variations, GPU vs oracle."""
import ctypes as C
import sys

import numpy as np

sys.path.insert(0, ".")
from tests.oracle_backend import load_oracle, render_scene  # noqa: E402
from tests.test_gpu_stroke_fuzz import SIZE, fuzz_scene  # noqa: E402
from z2d_b200 import abi, workloads  # noqa: E402
from z2d_b200.abi import AntiAliasMode, Format  # noqa: E402
from z2d_b200.cuda_backend import CudaBackend  # noqa: E402
from z2d_b200.host import Pixel, Surface  # noqa: E402

seed, di = int(sys.argv[1]), int(sys.argv[2])
cb = CudaBackend(0)
scene = fuzz_scene(seed, 300, AntiAliasMode.default)
orc = load_oracle(fast=True)
sfc = Surface(Format.rgba, SIZE, SIZE, None, cb)
nd = scene.nodes[scene.node_off[di]:scene.node_off[di + 1]].copy()
so = scene.stroke_opts[di:di + 1].copy()
print("nodes:")
for n in nd:
    print("  ", "MLCZ"[n["tag"]], [float(v) for v in n["p"]])
print("opts:", so)


def run(nodes, opts, label):
    sc = workloads.Scene(SIZE, SIZE, nodes, np.array([0, len(nodes)], dtype=np.int64), scene.patterns[di:di + 1].copy(), np.ones(1, np.uint32),
                         np.zeros(0, dtype=workloads.FILLOPTS_DT), opts, np.zeros(1, np.int64), keep=scene.keep)
    sfc.paint_pixel(Pixel.rgba(0, 0, 0, 0))
    cmds = sc.draw_cmds(sfc.handle)
    cb.submit(cmds.ctypes.data_as(C.POINTER(abi.DrawCmdPOD)), 1)
    got = sfc.download()
    ref = render_scene(orc, sc)
    bad = int((got.reshape(-1, 4) != ref.reshape(-1, 4)).any(axis=1).sum())
    print("   gpu edges", cb.stats()["edges"], end="")
    print(f"{label}: {bad} px differ (gpu covered {int((got.reshape(-1, 4)[:, 3] > 0).sum())}, oracle {int((ref.reshape(-1, 4)[:, 3] > 0).sum())})")


run(nd, so, "whole draw")
starts = [i for i, n in enumerate(nd) if n["tag"] == 0] + [len(nd)]
for a, b in zip(starts[:-1], starts[1:]):
    run(nd[a:b].copy(), so, f"sub-path nodes [{a},{b}) {''.join('MLCZ'[t] for t in nd['tag'][a:b])}")
for a in range(len(starts) - 1):
    for b in range(a + 2, len(starts)):
        run(nd[starts[a]:starts[b]].copy(), so, f"sub-paths {a}..{b - 1}")
for name, field, val in [("offset 0", "dash_offset", 0.0), ("identity ctm", "ctm", [1, 0, 0, 1, 0, 0]), ("2 dashes", "n_dashes", 2), ("no dashes", "n_dashes", 0),
                         ("join miter", "line_join_mode", 0)]:
    o = so.copy()
    o[field] = val
    run(nd, o, name)
