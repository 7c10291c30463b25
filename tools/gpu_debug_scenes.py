#!/usr/bin/env python3
"""Run every ported scene on the GPU and print per-scene mismatch statistics vs the oracle."""
import sys, os, time, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import specs, golden_util
from tests.oracle_backend import OracleBackend
from z2d_b200.cuda_backend import CudaBackend

ob, cb = OracleBackend(), CudaBackend()
zo, zc = specs.bind(ob), specs.bind(cb)
only = sys.argv[1:]
tot_bad = 0
def cmp(name, fn_c, fn_o):
    global tot_bad
    try:
        t = time.time(); got = fn_c(); tg = time.time() - t
        ref = fn_o()
        g, r = got.pixels().astype(np.int32), ref.pixels().astype(np.int32)
        d = np.abs(g - r)
        bad = int((d != 0).any(axis=-1).sum())
        ys, xs = np.nonzero((d != 0).any(axis=-1))
        where = f" first=({xs[0]},{ys[0]}) got={g[ys[0],xs[0]].tolist()} ref={r[ys[0],xs[0]].tolist()} bbox=x[{xs.min()},{xs.max()}] y[{ys.min()},{ys.max()}]" if bad else ""
        print(f"{name}: bad={bad} maxdiff={d.max()} gpu_s={tg:.3f}{where}", flush=True)
        tot_bad += bad != 0
    except Exception as e:
        print(f"{name}: EXC {type(e).__name__}: {e}", flush=True)
        tot_bad += 1
for stem, fn in sorted(specs.PATH_SCENES.items()):
    if only and not any(stem.startswith(o) for o in only): continue
    for aa, suf in golden_util.AA_SUFFIX:
        cmp(stem + suf, lambda: fn(zc, aa), lambda: fn(zo, aa))
for stem, fn in sorted(specs.COMPOSITOR_SCENES.items()):
    if only and not any(stem.startswith(o) for o in only): continue
    cmp(stem, lambda: fn(zc), lambda: fn(zo))
print("scenes with differences:", tot_bad)
