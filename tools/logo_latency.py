"""BASELINE config 1 (spec/080_fill_z2d_logo): latency-bound small scene, timed as back-to-back renders.

Glyph outlines -> nodes is host-side text layout outside the hot path, so the five node lists of the scene are built once;
each render then is: clear the 601x172 RGBA surface, five painter.fill calls, wait for completion.  Prints one JSON line with
microseconds per scene on the device and on one host core running the CPU oracle."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import specs  # noqa: E402
from tests.oracle_backend import OracleBackend  # noqa: E402
from z2d_b200 import host  # noqa: E402
from z2d_b200.abi import AntiAliasMode  # noqa: E402


class Recorder:
    """Backend wrapper that records the (pattern, nodes, opts) of every fill so the scene can be replayed without text layout."""

    def __init__(self, inner):
        self.inner, self.calls = inner, []

    def __getattr__(self, name):
        return getattr(self.inner, name)

    def fill(self, hd, pat, nodes, n, opts):
        self.calls.append((pat, nodes, n, opts))
        return self.inner.fill(hd, pat, nodes, n, opts)


def bench(backend, reps):
    rec = Recorder(backend)
    sfc = specs.PATH_SCENES["080_fill_z2d_logo"](specs.bind(rec), AntiAliasMode.default)
    zero = host.Pixel.rgba(0, 0, 0, 0)
    for _ in range(3):
        sfc.paint_pixel(zero)
        for pat, nodes, n, opts in rec.calls:
            backend.fill(sfc.handle, pat, nodes, n, opts)
        backend.sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        sfc.paint_pixel(zero)
        for pat, nodes, n, opts in rec.calls:
            backend.fill(sfc.handle, pat, nodes, n, opts)
        backend.sync()
    dt = time.perf_counter() - t0
    return dt / reps * 1e6, sum(c[2] for c in rec.calls), len(rec.calls)


if __name__ == "__main__":
    from z2d_b200.cuda_backend import CudaBackend
    gpu_us, nodes, calls = bench(CudaBackend(0), 1000)
    cpu_us, _, _ = bench(OracleBackend(fast=True), 50)
    print(json.dumps({"workload": "spec/080_fill_z2d_logo (BASELINE config 1): 601x172 RGBA8, 5 fills, default AA, node lists prebuilt",
                      "fills_per_scene": calls, "nodes_per_scene": nodes, "gpu_us_per_scene": gpu_us, "cpu_oracle_us_per_scene": cpu_us,
                      "note": "one batch per scene: fixed cost of a batch (uploads, ~27 launches, 2 count read-backs) dominates"}))
