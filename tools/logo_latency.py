"""BASELINE config 1 (spec/080_fill_z2d_logo) latency with the device stage times of one scene (see bench.py --workload c1)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_extra  # noqa: E402
from z2d_b200.cuda_backend import CudaBackend  # noqa: E402

cb = CudaBackend(0)
out = bench_extra.run_c1(cb, reps=1000, with_cpu="--cpu" in sys.argv)
st = cb.stats()
out["device_stage_ms_last_scene"] = {k: st[k] for k in ("ms_flatten", "ms_bin", "ms_lists", "ms_raster", "ms_total")}
out["counters"] = {k: st[k] for k in ("draws", "nodes", "edges", "band_edges", "tile_items", "tile_pairs", "covered_px")}
print(json.dumps(out))
