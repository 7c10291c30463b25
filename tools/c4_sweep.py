"""Run the config-4 compositor sweep and write gpurun_out/c4_cells.json + c4_table.md (under gpurun)."""
import json
import sys

import torch

sys.path.insert(0, ".")
import bench_extra  # noqa: E402
from bench import peaks  # noqa: E402
from z2d_b200.cuda_backend import CudaBackend  # noqa: E402

stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
cb = CudaBackend(0, stream=stream.cuda_stream)
peak, kind = peaks()
cells = bench_extra.run_c4(cb, peak, full="--quick" not in sys.argv)
json.dump({"peak_gbs": peak, "peak_kind": kind, "cells": cells, "summary": bench_extra.c4_summary(cells)}, open("gpurun_out/c4_cells.json", "w"), indent=0)
open("gpurun_out/c4_table.md", "w").write(bench_extra.c4_markdown(cells, peak))
for r in bench_extra.c4_summary(cells):
    print(r)
