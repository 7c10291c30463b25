"""Turn the CSV exports of tools/run_profile.sh (gpurun_out/) into the small, tracked summaries under profiles/.

usage: python tools/summarise_profile.py <round tag, e.g. r01b>
  profiles/<tag>_launches.json      per-kernel share of the bench step from the ncu launch list (gpu__time_duration.sum)
  profiles/<tag>_<kernel>_ncu.json  key metrics of the --set full capture (duration, DRAM bytes, issue/stall/occupancy)
  profiles/<tag>_launches.csv       the launch list itself
"""
import collections
import csv
import json
import shutil
import sys

tag = sys.argv[1]
G = "gpurun_out/"

rows = list(csv.reader(l for l in open(G + "launches.csv") if l.startswith('"')))
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    if len(r) <= vi:
        continue
    name = r[ki].split("(")[0].replace("z2d::", "").replace("void ", "")
    a = agg.setdefault(name, [0, 0.0, 0.0])
    v = float(r[vi].replace(",", ""))
    a[0] += 1
    a[1] += v
    a[2] = max(a[2], v)
total = sum(a[1] for a in agg.values())
out = {"command": "ncu --metrics gpu__time_duration.sum --clock-control none -c 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-c5 --chunk 0",
       "note": "per-launch times under ncu are serialised and cold-cache; the SHARE of each kernel is what carries over to the bench",
       "total_ms": total / 1e6,
       "kernels": [{"kernel": k, "launches": a[0], "total_ms": a[1] / 1e6, "max_ms": a[2] / 1e6, "share": a[1] / total}
                   for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])]}
json.dump(out, open(f"profiles/{tag}_launches.json", "w"), indent=1)
shutil.copy(G + "launches.csv", f"profiles/{tag}_launches.csv")

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__grid_size", "launch__block_size",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active", "smsp__sass_inst_executed_op_local_ld.sum",
        "smsp__sass_inst_executed_op_local_st.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed"]
for kern in ("raster", "flatten", "composite", "raster_strokes", "flatten_strokes", "stroke_walk", "stroke_units", "composite_gen", "composite_lut"):
    try:
        rows = list(csv.reader(open(G + f"{kern}_raw.csv")))
    except FileNotFoundError:
        continue
    hdr, units, val = rows[0], rows[1], rows[2]
    d = {"kernel": val[hdr.index("Kernel Name")] if "Kernel Name" in hdr else kern, "metrics": {}, "stalls_per_issue": {}}
    for i, h in enumerate(hdr):
        if h in KEYS:
            d["metrics"][h] = {"value": val[i], "unit": units[i]}
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            try:
                v = float(val[i])
            except ValueError:
                continue
            if v >= 0.05:
                d["stalls_per_issue"][h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = round(v, 3)
    m = d["metrics"]
    try:
        def num(k):
            v, u = m[k]["value"].replace(",", ""), m[k]["unit"]
            f = float(v)
            return f * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1}.get(u, 1)
        rd, wr, t = num("dram__bytes_read.sum"), num("dram__bytes_write.sum"), num("gpu__time_duration.sum")
        d["derived"] = {"dram_bytes": rd + wr, "duration_ms": t * 1e3, "dram_GBps": (rd + wr) / t / 1e9}
    except Exception as e:  # noqa: BLE001
        d["derived"] = {"error": str(e)}
    json.dump(d, open(f"profiles/{tag}_{kern}_ncu.json", "w"), indent=1)
    print(kern, d["derived"], d["stalls_per_issue"])
print(json.dumps(out["kernels"][:8], indent=0)[:900])
