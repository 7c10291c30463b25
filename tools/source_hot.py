"""Per-source-line share of executed warp instructions from `ncu --page source --csv --print-source cuda,sass` (tools/run_profile_raster.sh)."""
import collections
import csv
import sys

path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/raster_source_cuda.csv"
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 50
rows = list(csv.reader(open(path)))
cur, agg, tot, thr = None, {}, 0, {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        ie, it = r.index("Instructions Executed"), r.index("Thread Instructions Executed")
        continue
    try:
        ln, n, t = int(r[0]), int(r[ie]), int(r[it])
    except (ValueError, IndexError):
        continue
    if r[2] == "-":
        agg[(cur, ln)] = (n, t, r[1].strip())
        tot += n
print("total warp instructions", tot)
for (f, ln), (n, t, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top_n]:
    print(f"{f}:{ln:4d} {n / tot * 100:5.2f}% lanes {t / max(n, 1):4.1f}  {src[:120]}")
