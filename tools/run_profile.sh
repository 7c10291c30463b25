# Round profile recipe (run on the GPU box through gpurun): the ncu launch list of the bench command and one full capture of
# the dominant kernels, exported to CSV (the .ncu-rep files are too large to bring back).  tools/summarise_profile.py <tag>
# turns the CSVs into profiles/<tag>_*.json.
set -x
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-c5 --no-extras --chunk 0"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_launches.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_c3.csv $B --workload c3 > gpurun_out/ncu_launches_c3.log 2>&1
cap() {  # kernel regex, tag, launches to skip, command...
  K=$1; TAG=$2; SKIP=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c 1 -f -o /tmp/prof_$TAG "$@" > gpurun_out/ncu_$TAG.log 2>&1
  ncu -i /tmp/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv
  ncu -i /tmp/prof_$TAG.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/${TAG}_source_cuda.csv 2>/dev/null
}
cap '^k_raster_tiles$' raster 2 $B
cap 'k_flatten_nodes' flatten 9 $B
cap 'k_composite_fast' composite 1 $B
cap 'k_raster_tiles_rich' raster_strokes 2 $B --workload c3
cap 'k_stroke_walk' stroke_walk 2 $B --workload c3
cap 'k_stroke_units' stroke_units 2 $B --workload c3
cap 'k_composite_gen' composite_gen 2 python tools/c4_one.py rgba linear none src_over integer
cap 'k_composite_lut' composite_lut 2 python tools/c4_one.py alpha8 pixel none src_over integer
ls -la gpurun_out/ | tail -30
