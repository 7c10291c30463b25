# Round profile recipe (run on the GPU box through gpurun): the ncu launch list of the bench command and one full capture
# of the raster, flatten and compositor kernels, exported to CSV (the .ncu-rep files are too large to bring back).
set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --chunk 0 > gpurun_out/ncu_launches.log 2>&1
for spec in "k_raster_tiles:raster:2" "k_flatten_nodes:flatten:9" "k_composite_fast:composite:1"; do
  K=${spec%%:*}; rest=${spec#*:}; TAG=${rest%%:*}; SKIP=${rest#*:}
  ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c 1 -f -o /tmp/prof_$TAG python bench.py --steps 2 --warmup 1 --no-cpu-baseline --chunk 0 > gpurun_out/ncu_$TAG.log 2>&1
  ncu -i /tmp/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv
  ncu -i /tmp/prof_$TAG.ncu-rep --page details > gpurun_out/${TAG}_details.txt
done
ncu -i /tmp/prof_raster.ncu-rep --page source --csv > gpurun_out/raster_source.csv
ls -la gpurun_out/
