# Full ncu capture of the raster kernel with CUDA-source correlation, exported to CSV (run under gpurun).
# usage: tools/run_profile_raster.sh [extra bench.py args]
ncu --set full --clock-control none --import-source on -k regex:k_raster_tiles -s 2 -c 1 -f -o /tmp/raster python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-composite --no-c5 --no-extras --chunk 0 "$@" > gpurun_out/ncu_r.log 2>&1
ncu -i /tmp/raster.ncu-rep --page raw --csv > gpurun_out/raster_raw.csv
ncu -i /tmp/raster.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/raster_source_cuda.csv 2> gpurun_out/raster_source_cuda.err
ncu -i /tmp/raster.ncu-rep --page source --csv > gpurun_out/raster_source.csv
ls -la gpurun_out/ | tail -8
