# Round profile recipe (run on the GPU box through gpurun): bench lines, ncu launch list, one full capture of
# the raster kernel and of the compositor kernel, exported to CSV (the .ncu-rep files are too large to bring back).
set -x
python bench.py > gpurun_out/bench8.json 2> gpurun_out/bench8.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench8_ref.json 2> gpurun_out/bench8_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches8.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_l8.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_raster_tiles -s 2 -c 1 -f -o /tmp/raster python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_r8.log 2>&1
ncu -i /tmp/raster.ncu-rep --page raw --csv > gpurun_out/raster8_raw.csv
ncu -i /tmp/raster.ncu-rep --page source --csv > gpurun_out/raster8_source.csv
ncu -i /tmp/raster.ncu-rep --page details > gpurun_out/raster8_details.txt
ncu --set full --clock-control none -k regex:k_composite -s 2 -c 1 -f -o /tmp/composite python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_c8.log 2>&1
ncu -i /tmp/composite.ncu-rep --page raw --csv > gpurun_out/composite8_raw.csv
ncu -i /tmp/composite.ncu-rep --page details > gpurun_out/composite8_details.txt
ls -la gpurun_out/
