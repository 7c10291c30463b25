# usage: bash tools/run_profile_kernel.sh <kernel regex> <tag>   -- one full ncu capture, exported to CSV under gpurun_out/
set -x
K=$1; TAG=$2
ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -f -o /tmp/prof_$TAG python tools/warmup_probe.py 3 > gpurun_out/ncu_$TAG.log 2>&1
ncu -i /tmp/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv
ncu -i /tmp/prof_$TAG.ncu-rep --page source --csv > gpurun_out/${TAG}_source.csv
ncu -i /tmp/prof_$TAG.ncu-rep --page details > gpurun_out/${TAG}_details.txt
