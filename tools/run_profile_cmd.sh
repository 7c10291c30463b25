# usage: bash tools/run_profile_cmd.sh <kernel regex> <tag> <python script> [args]  -- one full ncu capture of a kernel of any
# driver script, exported to CSV under gpurun_out/ (the .ncu-rep stays in /tmp on the box: gpurun_out is capped at 64 MiB)
set -x
K=$1; TAG=$2; shift 2
ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -f -o /tmp/prof_$TAG python "$@" > gpurun_out/ncu_$TAG.log 2>&1
ncu -i /tmp/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv
ncu -i /tmp/prof_$TAG.ncu-rep --page source --csv > gpurun_out/${TAG}_source.csv
ncu -i /tmp/prof_$TAG.ncu-rep --page details > gpurun_out/${TAG}_details.txt
