# ncu --set full of one kernel with source correlation: tools/run_profile_kernel.sh <kernel regex> <tag> <command...>
k=$1; tag=$2; shift 2
ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o /tmp/$tag "$@" > gpurun_out/ncu_$tag.log 2>&1
ncu -i /tmp/$tag.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv
ncu -i /tmp/$tag.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/${tag}_source_cuda.csv 2>/dev/null
tail -3 gpurun_out/ncu_$tag.log
