#!/bin/bash
# ncu evidence for the section-8f kernels (current code) + launch list of a short bench run
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-generated --e2e-steps 1"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r1_launches_bench.csv $B > gpurun_out/ncu_launch.log 2>&1
echo "launch list exit $?"
cap() { timeout 600 ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c 1 -o gpurun_out/prof_$1 -f ${@:3} > gpurun_out/ncu_$1.log 2>&1; echo "ncu $1 exit $?"; }
cap kdi_refine_kernel 1 python tests/gpu_tools/refine_time.py 3000 60 1001
cap kdi_merge_maps_kernel 1 python tests/gpu_tools/merge_time.py 500 20 2
cap kdi_preprocess_kernel 1 python tests/gpu_tools/preprocess_time.py 100 60
cap kdi_average_neighbours_kernel 1 python tests/gpu_tools/preprocess_time.py 100 60
ls -la gpurun_out/*.ncu-rep
