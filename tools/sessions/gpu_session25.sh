#!/bin/bash
# ncu evidence for the refinement and merge-maps kernels + merge timing
mkdir -p gpurun_out
timeout 300 python tests/gpu_tools/merge_time.py 1000 20 2 2>&1 | tail -2 | tee gpurun_out/merge_time.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kdi_refine_kernel -s 1 -c 1 -o gpurun_out/prof_kdi_refine_kernel -f python tests/gpu_tools/refine_time.py 3000 60 1001 > gpurun_out/ncu_refine.log 2>&1; echo "ncu refine exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kdi_merge_maps_kernel -s 1 -c 1 -o gpurun_out/prof_kdi_merge_maps_kernel -f python tests/gpu_tools/merge_time.py 500 20 2 > gpurun_out/ncu_merge.log 2>&1; echo "ncu merge exit $?"
ls -la gpurun_out/*.ncu-rep
