#!/bin/bash
mkdir -p gpurun_out
for cfg in "OVERLAP=1" "OVERLAP=1 KDI_CARVEOUT=-1" "OVERLAP=1 KDI_GEMM_CARVEOUT=-1 KDI_CARVEOUT=-1" "OVERLAP=1 MAX_STAGES=4"; do
  env $cfg KDI_TIMELINE=1 timeout 300 python tools/timeline.py 2>&1 | awk '/====/{p=1} p' 
done > gpurun_out/timeline.log 2>&1
cat gpurun_out/timeline.log
