#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?")
tail -6 gpurun_out/pytest_gpu.log
for cfg in "KEEP=20 M=80000 N=12500 OVERLAP=0" "KEEP=20 M=10000 N=100000"; do
  env $cfg KDI_TIMELINE=1 timeout 300 python tools/timeline.py 2>&1 | awk '/====/{p=1} p' 
done
timeout 900 python tests/gpu_tools/run_config.py --config 3 --sample-oracle 0 > gpurun_out/config3.json 2> gpurun_out/config3.err
echo "config 3 exit $?"; cat gpurun_out/config3.json; tail -3 gpurun_out/config3.err
