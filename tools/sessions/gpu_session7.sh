#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?")
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python tests/gpu_tools/run_config.py --config 3 > gpurun_out/config3.json 2> gpurun_out/config3.err
echo "config 3 exit $?"; cat gpurun_out/config3.json; tail -3 gpurun_out/config3.err
OVERLAP=0 KDI_TIMELINE=1 M=40000 timeout 300 python tools/timeline.py 2>&1 | tail -12
