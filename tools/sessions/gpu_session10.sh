#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_projection.py -q -m gpu > gpurun_out/pytest_proj.log 2>&1; echo "pytest exit $?")
tail -8 gpurun_out/pytest_proj.log
timeout 300 python tools/project_time.py 2>&1 | tail -4
