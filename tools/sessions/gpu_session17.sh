#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?")
tail -25 gpurun_out/pytest_gpu.log
timeout 300 python tools/project_time.py 2>&1 | tail -4
