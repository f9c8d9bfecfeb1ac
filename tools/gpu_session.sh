#!/bin/bash
# Session 40: new projection tests (own arithmetic vs math library, schedules of a generated dictionary).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_projection.py -m gpu -q > gpurun_out/s40_pytest.log 2>&1
echo "pytest exit $?"; tail -30 gpurun_out/s40_pytest.log
