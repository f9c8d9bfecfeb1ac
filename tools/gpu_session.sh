#!/bin/bash
# Session 67 (2 GPUs): the sharded test (peer and collective exchange, certificate modes) and smoke()'s 2-rank check.
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q > gpurun_out/s67_sharded.log 2>&1
echo "sharded exit $?"; tail -5 gpurun_out/s67_sharded.log | cut -c1-300
