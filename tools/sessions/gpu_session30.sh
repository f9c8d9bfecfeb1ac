#!/bin/bash
# merge-maps kernel with the register sort, projection with one PC per rotation
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_merge_maps.py tests/test_gpu_projection.py -q -m gpu -x 2>&1 | tail -8
timeout 300 python tests/gpu_tools/merge_time.py 1000 20 2 2>&1 | tail -1 | tee gpurun_out/merge_time.json
timeout 300 python tests/gpu_tools/merge_time.py 500 50 3 2>&1 | tail -1
