#!/bin/bash
# refinement kernel: parity tests, timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_refinement.py tests/test_merge_maps.py -q -m gpu -x 2>&1 | tail -25
timeout 600 python tests/gpu_tools/refine_time.py 10000 60 1001 2>&1 | tail -3 | tee gpurun_out/refine_time.json
