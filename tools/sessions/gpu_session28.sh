#!/bin/bash
# preprocessing kernels after the sliding-window / vector IO rewrite
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_preprocess.py -q -m gpu -x 2>&1 | tail -15
timeout 600 python tests/gpu_tools/preprocess_time.py 200 60 2>&1 | tail -2 | tee gpurun_out/preprocess_time.json
