#!/bin/bash
# preprocessing kernels: parity tests + timing; refinement tests again with the final register allocation
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_preprocess.py tests/test_refinement.py -q -m gpu -x 2>&1 | tail -25
timeout 600 python tests/gpu_tools/preprocess_time.py 200 60 2>&1 | tail -2 | tee gpurun_out/preprocess_time.json
