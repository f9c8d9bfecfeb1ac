#!/bin/bash
# refinement kernel: occupancy variants
for mb in 2 3 4; do echo "MINB $mb"; KDI_REFINE_MINB=$mb timeout 300 python tests/gpu_tools/refine_time.py 10000 60 1001 2>&1 | tail -1; done
