#!/bin/bash
# Session 75: compute-sanitizer memcheck over the certificate tests at full dictionary size (bound-first, strict, shard stages).
mkdir -p gpurun_out
timeout 55 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "strict_certificate_is_a_bound and ncc" > gpurun_out/s75_memcheck.log 2>&1
echo "memcheck exit $?"; grep -a "ERROR SUMMARY\|passed\|failed" gpurun_out/s75_memcheck.log | tail -3
