#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tests/gpu_tools/realistic_check.py 2>&1 | tee gpurun_out/realistic.jsonl | tail -8
