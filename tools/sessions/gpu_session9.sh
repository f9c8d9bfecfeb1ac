#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?")
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python tools/project_time.py 2>&1 | tail -6
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench exit $?"; cat gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
