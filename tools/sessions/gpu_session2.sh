#!/bin/bash
# GPU session 2: tests, L2 sweep (timing + DRAM bytes), bench
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?")
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python tools/gemm_sweep.py > gpurun_out/sweep_time.jsonl 2> gpurun_out/sweep_time.err
echo "sweep exit $?"; cat gpurun_out/sweep_time.jsonl
REPS=2 timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,sm__cycles_elapsed.avg.per_second --clock-control none -k regex:kdi_gemm_kernel --csv --log-file gpurun_out/sweep_ncu.csv python tools/gemm_sweep.py > gpurun_out/sweep_ncu.log 2>&1
echo "ncu sweep exit $?"
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench exit $?"; cat gpurun_out/bench_n1.json
