#!/bin/bash
# GPU session 3 (2 GPUs): sharded test + bench n=2, interleaved sweep of strip size / stages
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest exit $?")
tail -5 gpurun_out/pytest_gpu2.log
ROUNDS=4 REPS=3 SETTINGS="0,0,0,0;0,4,0,0;0,6,0,0;0,3,0,0;0,0,0,0,5;0,4,0,0,5;0,6,0,0,4;0,12,0,0;14,6,0,0" timeout 600 python tools/gemm_sweep.py > gpurun_out/sweep2_time.jsonl 2> gpurun_out/sweep2_time.err
echo "sweep exit $?"; cat gpurun_out/sweep2_time.jsonl; tail -3 gpurun_out/sweep2_time.err
KDI_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "bench n2 exit $?"; tail -4 gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err
