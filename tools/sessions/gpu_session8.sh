#!/bin/bash
# 8-GPU session: BASELINE configs 4 and 5 at full size (verified), bench at N = 8 and 4
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node 8 --master-port 29541 tests/gpu_tools/run_config.py --config 4 > gpurun_out/config4.json 2> gpurun_out/config4.err
echo "config 4 exit $?"; grep '^{' gpurun_out/config4.json; tail -3 gpurun_out/config4.err
timeout 900 $TR --nproc-per-node 8 --master-port 29542 tests/gpu_tools/run_config.py --config 5 --sample-oracle 0 > gpurun_out/config5.json 2> gpurun_out/config5.err
echo "config 5 exit $?"; grep '^{' gpurun_out/config5.json; tail -3 gpurun_out/config5.err
KDI_TRACE=1 timeout 600 $TR --nproc-per-node 8 --master-port 29543 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
echo "bench n8 exit $?"; grep 'kdi trace' gpurun_out/bench_n8.json | tail -3; grep '^{' gpurun_out/bench_n8.json
KDI_TRACE=1 timeout 600 $TR --nproc-per-node 4 --master-port 29544 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err
echo "bench n4 exit $?"; grep 'kdi trace' gpurun_out/bench_n4.json | tail -3; grep '^{' gpurun_out/bench_n4.json
