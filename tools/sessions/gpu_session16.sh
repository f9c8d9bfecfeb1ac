#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest exit $?")
tail -6 gpurun_out/pytest_gpu2.log
KDI_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "bench n2 exit $?"; grep "kdi trace" gpurun_out/bench_n2.json | tail -4; grep '^{' gpurun_out/bench_n2.json; tail -3 gpurun_out/bench_n2.err
