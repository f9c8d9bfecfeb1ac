#!/bin/bash
# Session 42: kc >= 64 candidate lists with the index halves in an L2-resident scratch (one more stage).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/s42_pytest.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/s42_pytest.log
env M=100000 N=37500 KEEP=50 ROUNDS=4 REPS=2 SETTINGS="split=1;stages=5;stages=6" timeout 900 python tools/schedule_sweep.py > gpurun_out/s42_sweep_c4.jsonl 2> gpurun_out/s42_sweep_c4.err
cat gpurun_out/s42_sweep_c4.jsonl; tail -3 gpurun_out/s42_sweep_c4.err
env M=20000 N=100000 KEEP=100 ROUNDS=3 REPS=2 SETTINGS="split=1;stages=3;stages=4" timeout 900 python tools/schedule_sweep.py > gpurun_out/s42_sweep_k100.jsonl 2> gpurun_out/s42_sweep_k100.err
cat gpurun_out/s42_sweep_k100.jsonl; tail -3 gpurun_out/s42_sweep_k100.err
