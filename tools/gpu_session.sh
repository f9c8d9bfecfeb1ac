#!/bin/bash
# Session 45: index scratch handed out by ticket - parity suite, keep_n = 100 timing.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/s45_pytest.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/s45_pytest.log
env M=20000 N=100000 KEEP=100 ROUNDS=3 REPS=2 SETTINGS="split=1" timeout 900 python tools/schedule_sweep.py 2>&1 | tail -1
env M=100000 N=37500 KEEP=50 ROUNDS=3 REPS=2 SETTINGS="split=1" timeout 900 python tools/schedule_sweep.py 2>&1 | tail -1
