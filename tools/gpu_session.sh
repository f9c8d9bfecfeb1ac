#!/bin/bash
# Session 43: how much of the kc = 64 / kc = 32 GEMM time is the candidate insertion (bare GEMM vs full).
mkdir -p gpurun_out
for ni in 0 1 0 1; do
echo "C4 shard shape (kc 64), KDI_GEMM_NO_INSERT=$ni"
env KDI_GEMM_NO_INSERT=$ni M=100000 N=37500 KEEP=50 ROUNDS=2 REPS=2 SETTINGS="split=1" timeout 600 python tools/schedule_sweep.py 2>&1 | tail -1
done
for ni in 0 1 0 1; do
echo "C2 shape (kc 32), KDI_GEMM_NO_INSERT=$ni"
env KDI_GEMM_NO_INSERT=$ni ROUNDS=3 REPS=3 SETTINGS="split=1" timeout 600 python tools/schedule_sweep.py 2>&1 | tail -1
done
