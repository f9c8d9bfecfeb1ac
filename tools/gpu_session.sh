#!/bin/bash
# Session 57: the 512 x 256 tile at BASELINE config 2 under sustained load (long interleaved runs).
mkdir -p gpurun_out
env ROUNDS=5 REPS=20 SETTINGS="dual=0;dual=2" timeout 600 python tools/schedule_sweep.py > gpurun_out/s57_sweep_c2.jsonl 2> gpurun_out/s57_sweep_c2.err
cat gpurun_out/s57_sweep_c2.jsonl | cut -c1-260; tail -3 gpurun_out/s57_sweep_c2.err
env M=20000 ROUNDS=4 REPS=12 SETTINGS="dual=0;dual=2" timeout 600 python tools/schedule_sweep.py > gpurun_out/s57_sweep_c2_m20k.jsonl 2> gpurun_out/s57_sweep_c2_m20k.err
cat gpurun_out/s57_sweep_c2_m20k.jsonl | cut -c1-260
