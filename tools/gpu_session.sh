#!/bin/bash
# The command set of the current GPU session (rewritten per session; results land in gpurun_out/).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/s1_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/s1_pytest.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/s1_pytest.log
timeout 600 python tools/schedule_sweep.py > gpurun_out/s1_sweep.jsonl 2> gpurun_out/s1_sweep.err
echo "sweep exit $?"; cat gpurun_out/s1_sweep.jsonl; tail -3 gpurun_out/s1_sweep.err
KDI_TIMELINE=1 ROUNDS=1 REPS=3 SETTINGS="flags=1;flags=1,groups=4,post=1,sms=140,serial=1;flags=1,groups=4,post=1,part=8" \
  timeout 300 python tools/schedule_sweep.py > gpurun_out/s1_timeline.out 2> gpurun_out/s1_timeline.txt
echo "timeline exit $?"
timeout 600 python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/s1_bench.json 2> gpurun_out/s1_bench.err
echo "bench exit $?"; cat gpurun_out/s1_bench.json | head -c 3000
