#!/bin/bash
# The command set of the current GPU session (rewritten per session; results land in gpurun_out/).
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/s6_pytest.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/s6_pytest.log
timeout 600 python tools/schedule_sweep.py > gpurun_out/s6_sweep.jsonl 2> gpurun_out/s6_sweep.err
echo "sweep exit $?"; cat gpurun_out/s6_sweep.jsonl; tail -3 gpurun_out/s6_sweep.err
KDI_TIMELINE=1 ROUNDS=1 REPS=3 SETTINGS="flags=1;flags=1,groups=4,post=1,cores=1;flags=1,groups=4,post=1,cores=2" \
  timeout 300 python tools/schedule_sweep.py > gpurun_out/s6_timeline.out 2> gpurun_out/s6_timeline.txt
echo "timeline exit $?"
M=100000 N=37500 KEEP=50 ROUNDS=2 REPS=2 SETTINGS="flags=0;flags=1;flags=1,groups=16,post=1,cores=1" timeout 300 python tools/schedule_sweep.py > gpurun_out/s6_sweep_c4.jsonl 2> gpurun_out/s6_sweep_c4.err
echo "c4 sweep exit $?"; cat gpurun_out/s6_sweep_c4.jsonl
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/s6_bench.json 2> gpurun_out/s6_bench.err
echo "bench exit $?"; head -c 7000 gpurun_out/s6_bench.json; tail -5 gpurun_out/s6_bench.err
