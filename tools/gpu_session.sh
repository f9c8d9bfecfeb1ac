#!/bin/bash
# The command set of the current GPU session (rewritten per session; results land in gpurun_out/).
mkdir -p gpurun_out
timeout 600 python tools/schedule_sweep.py > gpurun_out/s12_sweep.jsonl 2> gpurun_out/s12_sweep.err
echo "sweep exit $?"; cat gpurun_out/s12_sweep.jsonl; tail -3 gpurun_out/s12_sweep.err | cut -c1-300
KDI_TIMELINE=1 ROUNDS=1 REPS=3 SETTINGS="split=0" timeout 300 python tools/schedule_sweep.py > gpurun_out/s12_timeline.out 2> gpurun_out/s12_timeline.txt
awk '/kdi timeline/{c++} c==3' gpurun_out/s12_timeline.txt
timeout 300 python tools/project_time.py > gpurun_out/s12_project_time.txt 2>&1; tail -5 gpurun_out/s12_project_time.txt
GEN=1 KDI_TIMELINE=1 timeout 300 python tools/timeline.py > gpurun_out/s12_gen_timeline.out 2> gpurun_out/s12_gen_timeline.txt
awk '/kdi timeline/{c++} c==4' gpurun_out/s12_gen_timeline.txt; tail -1 gpurun_out/s12_gen_timeline.out
bash tools/gpu_profile.sh
