#!/bin/bash
# The command set of the current GPU session (rewritten per session; results land in gpurun_out/).
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/s16_pytest.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/s16_pytest.log
SETTINGS="view=1;view=0;view=1,split=0;view=0,split=0" timeout 600 python tools/schedule_sweep.py > gpurun_out/s16_sweep.jsonl 2> gpurun_out/s16_sweep.err
echo "sweep exit $?"; cat gpurun_out/s16_sweep.jsonl
KDI_GEMM_FULL_K=1 SETTINGS="view=1;view=0" timeout 600 python tools/schedule_sweep.py > gpurun_out/s16_sweep_fullk.jsonl 2> gpurun_out/s16_sweep_fullk.err
echo "sweep full-K exit $?"; cat gpurun_out/s16_sweep_fullk.jsonl
SETTINGS="view=1;view=0" timeout 600 python tools/schedule_sweep.py > gpurun_out/s16_sweep2.jsonl 2> gpurun_out/s16_sweep2.err
echo "sweep (again, short K tail) exit $?"; cat gpurun_out/s16_sweep2.jsonl
KDI_TIMELINE=1 ROUNDS=1 REPS=3 SETTINGS="view=1;view=0;view=1,split=0" timeout 300 python tools/schedule_sweep.py > gpurun_out/s16_timeline.out 2> gpurun_out/s16_timeline.txt
awk '/kdi timeline/{c++} c==3||c==6||c==9' gpurun_out/s16_timeline.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kdi_normalize_warp_rows -s 3 -c 1 -o gpurun_out/r2b_prof_kdi_normalize_warp_rows_dict -f env ROUNDS=1 REPS=2 SETTINGS=split=0 python tools/schedule_sweep.py > gpurun_out/ncu_norm_dict.log 2>&1
echo "ncu normalize dict exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kdi_select_rescore -s 1 -c 1 -o gpurun_out/r2b_prof_kdi_select_rescore_view -f env ROUNDS=1 REPS=2 SETTINGS=split=0 python tools/schedule_sweep.py > gpurun_out/ncu_rescore_view.log 2>&1
echo "ncu rescore view exit $?"
