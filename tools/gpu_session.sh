#!/bin/bash
# The command set of the current GPU session (rewritten per session; results land in gpurun_out/).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s14_pytest.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/s14_pytest.log
SETTINGS="split=1;split=0" timeout 600 python tools/schedule_sweep.py > gpurun_out/s14_sweep.jsonl 2> gpurun_out/s14_sweep.err
echo "sweep exit $?"; cat gpurun_out/s14_sweep.jsonl
KDI_TIMELINE=1 ROUNDS=1 REPS=3 SETTINGS="split=1;split=0" timeout 300 python tools/schedule_sweep.py > gpurun_out/s14_timeline.out 2> gpurun_out/s14_timeline.txt
awk '/kdi timeline/{c++} c==3||c==6' gpurun_out/s14_timeline.txt
M=100000 N=37500 KEEP=50 KDI_TIMELINE=1 ROUNDS=1 REPS=2 SETTINGS="split=1" timeout 300 python tools/schedule_sweep.py > gpurun_out/s14_c4.out 2> gpurun_out/s14_c4.txt
awk '/kdi timeline/{c++} c==2' gpurun_out/s14_c4.txt | grep -v gemm_topk; cat gpurun_out/s14_c4.out
S="env ROUNDS=1 REPS=3 SETTINGS=split=1 python tools/schedule_sweep.py"
for k in kdi_select_warp_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o gpurun_out/r2_prof_$k -f $S > gpurun_out/ncu_$k.log 2>&1
  echo "ncu $k exit $?"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kdi_normalize_f32_regs -s 3 -c 1 -o gpurun_out/r2_prof_kdi_normalize_f32_regs_dict -f env ROUNDS=1 REPS=2 SETTINGS=split=0 python tools/schedule_sweep.py > gpurun_out/ncu_norm_dict.log 2>&1
echo "ncu normalize dict exit $?"
