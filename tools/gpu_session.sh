#!/bin/bash
# The command set of the current GPU session (rewritten per session; results land in gpurun_out/).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s13_pytest.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/s13_pytest.log
SETTINGS="split=1;split=0" timeout 600 python tools/schedule_sweep.py > gpurun_out/s13_sweep.jsonl 2> gpurun_out/s13_sweep.err
echo "sweep exit $?"; cat gpurun_out/s13_sweep.jsonl
KDI_TIMELINE=1 ROUNDS=1 REPS=3 SETTINGS="split=1;split=0" timeout 300 python tools/schedule_sweep.py > gpurun_out/s13_timeline.out 2> gpurun_out/s13_timeline.txt
awk '/kdi timeline/{c++} c==3||c==6' gpurun_out/s13_timeline.txt
for o in "18=0"; do
  OPTS=$o KDI_TIMELINE=1 CONFIG=3 SAMPLE64=64 timeout 300 python tools/config_timeline.py > gpurun_out/s13_c3.json 2> gpurun_out/s13_c3.txt
  grep normalize gpurun_out/s13_c3.txt | tail -2; python -c "
import json;r=json.load(open('gpurun_out/s13_c3.json'));print(r['ms_per_step'], r['rank0_stage_ms'], r['checks'])"
done
S="env ROUNDS=1 REPS=3 SETTINGS=split=1 python tools/schedule_sweep.py"
for k in kdi_select_rescore_kernel kdi_select_warp_kernel kdi_normalize_f32_regs; do
  skip=2; [ $k = kdi_normalize_f32_regs ] && skip=5
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -o gpurun_out/r2_prof_$k -f $S > gpurun_out/ncu_$k.log 2>&1
  echo "ncu $k exit $?"
done
