#!/bin/bash
# Session 48: staged prepare kernel copying kept columns run by run - parity, then BASELINE configs[2] A/B.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/s48_pytest.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/s48_pytest.log
for cols in 0 1 0 1; do
echo "KDI_OPT_STAGED_COLS=$cols"
env KDI_TIMELINE=1 CONFIG=3 SAMPLE64=64 OPTS="23=$cols" timeout 600 python tools/config_timeline.py > gpurun_out/s48_c3_cols$cols.txt 2>&1
tail -1 gpurun_out/s48_c3_cols$cols.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['rank0_stage_ms'], d['checks']['f64_rows_identical'])"
done
