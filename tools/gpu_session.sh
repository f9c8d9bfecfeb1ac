#!/bin/bash
# Session 47: register-resident gather normalise for masked rows - parity, then BASELINE configs[2] A/B.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/s47_pytest.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/s47_pytest.log
for regs in 1 0 1 0; do
echo "KDI_OPT_GATHER_REGS=$regs"
env KDI_TIMELINE=1 CONFIG=3 SAMPLE64=64 OPTS="23=$regs" timeout 600 python tools/config_timeline.py > gpurun_out/s47_c3_regs$regs.txt 2>&1
grep "normalize" gpurun_out/s47_c3_regs$regs.txt | tail -2
tail -1 gpurun_out/s47_c3_regs$regs.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['rank0_stage_ms'], d['checks'])"
done
