#!/bin/bash
# The command set of the current GPU session (rewritten per session; results land in gpurun_out/).
mkdir -p gpurun_out
for o in "18=1" "18=0"; do
  OPTS=$o KDI_TIMELINE=1 CONFIG=3 timeout 300 python tools/config_timeline.py > gpurun_out/s10_c3_$o.json 2> gpurun_out/s10_c3_$o.txt
  echo "c3 $o exit $?"; tail -12 gpurun_out/s10_c3_$o.txt; python -c "
import json;r=json.load(open('gpurun_out/s10_c3_$o.json'));print(r['ms_per_step'], r['rank0_stage_ms'], r['checks'])"
done
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s10_pytest.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/s10_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/s10_bench.json 2> gpurun_out/s10_bench.err
echo "bench exit $?"; python -c "
import json;b=json.load(open('gpurun_out/s10_bench.json'));print(b['ms_per_step'], b['detail']['stage_ms']); print('e2e', b['e2e']['ms_per_step'], 'pageable', b['e2e_pageable']['ms_per_step'], 'generated', b['e2e_generated']['ms_per_step'], b['e2e_generated']['rank0_stage_ms'])"
