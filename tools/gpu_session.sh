#!/bin/bash
# Session 35: compute-sanitizer racecheck / synccheck over smoke(); memcheck over the whole GPU suite.
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 9 --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s35_synccheck_smoke.log 2>&1
echo "synccheck smoke exit $?"; grep -E "ERROR SUMMARY|smoke ok|Error|Barrier|barrier" gpurun_out/s35_synccheck_smoke.log | head -10
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s35_racecheck_smoke.log 2>&1
echo "racecheck smoke exit $?"; grep -E "RACECHECK SUMMARY|ERROR SUMMARY|smoke ok|hazard|Race" gpurun_out/s35_racecheck_smoke.log | head -20
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests -m gpu -q -x > gpurun_out/s35_memcheck_all.log 2>&1
echo "memcheck all exit $?"; grep -E "ERROR SUMMARY|Invalid|passed|failed|Error" gpurun_out/s35_memcheck_all.log | head -20
