#!/bin/bash
# The command set of the current GPU session (rewritten per session; results land in gpurun_out/).
mkdir -p gpurun_out
timeout 120 tools/probes/coresidency_probe > gpurun_out/s4_probe.txt 2>&1
echo "probe exit $?"; cat gpurun_out/s4_probe.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "config2_full or schedule_variants" > gpurun_out/s4_pytest_a.log 2>&1
echo "pytest A exit $?"; tail -15 gpurun_out/s4_pytest_a.log
timeout 600 python tools/schedule_sweep.py > gpurun_out/s4_sweep.jsonl 2> gpurun_out/s4_sweep.err
echo "sweep exit $?"; cat gpurun_out/s4_sweep.jsonl; tail -3 gpurun_out/s4_sweep.err
