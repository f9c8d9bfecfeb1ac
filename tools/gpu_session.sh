#!/bin/bash
# Session 19: two GPUs - the whole GPU suite (incl. the sharded tests), all failures reported.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/s19_pytest.log 2>&1
echo "pytest exit $?"; tail -15 gpurun_out/s19_pytest.log
