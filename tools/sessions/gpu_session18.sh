#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest exit $?")
tail -8 gpurun_out/pytest_gpu2.log
(timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3)
