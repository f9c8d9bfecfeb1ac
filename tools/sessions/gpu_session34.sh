#!/bin/bash
# final confirmation of the round: the whole GPU suite on the final build
(timeout 60 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest exit $?")
tail -4 gpurun_out/pytest_gpu_final.log
