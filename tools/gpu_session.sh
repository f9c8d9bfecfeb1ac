#!/bin/bash
# Session 27: ncu capture of the projection kernel (own arithmetic, tap tables).
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kdi_project_kernel -s 2 -c 1 -o gpurun_out/r2_prof_kdi_project_kernel -f env N=100000 CONFIGS=1001:0 python tools/project_time.py > gpurun_out/ncu_project.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_project.log
