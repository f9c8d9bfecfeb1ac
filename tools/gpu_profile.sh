#!/bin/bash
# Final evidence for profiles/: launch list of a short bench run + one full ncu capture per kernel.
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-generated --e2e-steps 1"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r1_launches_bench.csv $B > gpurun_out/ncu_launch.log 2>&1
echo "launch list exit $?"
for k in kdi_gemm_kernel kdi_select_rescore_kernel kdi_select_warp_kernel kdi_normalize_f32_regs kdi_normalize_staged; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -o gpurun_out/prof_$k -f $B > gpurun_out/ncu_$k.log 2>&1
  echo "ncu $k exit $?"
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:kdi_project_kernel -s 1 -c 1 -o gpurun_out/prof_kdi_project_kernel -f python tools/project_time.py > gpurun_out/ncu_project.log 2>&1
echo "ncu project exit $?"
ls -la gpurun_out/*.ncu-rep
