#!/bin/bash
# Evidence for profiles/ (round 2): launch list of a short bench run + one `ncu --set full` capture per
# kernel of the path.  Reports land in gpurun_out/; tools/ncu_summary.py turns them into the summaries
# kept under profiles/.  (The kernels of the peer-memory exchange only run with several ranks, which ncu
# must not wrap: their times come from the KDI_TIMELINE trace, profiles/r2_timeline_n8_rank0.txt.)
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-generated --no-extras --e2e-steps 1"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv $B > gpurun_out/ncu_launch.log 2>&1
echo "launch list exit $?"
S="env ROUNDS=1 REPS=3 SETTINGS=split=1 python tools/schedule_sweep.py"
for k in kdi_gemm_kernel kdi_select_rescore_kernel kdi_select_warp_kernel; do
  skip=2; [ $k = kdi_gemm_kernel ] && skip=4
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -o gpurun_out/r2_prof_$k -f $S > gpurun_out/ncu_$k.log 2>&1
  echo "ncu $k exit $?"
done
# the dictionary prepare launch (warp per row, view mode: 16-bit rows only) - the 2nd normalise launch of a call
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kdi_normalize_warp_rows -s 3 -c 1 -o gpurun_out/r2_prof_kdi_normalize_warp_rows_dict -f $S > gpurun_out/ncu_norm_dict.log 2>&1
echo "ncu normalize (dictionary) exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kdi_normalize_warp_rows -s 2 -c 1 -o gpurun_out/r2_prof_kdi_normalize_warp_rows_exp -f $S > gpurun_out/ncu_norm_exp.log 2>&1
echo "ncu normalize (experimental) exit $?"
# the 64-entry-list variant at the shape of one rank's share of BASELINE configs[3]
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kdi_gemm_kernel -s 1 -c 1 -o gpurun_out/r2_prof_kdi_gemm_kernel_kc64 -f \
  env M=100000 N=37500 KEEP=50 ROUNDS=1 REPS=2 SETTINGS=split=1 python tools/schedule_sweep.py > gpurun_out/ncu_gemm_kc64.log 2>&1
echo "ncu gemm kc64 exit $?"
# masked uint8 rows (BASELINE configs[2] at quarter scale): the staged normalise kernel
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kdi_normalize_staged -s 0 -c 1 -o gpurun_out/r2_prof_kdi_normalize_staged -f \
  env CONFIG=3 SCALE=0.25 SAMPLE64=0 python tools/config_timeline.py > gpurun_out/ncu_normalize_staged.log 2>&1
echo "ncu normalize_staged exit $?"
ls -la gpurun_out/*.ncu-rep
