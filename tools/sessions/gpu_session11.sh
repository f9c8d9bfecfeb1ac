#!/bin/bash
mkdir -p gpurun_out
# C4 shard shape: 100 000 x 37 500, keep_n 50 (kc = 64) vs keep_n 20 (kc = 32): GEMM time + ncu
for k in 50 20; do
  OVERLAP=0 KEEP=$k M=100000 N=37500 KDI_TIMELINE=1 timeout 300 python tools/timeline.py 2>&1 | awk '/====/{p=1} p'
done
OVERLAP=0 KEEP=50 M=100000 N=37500 timeout 600 ncu --set full --clock-control none --import-source on -k regex:kdi_gemm_kernel -s 2 -c 1 -o gpurun_out/prof_gemm_kc64 -f python tools/timeline.py > gpurun_out/ncu_kc64.log 2>&1
echo "ncu exit $?"
ncu -i gpurun_out/prof_gemm_kc64.ncu-rep --page raw --csv 2>/dev/null | python - <<'PY'
import csv,sys
rows=list(csv.reader(sys.stdin))
hdr=rows[0]; vals=rows[2] if len(rows)>2 else rows[1]
want=["gpu__time_duration.sum","sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed","dram__bytes_read.sum","dram__bytes_write.sum","lts__t_sector_hit_rate.pct","sm__cycles_elapsed.avg.per_second","launch__registers_per_thread","smsp__inst_executed.sum","launch__shared_mem_per_block_dynamic"]
for h,v in zip(hdr,vals):
    if h in want: print(h,v)
PY
