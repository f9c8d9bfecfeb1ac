#!/bin/bash
# Session 51: ncu on the cuBLAS kernels of the same-box comparison (tensor-pipe activity, clock).
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.avg.per_second,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tensor.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct,l1tex__m_xbar2l1tex_read_bytes.sum,launch__grid_size,launch__block_size,launch__cluster_dim_x,launch__shared_mem_per_block_dynamic --clock-control none -s 8 -c 12 --csv --log-file gpurun_out/s51_cublas_ncu.csv python tools/probes/cublas_ncu_target.py > gpurun_out/s51_cublas_ncu.log 2>&1
echo "ncu exit $?"; python - <<'PY'
import csv
rows=list(csv.DictReader(l for l in open('gpurun_out/s51_cublas_ncu.csv') if not l.startswith('==')))
by={}
for r in rows:
    by.setdefault((r['ID'], r['Kernel Name'][:60]), {})[r['Metric Name']] = r['Metric Value']
for k,v in list(by.items()):
    print(k, v)
PY
