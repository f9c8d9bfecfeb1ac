#!/bin/bash
mkdir -p gpurun_out
export M=100000 N=37500
S="0,0,0,0;0,8,0,0;0,4,0,0;0,15,0,0;12,8,0,0;12,4,0,0;8,8,0,0;16,8,0,0;12,15,0,0;25,8,1,0"
ROUNDS=2 REPS=2 SETTINGS="$S" timeout 900 python tools/gemm_sweep.py > gpurun_out/sweep_c4_time.jsonl 2> gpurun_out/sweep_c4_time.err
echo "sweep exit $?"; cat gpurun_out/sweep_c4_time.jsonl; tail -2 gpurun_out/sweep_c4_time.err
ROUNDS=1 REPS=1 SETTINGS="$S" timeout 900 ncu --metrics dram__bytes_read.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,sm__cycles_elapsed.avg.per_second --clock-control none -k regex:kdi_gemm_kernel --csv --log-file gpurun_out/sweep_c4_ncu.csv python tools/gemm_sweep.py > gpurun_out/sweep_c4_ncu.log 2>&1
echo "ncu exit $?"
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/sweep_c4_ncu.csv')))
hdr=None; per=collections.OrderedDict()
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r)); per.setdefault(d['ID'],{})[d['Metric Name']]=d['Metric Value']
for k,v in per.items(): print(k, v)
PY
