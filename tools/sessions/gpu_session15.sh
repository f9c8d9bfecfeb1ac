#!/bin/bash
mkdir -p gpurun_out
run() { # label M N overlap settings
  M=$2 N=$3 OVERLAP=$4 ROUNDS=1 REPS=1 SETTINGS="$5" timeout 600 ncu --metrics dram__bytes_read.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,l1tex__m_xbar2l1tex_read_bytes.sum --clock-control none -k regex:kdi_gemm_kernel --csv --log-file gpurun_out/l2_$1.csv python tools/gemm_sweep.py > /dev/null 2>&1
  python - "$1" <<'PY'
import csv,collections,sys
rows=list(csv.reader(open('gpurun_out/l2_%s.csv'%sys.argv[1])))
hdr=None; per=collections.OrderedDict()
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r)); per.setdefault(int(d['ID']),{})[d['Metric Name']]=float(d['Metric Value'])
t=sum(v['gpu__time_duration.sum'] for v in per.values())/1e6
b=sum(v['dram__bytes_read.sum'] for v in per.values())/1e9
x=sum(v['l1tex__m_xbar2l1tex_read_bytes.sum'] for v in per.values())/1e9
print(sys.argv[1], len(per),'launches: %.2f ms, dram read %.2f GB, L2->SM %.1f GB'%(t,b,x))
PY
}
run smallM_fullN 2560 100000 0 "0,0,0,0"
run fullM_smallN 10000 5120 0 "0,0,0,0"
run c2_serial 10000 100000 0 "0,0,0,0"
run c2_overlap 10000 100000 1 "0,0,0,0"
run c2_overlap_st6 10000 100000 1 "0,6,0,0"
run c2_overlap_st4 10000 100000 1 "0,4,0,0"
run c2_overlap_sb37 10000 100000 1 "37,0,0,0"
run c2_overlap_sb10 10000 100000 1 "10,0,0,0"
run c2_overlap_sb14 10000 100000 1 "14,0,0,0"
