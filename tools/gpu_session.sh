#!/bin/bash
# Session 58: eight GPUs - the N = 8 bench (default flags, as the driver's scaling run launches it) with the final tree.
mkdir -p gpurun_out
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/s58_bench_n8.json 2> gpurun_out/s58_bench_n8.err
echo "bench n8 exit $?"; python - <<'PY'
import json
for l in open('gpurun_out/s58_bench_n8.json'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print({k:d[k] for k in ('value','ms_per_step','n_gpus','e2e','e2e_generated','parity') if k in d})
        for k,v in (d.get('extra') or {}).items():
            print(k, {kk: v.get(kk) for kk in ('ms_per_step','patterns_per_s','rank0_stage_ms','rank0_gemm_tflops_algorithmic','checks','error')})
PY
tail -2 gpurun_out/s58_bench_n8.err
