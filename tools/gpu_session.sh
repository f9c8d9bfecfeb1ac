#!/bin/bash
# Session 37: eight GPUs - the N = 8 bench with the full-size configurations 4, 5 and 5-generated.
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu > gpurun_out/s37_bench_n8.json 2> gpurun_out/s37_bench_n8.err
echo "bench n8 exit $?"; python - <<'PY'
import json
for l in open('gpurun_out/s37_bench_n8.json'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print({k:d[k] for k in ('value','ms_per_step','n_gpus','e2e','e2e_generated','parity') if k in d}); print(json.dumps(d.get('extra')))
PY
tail -3 gpurun_out/s37_bench_n8.err
