#!/bin/bash
# Session 23: two GPUs - whole suite (sharded tests with the forced view mode), short N = 2 bench.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/s23_pytest.log 2>&1
echo "pytest exit $?"; tail -6 gpurun_out/s23_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 --no-generated --no-cpu > gpurun_out/s23_bench_n2.json 2> gpurun_out/s23_bench_n2.err
echo "bench n2 exit $?"; python - <<'PY'
import json
for l in open('gpurun_out/s23_bench_n2.json'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print({k:d[k] for k in ('value','ms_per_step','n_gpus','e2e','parity') if k in d}); print(d.get('detail'))
PY
