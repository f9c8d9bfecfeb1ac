#!/bin/bash
# Session 33: two GPUs - sharded tests, smoke(), short N = 2 bench with the generated-dictionary leg.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -q > gpurun_out/s33_pytest_sharded.log 2>&1
echo "pytest sharded exit $?"; tail -3 gpurun_out/s33_pytest_sharded.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s33_smoke.log 2>&1
echo "smoke exit $?"; tail -4 gpurun_out/s33_smoke.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu --no-extras > gpurun_out/s33_bench_n2.json 2> gpurun_out/s33_bench_n2.err
echo "bench n2 exit $?"; python - <<'PY'
import json
for l in open('gpurun_out/s33_bench_n2.json'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print({k:d[k] for k in ('value','ms_per_step','n_gpus','e2e','e2e_generated','parity') if k in d})
PY
