#!/bin/bash
# The command set of the current GPU session (rewritten per session; results land in gpurun_out/).
# Session 18: two GPUs - the whole GPU suite (incl. the sharded tests), smoke, a short 2-rank bench.
mkdir -p gpurun_out
nvidia-smi -L
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s18_pytest.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/s18_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s18_smoke.log 2>&1
echo "smoke exit $?"; tail -3 gpurun_out/s18_smoke.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 --no-generated > gpurun_out/s18_bench_n2.json 2> gpurun_out/s18_bench_n2.err
echo "bench n2 exit $?"; python - <<'PY'
import json
for l in open('gpurun_out/s18_bench_n2.json'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print({k:d[k] for k in ('value','ms_per_step','n_gpus','e2e','roofline','parity') if k in d})
PY
SETTINGS="view=1;view=0" timeout 600 python tools/schedule_sweep.py > gpurun_out/s18_sweep.jsonl 2> gpurun_out/s18_sweep.err
echo "sweep exit $?"; cat gpurun_out/s18_sweep.jsonl
