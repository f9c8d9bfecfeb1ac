#!/bin/bash
# What the driver runs at round end on one GPU: whole suite, smoke(), our bench arm with default flags.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/final_pytest.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/final_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1
echo "smoke exit $?"; tail -2 gpurun_out/final_smoke.log | cut -c1-400
timeout 1200 python bench.py > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err
echo "bench exit $?"; python - <<'PY'
import json
for l in open('gpurun_out/final_bench_n1.json'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print({k:d[k] for k in ('value','ms_per_step','e2e','roofline','clocks') if k in d}); print(json.dumps(d.get('parity'))[:1200]); print(json.dumps(d.get('extra'))[:600])
PY
