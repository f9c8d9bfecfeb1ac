#!/bin/bash
# What the driver runs at round end on one GPU: whole suite, smoke(), our bench arm with default flags
# (+ the certificate tests with their printed counts and the interleaved A/B of the certificate modes).
mkdir -p gpurun_out
ROUNDS=6 REPS=3 SETTINGS="cert=0;cert=2" timeout 300 python tools/schedule_sweep.py > gpurun_out/final_cert_sweep.jsonl 2> gpurun_out/final_cert_sweep.err
echo "sweep exit $?"; cut -c1-330 gpurun_out/final_cert_sweep.jsonl
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "strict" > gpurun_out/final_strict.log 2>&1
echo "strict tests exit $?"; grep -a "certificate" gpurun_out/final_strict.log | sed 's/^\.//' | head
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
        d=json.loads(l); print({k:d[k] for k in ('value','ms_per_step') if k in d}, d['roofline']['achieved'], d['e2e']['ms_per_step'], d['clocks']); print(json.dumps(d.get('parity'))[:1200]); print(json.dumps(d.get('extra'))[:300])
PY
