#!/bin/bash
# Session 64: strict certificate (KDI_OPT_CERT_STRICT) - its tests, the whole suite, smoke, default bench.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "strict" > gpurun_out/s64_strict.log 2>&1
echo "strict tests exit $?"; grep -a "strict certificate\|passed\|failed\|Error\|assert" gpurun_out/s64_strict.log | head -20
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s64_pytest.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/s64_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s64_smoke.log 2>&1
echo "smoke exit $?"; tail -1 gpurun_out/s64_smoke.log | cut -c1-300
timeout 1200 python bench.py > gpurun_out/s64_bench_n1.json 2> gpurun_out/s64_bench_n1.err
echo "bench exit $?"; python - <<'PY'
import json
for l in open('gpurun_out/s64_bench_n1.json'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print({k:d[k] for k in ('value','ms_per_step') if k in d}); print(json.dumps(d.get('parity'))[:1200])
PY
