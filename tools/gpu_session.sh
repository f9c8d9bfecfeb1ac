#!/bin/bash
# Session 21: whole suite, sweep, default bench (N = 1, with the extras), profile set for profiles/.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/s21_pytest.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/s21_pytest.log
SETTINGS="view=1;view=0" timeout 600 python tools/schedule_sweep.py > gpurun_out/s21_sweep.jsonl 2> gpurun_out/s21_sweep.err
echo "sweep exit $?"; cat gpurun_out/s21_sweep.jsonl
KDI_TIMELINE=1 ROUNDS=1 REPS=3 SETTINGS="view=1;view=0" timeout 300 python tools/schedule_sweep.py > gpurun_out/s21_timeline.out 2> gpurun_out/s21_timeline.txt
awk '/kdi timeline/{c++} c==3||c==6' gpurun_out/s21_timeline.txt
timeout 1500 python bench.py > gpurun_out/s21_bench_n1.json 2> gpurun_out/s21_bench_n1.err
echo "bench exit $?"; python - <<'PY'
import json
for l in open('gpurun_out/s21_bench_n1.json'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print({k:d[k] for k in ('value','ms_per_step','e2e','e2e_pageable','e2e_generated','roofline','parity','cpu_baseline','gpu_launches','clocks') if k in d}); print(json.dumps(d.get('extra'))[:1500])
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/s21_bench_ref.json 2> gpurun_out/s21_bench_ref.err
echo "reference arm exit $?"; tail -c 600 gpurun_out/s21_bench_ref.json
bash tools/gpu_profile.sh
