#!/bin/bash
# Session 38: A/B of the result copy-out (pinned ring vs the driver's staging) on the host-input legs.
mkdir -p gpurun_out
for v in 0 1 0 1; do
KDI_COPY_OUT_DIRECT=$v timeout 600 python bench.py --steps 5 --warmup 3 --e2e-steps 8 --no-cpu --no-extras --no-generated > gpurun_out/s38_bench_$v.json 2> gpurun_out/s38_bench_$v.err
python - <<PY
import json
for l in open('gpurun_out/s38_bench_$v.json'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print("direct=$v", round(d['e2e']['ms_per_step'],2), round(d['e2e_pageable']['ms_per_step'],2))
PY
done
