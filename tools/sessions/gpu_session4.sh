#!/bin/bash
# GPU session 4: tests (overlapped schedule), bench with overlap on/off
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?")
tail -15 gpurun_out/pytest_gpu.log
for ov in 1 0 1 0; do
timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu --e2e-steps 2 --overlap $ov > gpurun_out/bench_ov$ov.json 2> gpurun_out/bench_ov$ov.err
echo "bench overlap=$ov exit $?"; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_ov$ov.json").read().strip().splitlines()[-1])
print("value=%.0f ms=%.3f e2e=%.0f"%(d["value"],d["ms_per_step"],d["e2e"]["value"]), d["config"]["stage_ms"], d["clocks"], "frac %.3f"%d["roofline"]["frac"])
PY
tail -3 gpurun_out/bench_ov$ov.err
done
