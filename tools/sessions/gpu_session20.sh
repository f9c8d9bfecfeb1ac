#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?")
tail -6 gpurun_out/pytest_gpu.log
timeout 900 python tests/gpu_tools/run_config.py --config 3 --sample-oracle 0 > gpurun_out/config3.json 2> gpurun_out/config3.err
echo "config 3 exit $?"; cat gpurun_out/config3.json; tail -3 gpurun_out/config3.err
for cfg in "KEEP=20 M=80000 N=12500" "KEEP=50 M=100000 N=37500"; do
  env $cfg KDI_TIMELINE=1 timeout 300 python tools/timeline.py 2>&1 | awk '/====/{p=1} p' | grep -v "gemm_topk"
done
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_n1d.json 2> gpurun_out/bench_n1d.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n1d.json").read().strip().splitlines()[-1])
print("value=%.0f ms=%.3f e2e=%.0f"%(d["value"],d["ms_per_step"],d["e2e"]["value"]), d["config"]["stage_ms"], "frac %.3f"%d["roofline"]["frac"], "gen", d["e2e_generated"]["value"], d["e2e_generated"]["ms_per_step"])
PY
