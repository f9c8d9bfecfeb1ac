#!/bin/bash
# One GPU session: tests, bench variants, ncu launch list + full capture of the GEMM kernel.
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?")
tail -3 gpurun_out/pytest_gpu.log
for cg in 1 2; do
  timeout 400 python bench.py --steps 10 --warmup 3 --cta-group $cg $([ $cg = 2 ] && echo --no-cpu) > gpurun_out/bench_cg$cg.json 2> gpurun_out/bench_cg$cg.err
  echo "bench cg=$cg exit $?"
done
for st in 4 16 32 64; do
  for cg in 1 2; do
    timeout 200 python bench.py --steps 5 --warmup 3 --cta-group $cg --strip-tiles $st --no-cpu --e2e-steps 1 > gpurun_out/bench_cg${cg}_st$st.json 2>/dev/null
  done
done
timeout 200 python bench.py --steps 5 --warmup 3 --cta-group 2 --compute-dtype bf16 --no-cpu --e2e-steps 1 > gpurun_out/bench_cg2_bf16.json 2>/dev/null
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value=%.0f e2e=%.0f" % (d["value"], d["e2e"]["value"]), d["config"]["stage_ms"], "flag", d["config"]["flagged_rows_last_step"], "frac %.3f" % d["roofline"]["frac"])
    except Exception as e:
        print(f, "ERR", e)
PY
# launch list (every kernel with its device time) of one short bench run
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 --cta-group 2 > gpurun_out/ncu_launch.log 2>&1
echo "ncu launches exit $?"
for cg in 1 2; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:kdi_gemm_kernel -s 3 -c 1 -o gpurun_out/prof_gemm_cg$cg -f python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 --cta-group $cg > gpurun_out/ncu_full_cg$cg.log 2>&1
  echo "ncu full cg=$cg exit $?"
done
ls -la gpurun_out
