#!/bin/bash
# The command set of the current GPU session (rewritten per session; results land in gpurun_out/).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s9_pytest.log 2>&1
echo "pytest exit $?"; tail -8 gpurun_out/s9_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/s9_bench.json 2> gpurun_out/s9_bench.err
echo "bench exit $?"; head -c 1500 gpurun_out/s9_bench.json; tail -3 gpurun_out/s9_bench.err
python - <<'PY'
import json
b=json.load(open('gpurun_out/s9_bench.json'))
print(json.dumps(b.get('extra'))[:2500])
PY
