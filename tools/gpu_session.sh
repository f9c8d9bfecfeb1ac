#!/bin/bash
# Session 28: generated dictionary projected in one piece beside the experimental upload; whole suite.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/s28_pytest.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/s28_pytest.log
for i in 1 2; do
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu --no-extras > gpurun_out/s28_bench_n1_$i.json 2> gpurun_out/s28_bench_n1_$i.err
echo "bench exit $?"; python - <<PY
import json
for l in open('gpurun_out/s28_bench_n1_$i.json'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print({k:d[k] for k in ('value','ms_per_step','e2e_generated') if k in d})
PY
done
KDI_TIMELINE=1 timeout 300 python - <<'PY' > gpurun_out/s28_timeline_generated.txt 2>&1
import numpy as np, torch, sys
sys.path.insert(0, '.')
import kikuchipy_b200 as kb
from kikuchipy_b200 import synthetic as po
mu, ml = po.synthetic_master_pattern(1001, seed=5)
dc = kb.direction_cosines([-0.9, 0.85, -0.7, 0.95], 0.5, 60, 60, po.tilted_detector_matrix(70.0))
rot = po.random_rotations(100000, seed=4)
gen = kb.get_patterns(mu, ml, rot, direction_cosines=dc, detector_shape=(60, 60))
exp = np.random.default_rng(1).integers(0, 256, (10000, 60, 60), dtype=np.uint8)
for _ in range(3):
    res = kb.dictionary_indexing(exp, gen, metric="ncc", keep_n=20, verbose=False)
PY
tail -30 gpurun_out/s28_timeline_generated.txt
