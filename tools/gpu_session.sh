#!/bin/bash
# Session 31: refinement after the 32-bit index division; PCIe probe; timeline of the host-dictionary path.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_refinement.py tests/test_gpu_projection.py -m gpu -q > gpurun_out/s31_pytest.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/s31_pytest.log
timeout 600 python tests/gpu_tools/refine_time.py > gpurun_out/s31_refine_time.txt 2>&1; tail -3 gpurun_out/s31_refine_time.txt
timeout 300 python tools/probes/h2d_probe.py > gpurun_out/s31_h2d_probe.txt 2>&1; cat gpurun_out/s31_h2d_probe.txt
KDI_TIMELINE=1 timeout 300 python - <<'PY' > gpurun_out/s31_timeline_host_dict.txt 2>&1
import numpy as np, torch, sys
sys.path.insert(0, '.')
import kikuchipy_b200 as kb
ctx = kb.default_context(0)
exp = ctx.pinned_empty((10000, 60, 60), np.uint8); exp[:] = np.random.default_rng(1).integers(0, 256, exp.shape, dtype=np.uint8)
dic = ctx.pinned_empty((100000, 60, 60), np.float32); dic[:] = np.random.default_rng(2).random(dic.shape, dtype=np.float32)
for _ in range(3):
    res = kb.dictionary_indexing(exp, dic, metric="ncc", keep_n=20, verbose=False)
PY
tail -45 gpurun_out/s31_timeline_host_dict.txt
