"""Timing of merge_crystal_maps on the device against the oracle (run on the GPU box).

  python tests/gpu_tools/merge_time.py [map side] [scores per point] [maps]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))
import kikuchipy_b200 as kb  # noqa: E402
from oracle import merge_oracle as mo  # noqa: E402

side = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
k = int(sys.argv[3]) if len(sys.argv) > 3 else 2
rng = np.random.default_rng(0)
m = side * side
scores = [(-np.sort(-rng.random((m, n), dtype=np.float32), axis=1)) for _ in range(k)]
rots = [rng.normal(size=(m, n, 4)) for _ in range(k)]
idx = [rng.integers(0, 100000, (m, n)) for _ in range(k)]
ctx = kb.default_context()
args = (scores, rots, idx, [None] * k, [np.zeros(m, np.uint8)] * k, m, 3, 1, False)
ctx.merge_crystal_maps(*args)
t0 = time.time()
out = ctx.merge_crystal_maps(*args)
wall = time.time() - t0
ms = ctx.timings()["total_ms"]
bytes_alg = m * k * n * (2 * 4 + 32 + 2 * 8) + m * n * (4 + 32 + 4) + m * k * n * (4 + 8) + m * 8
t0 = time.time()
want = mo.merge_arrays([{"scores": s, "rotations": r, "simulation_indices": i, "phase_id": np.zeros(m)} for s, r, i in zip(scores, rots, idx)],
                       [None] * k, m, 3, None, True)
cpu = time.time() - t0
same = all(np.array_equal(out[key].reshape(want[key].shape), want[key], equal_nan=True) for key in
           ("phase_id", "scores", "merged_scores", "simulation_indices", "merged_simulation_indices", "rotations"))
print(json.dumps({"map": [side, side], "scores_per_point": n, "maps": k, "kernel_ms": round(ms, 3), "call_ms": round(wall * 1e3, 1),
                  "algorithmic_GB": round(bytes_alg / 1e9, 3), "GBps": round(bytes_alg / ms / 1e6, 1),
                  "oracle_s": round(cpu, 2), "identical_to_oracle": bool(same)}))
