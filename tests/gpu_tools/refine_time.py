"""Timing of the refinement kernel on a realistic size (run on the GPU box through gpurun), with the
oracle (NumPy port of the reference's solver) timed beside it on a few patterns.

  python tests/gpu_tools/refine_time.py [n_patterns] [detector side] [master pattern side]

Prints one JSON line: patterns/s on the device (kernel time, CUDA events, and wall time of the
whole call from host patterns), mean objective evaluations per pattern, pixel-evaluations/s, the
CPU port's patterns/s and the parity of the first patterns against it."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))
import kikuchipy_b200 as kb  # noqa: E402
from kikuchipy_b200 import _lib  # noqa: E402
from oracle import refinement_oracle as ro  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
side = int(sys.argv[2]) if len(sys.argv) > 2 else 60
mps = int(sys.argv[3]) if len(sys.argv) > 3 else 1001
n_cpu = 16
base = ro.synthetic_case(n=256, nrows=side, ncols=side, mp_size=mps, seed=11, noise=0.05, perturb_deg=1.0)
reps = -(-n // 256)
pats = np.tile(base["patterns"], (reps, 1))[:n]
x0 = np.tile(base["start_eulers"], (reps, 1))[:n][:, None, :]
p = base["problem"]
ctx = kb.default_context()
mp = ctx.master_pattern(p.mu, p.ml, p.dc)
out = {}
for mode, name in ((_lib.REFINE_ORI, "orientation"), (_lib.REFINE_ORI_PC, "orientation+pc")):
    xx = x0 if mode == _lib.REFINE_ORI else np.concatenate([x0, np.tile(base["pc"], (n, 1, 1))], axis=2)
    ctx.refine(mp, mode, pats[:512], side, side, False, xx[:512], om_detector_to_sample=p.om)  # warm-up
    t0 = time.time()
    res = ctx.refine(mp, mode, pats, side, side, False, xx, om_detector_to_sample=p.om)
    wall = time.time() - t0
    ms = ctx.timings()["total_ms"]
    evals = float(res[:, 1].mean())
    out[name] = {"patterns": n, "kernel_ms": round(ms, 3), "patterns_per_s_kernel": round(n / ms * 1e3, 1),
                 "patterns_per_s_call": round(n / wall, 1), "mean_evaluations": round(evals, 1),
                 "pixel_evaluations_per_s": float(f"{n * evals * side * side / ms * 1e3:.4g}"),
                 "mean_score": round(float(res[:, 0].mean()), 6)}
    if mode == _lib.REFINE_ORI:
        t0 = time.time()
        want = ro.refine_orientation(p, pats[:n_cpu], x0[:n_cpu], False)
        cpu = time.time() - t0
        out[name]["cpu_port_patterns_per_s_1core"] = round(n_cpu / cpu, 2)
        out[name]["max_dscore_vs_port"] = float(np.abs(res[:n_cpu, 0] - want[:, 0]).max())
        out[name]["identical_searches"] = int(np.sum(res[:n_cpu, 1] == want[:, 1]))
print(json.dumps({"detector": [side, side], "master_pattern": mps, **out}))
