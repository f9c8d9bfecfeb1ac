"""Timing of the preprocessing kernels (run on the GPU box): static + dynamic background removal
in one launch and neighbour averaging on a 200 x 200 map of 60 x 60 uint8 patterns, device-resident,
with the oracle (the SciPy/NumPy port of the reference) timed beside it on a few patterns."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))
import kikuchipy_b200 as kb  # noqa: E402
from oracle import preprocess_oracle as pp  # noqa: E402

ny = nx = int(sys.argv[1]) if len(sys.argv) > 1 else 200
side = int(sys.argv[2]) if len(sys.argv) > 2 else 60
rng = np.random.default_rng(0)
pats = rng.integers(0, 256, (ny, nx, side, side), dtype=np.uint8)
bg = pats[:4, :4].mean(axis=(0, 1)).astype(np.uint8)
dev = torch.from_numpy(pats).cuda()
ctx = kb.default_context()
out = {"map": [ny, nx], "detector": [side, side], "bytes_in": int(pats.nbytes)}
for name, fn in (
    ("static+dynamic(frequency)", lambda: kb.preprocess(dev, static_bg=bg)),
    ("static+dynamic(spatial)", lambda: kb.preprocess(dev, static_bg=bg, filter_domain="spatial")),
    ("static only", lambda: kb.remove_static_background(dev, "subtract", bg)),
    ("average 3x3 circular", lambda: kb.average_neighbour_patterns(dev)),
):
    fn()
    torch.cuda.synchronize()
    fn()
    ms = ctx.timings()["total_ms"]
    out[name] = {"kernel_ms": round(ms, 3), "patterns_per_s": round(ny * nx / ms * 1e3), "GBps_algorithmic": round(2 * pats.nbytes / ms / 1e6, 1)}
n_cpu = 200
t0 = time.time()
pp.remove_dynamic_background(pp.remove_static_background(pats.reshape(-1, side, side)[:n_cpu], bg))
out["cpu_port_static+dynamic_patterns_per_s_1core"] = round(n_cpu / (time.time() - t0), 1)
print(json.dumps(out))
