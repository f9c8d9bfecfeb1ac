"""Time the projection kernel (raw output and fused with the prepare step) for 100 000 rotations."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import kikuchipy_b200 as kb
from kikuchipy_b200 import _lib
from kikuchipy_b200 import synthetic as po
ctx = kb.default_context(0)
N = int(os.environ.get("N", "100000"))
# CONFIGS="size:libm,..." (libm = 1: the CUDA math library version of the per-pixel arithmetic)
CONFIGS = [tuple(int(x) for x in c.split(":")) for c in os.environ.get("CONFIGS", "401:0,1001:1,1001:0").split(",")]
for size, libm in CONFIGS:
    ctx.set_option(_lib.OPT_PROJECT_LIBM, libm)
    mu, ml = po.synthetic_master_pattern(size, seed=5)
    dc = kb.direction_cosines([-0.9, 0.85, -0.7, 0.95], 0.5, 60, 60, po.tilted_detector_matrix(70.0))
    rot = torch.from_numpy(po.random_rotations(N, seed=4)).cuda()
    mp = ctx.master_pattern(mu, ml, dc)
    out = torch.empty((N, 3600), dtype=torch.float32, device="cuda")
    for name, fn in (("raw float32 out", lambda: ctx.project_patterns(mp, rot, out=out)),
                     ("fused with prepare", lambda: ctx.patterns_projected(mp, rot, _lib.KDI_NCC).close())):
        fn(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / 5 * 1e3
        print(f"master {size}x{size}, {'CUDA math library' if libm else 'own sequences'}: {name}: {ms:.3f} ms for {N} patterns ({N * 3600 / ms / 1e6:.1f} Gpixel/s)")
    mp.close()
