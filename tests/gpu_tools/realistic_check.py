"""Realism check of the candidate certificate: a dictionary projected from a (synthetic, smooth)
master pattern, experimental patterns = projections at slightly perturbed dictionary rotations
+ noise, 8-bit.  Neighbouring dictionary entries are strongly correlated here (unlike uniform
random dictionaries), which is what real EBSD dictionaries look like.  Reports the rows the
certificate sends to the exact path and compares a sample against the forced-exact path."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import kikuchipy_b200 as kb
from kikuchipy_b200 import _lib
from oracle import projection_oracle as po

M, N = int(os.environ.get("M", "10000")), int(os.environ.get("N", "100000"))
SIG = (60, 60)
ctx = kb.default_context(0)
mu, ml = po.synthetic_master_pattern(1001, seed=5)
dc = kb.direction_cosines([-0.9, 0.85, -0.7, 0.95], 0.5, SIG[0], SIG[1], po.tilted_detector_matrix(70.0))
rot = po.random_rotations(N, seed=4)
mp = ctx.master_pattern(mu, ml, dc)
rng = np.random.default_rng(11)
j = rng.integers(0, N, M)
# perturb by ~1 degree: q' = normalise(q + 0.01 * gaussian)
q = rot[j] + 0.01 * rng.normal(size=(M, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
pat = torch.empty((M, 3600), dtype=torch.float32, device="cuda")
ctx.project_patterns(mp, torch.from_numpy(q).cuda(), out=pat)
for noise in (0.05, 0.3):
    g = torch.Generator(device="cuda"); g.manual_seed(3)
    lo, hi = pat.min(1, keepdim=True).values, pat.max(1, keepdim=True).values
    p = (pat - lo) / (hi - lo)
    exp = torch.clamp(torch.round(255 * ((1 - noise) * p + noise * torch.rand(p.shape, device="cuda", generator=g))), 0, 255).to(torch.uint8)
    for dtype_opt, name in ((0, "fp16"), (1, "bf16")):
        ctx.set_option(_lib.OPT_COMPUTE_DTYPE, dtype_opt)
        idx = torch.empty((M, 20), dtype=torch.int64, device="cuda"); sc = torch.empty((M, 20), dtype=torch.float32, device="cuda")
        for _ in range(2):
            ctx.dictionary_indexing_projected(exp, M, mp, rot, _lib.KDI_NCC, 20, out=(idx, sc))
        tm = ctx.timings()
        # forced-exact on a sample
        rows = torch.arange(0, M, max(1, M // 128), device="cuda")
        ctx.set_option(_lib.OPT_FORCE_EXACT, 1)
        i2 = torch.empty((rows.numel(), 20), dtype=torch.int64, device="cuda"); s2 = torch.empty((rows.numel(), 20), dtype=torch.float32, device="cuda")
        ctx.dictionary_indexing_projected(exp[rows].contiguous(), rows.numel(), mp, rot, _lib.KDI_NCC, 20, out=(i2, s2))
        ctx.set_option(_lib.OPT_FORCE_EXACT, 0)
        same = bool(torch.equal(i2, idx[rows])) and bool(torch.equal(s2, sc[rows]))
        gap = (sc[:, 0] - sc[:, 19]).median().item()
        print(json.dumps({"noise": noise, "candidates": name, "flagged_rows": int(tm["flagged_rows"]), "total_ms": round(tm["total_ms"], 3),
                          "fallback_ms": round(tm["fallback_ms"], 3), "rescore_ms": round(tm["rescore_ms"], 3),
                          "best_is_planted_neighbour": float((idx[:, 0] == torch.from_numpy(j).cuda()).float().mean()),
                          "median_score_best": round(sc[:, 0].median().item(), 4), "median_gap_1_to_20": round(gap, 5),
                          "sample_equals_forced_exact": same}), flush=True)
ctx.set_option(_lib.OPT_COMPUTE_DTYPE, 0)
