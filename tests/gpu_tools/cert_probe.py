"""Statistics of (tensor-core score - exact score) on the candidates, per metric / shape."""
import os, sys
import numpy as np
import torch
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
import kikuchipy_b200 as kb
from kikuchipy_b200 import _lib
from oracle import di_oracle as orc

ctx = kb.default_context(0)

def probe(M, N, sig, metric, mask, k=20, planted=True, bf16=False):
    g = torch.Generator(device="cuda"); g.manual_seed(11)
    dic = torch.rand((N,) + sig, device="cuda", generator=g)
    if planted:
        j = torch.randint(0, N, (M,), device="cuda", generator=g)
        noise = torch.rand((M,) + sig, device="cuda", generator=g)
        exp = torch.clamp(torch.round(255.0 * (0.7 * dic[j] + 0.3 * noise)), 0, 255).to(torch.uint8)
    else:
        exp = torch.randint(0, 256, (M,) + sig, dtype=torch.uint8, device="cuda", generator=g)
    ctx.set_signal_mask(orc.circular_signal_mask(sig) if mask else None)
    ctx.set_option(_lib.OPT_COMPUTE_DTYPE, 1 if bf16 else 0)
    code = _lib.KDI_NCC if metric == "ncc" else _lib.KDI_NDP
    shard, approx, gidx = ctx.shard_candidates(exp, M, dic, N, code, k)
    exact = shard.rescore_owned(gidx)
    idx, sc, flags = shard.finalize(approx, gidx, exact, k, N)
    shard.close()
    ctx.set_signal_mask(None); ctx.set_option(_lib.OPT_COMPUTE_DTYPE, 0)
    d = (exact - approx).double()
    bias = d.mean(1); std = d.std(1)
    srt = torch.sort(exact, dim=1, descending=True).values
    gap = (srt[:, k - 1] - approx[:, -1]).double()
    print(f"{metric} {sig} mask={mask} planted={planted} bf16={bf16} M={M} N={N} kc={approx.shape[1]}: "
          f"bias mean {bias.mean():.3e} (min {bias.min():.3e} max {bias.max():.3e}) | row std mean {std.mean():.3e} max {std.max():.3e} | "
          f"rms {d.pow(2).mean().sqrt():.3e} | gap E_k - t: median {gap.median():.3e} min {gap.min():.3e} | flagged {flags.numel()}/{M} "
          f"| score range {float(exact[:, 1:].min()):.4f}..{float(exact[:, 1:].max()):.4f}")

probe(2000, 100000, (60, 60), "ncc", False, planted=False)
probe(2000, 100000, (60, 60), "ncc", False)
probe(2000, 100000, (60, 60), "ndp", False)
probe(2000, 100000, (60, 60), "ndp", False, planted=False)
probe(2000, 100000, (120, 120), "ndp", True)
probe(2000, 100000, (120, 120), "ncc", True)
probe(2000, 100000, (60, 60), "ndp", False, bf16=True)
probe(2000, 100000, (60, 60), "ncc", False, bf16=True, planted=False)
probe(2000, 37500, (60, 60), "ncc", False, k=50, planted=False)
