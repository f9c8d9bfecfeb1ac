"""Sweep the L2-reuse knobs of the GEMM + top-k kernel on BASELINE config 2 (10 000 x 100 000 x
3 600, device-resident).  Plain run: CUDA-event time of the kernel per setting.  Under
`ncu --metrics dram__bytes_read.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum
-k regex:kdi_gemm_kernel`: DRAM traffic per launch (launch order = setting order x REPS)."""
import sys, os, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import kikuchipy_b200 as kb
from kikuchipy_b200 import _lib

REPS = int(os.environ.get("REPS", "4"))
M, N, SIG = int(os.environ.get("M", "10000")), int(os.environ.get("N", "100000")), (60, 60)
ctx = kb.default_context(0)
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(1)
exp = torch.randint(0, 256, (M,) + SIG, dtype=torch.uint8, device=dev, generator=g)
g.manual_seed(2)
dic = torch.rand((N,) + SIG, dtype=torch.float32, device=dev, generator=g)
idx = torch.empty((M, 20), dtype=torch.int64, device=dev)
sc = torch.empty((M, 20), dtype=torch.float32, device=dev)
torch.cuda.synchronize()

# (superblock, strip_tiles, l2_policy, rotate[, max_stages])
SETTINGS = [(0, 0, 0, 0), (0, 0, 1, 0), (0, 0, 2, 0), (0, 0, 3, 0), (0, 0, 0, 1), (0, 0, 1, 1),
            (0, 4, 0, 0), (0, 4, 0, 1), (0, 16, 0, 1), (10, 0, 0, 0), (10, 4, 0, 1), (5, 4, 0, 0),
            (40, 0, 0, 0), (40, 4, 0, 0), (40, 4, 0, 1), (40, 4, 1, 1), (40, 9, 1, 1), (40, 16, 0, 0)]
if os.environ.get("SETTINGS"):
    SETTINGS = [tuple(int(x) for x in s.split(",")) for s in os.environ["SETTINGS"].split(";")]
if os.environ.get("OVERLAP"):
    ctx.set_option(_lib.OPT_OVERLAP, int(os.environ["OVERLAP"]))
ROUNDS = int(os.environ.get("ROUNDS", "1"))  # > 1: settings interleaved (A B C A B C ...) so that clock / power drift
                                            # over the run does not favour whichever setting happens to run first
ref = None
acc = {s: {"gemm": [], "total": [], "rescore": [], "flagged": 0, "same": True} for s in SETTINGS}
for rnd in range(ROUNDS):
    for setting in SETTINGS:
        sb, st, pol, rot = setting[:4]
        ctx.set_option(_lib.OPT_MAX_STAGES, setting[4] if len(setting) > 4 else 0)
        ctx.set_option(_lib.OPT_SUPERBLOCK, sb); ctx.set_option(_lib.OPT_STRIP_TILES, st)
        ctx.set_option(_lib.OPT_L2_POLICY, pol); ctx.set_option(_lib.OPT_TILE_ROTATE, rot)
        a = acc[setting]
        for _ in range(REPS):
            ctx.dictionary_indexing(exp, M, dic, N, _lib.KDI_NCC, 20, out=(idx, sc))
            t = ctx.timings()
            a["gemm"].append(t["gemm_topk_ms"]); a["total"].append(t["total_ms"]); a["rescore"].append(t["rescore_ms"])
            a["flagged"] = t["flagged_rows"]
        if ref is None:
            ref = idx.clone()
        a["same"] = a["same"] and bool(torch.equal(ref, idx))
for setting, a in acc.items():
    sb, st, pol, rot = setting[:4]
    print(json.dumps({"superblock": sb, "strip_tiles": st, "l2_policy": pol, "rotate": rot,
                      "max_stages": setting[4] if len(setting) > 4 else 0,
                      "gemm_ms_mean": round(float(np.mean(a["gemm"])), 3), "gemm_ms_min": round(min(a["gemm"]), 3),
                      "total_ms_mean": round(float(np.mean(a["total"])), 3), "rescore_ms_mean": round(float(np.mean(a["rescore"])), 3),
                      "flagged": a["flagged"], "same_idx": a["same"], "n": len(a["gemm"])}), flush=True)
