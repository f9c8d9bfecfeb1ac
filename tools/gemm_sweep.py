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

# (superblock, strip_tiles, l2_policy, rotate)
SETTINGS = [(0, 0, 0, 0), (0, 0, 1, 0), (0, 0, 2, 0), (0, 0, 3, 0), (0, 0, 0, 1), (0, 0, 1, 1),
            (0, 4, 0, 0), (0, 4, 0, 1), (0, 16, 0, 1), (10, 0, 0, 0), (10, 4, 0, 1), (5, 4, 0, 0),
            (40, 0, 0, 0), (40, 4, 0, 0), (40, 4, 0, 1), (40, 4, 1, 1), (40, 9, 1, 1), (40, 16, 0, 0)]
if os.environ.get("SETTINGS"):
    SETTINGS = [tuple(int(x) for x in s.split(",")) for s in os.environ["SETTINGS"].split(";")]
ref = None
for sb, st, pol, rot in SETTINGS:
    ctx.set_option(_lib.OPT_SUPERBLOCK, sb); ctx.set_option(_lib.OPT_STRIP_TILES, st)
    ctx.set_option(_lib.OPT_L2_POLICY, pol); ctx.set_option(_lib.OPT_TILE_ROTATE, rot)
    ms, tot = [], []
    for _ in range(REPS):
        ctx.dictionary_indexing(exp, M, dic, N, _lib.KDI_NCC, 20, out=(idx, sc))
        t = ctx.timings(); ms.append(t["gemm_topk_ms"]); tot.append(t["total_ms"])
    if ref is None:
        ref = idx.clone()
    same = bool(torch.equal(ref, idx))
    print(json.dumps({"superblock": sb, "strip_tiles": st, "l2_policy": pol, "rotate": rot,
                      "gemm_ms": [round(x, 3) for x in ms], "total_ms": round(min(tot), 3),
                      "rescore_ms": round(t["rescore_ms"], 3), "flagged": t["flagged_rows"], "same_idx": same}), flush=True)
