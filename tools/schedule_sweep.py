"""Sweep the schedule options of the device-resident pipeline on BASELINE config 2 (10 000 x 100 000 x
3 600) - or any M / N / KEEP from the environment.  Settings are interleaved (A B C A B C ...) so clock
and power drift do not favour one of them; every setting must reproduce the first one's indices.
KDI_TIMELINE=1 prints the per-launch timeline of the last call of each setting."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import kikuchipy_b200 as kb
from kikuchipy_b200 import _lib

M, N = int(os.environ.get("M", "10000")), int(os.environ.get("N", "100000"))
SIG = (int(os.environ.get("SY", "60")), int(os.environ.get("SX", "60")))
KEEP = int(os.environ.get("KEEP", "20"))
REPS, ROUNDS = int(os.environ.get("REPS", "3")), int(os.environ.get("ROUNDS", "4"))
O = _lib
NAMES = {"flags": O.OPT_DEP_FLAGS, "groups": O.OPT_MIN_GROUPS, "post": O.OPT_POST_PER_GROUP, "sms": O.OPT_GEMM_SMS,
         "serial": O.OPT_GEMM_SERIAL, "part": O.OPT_SM_PARTITION, "cores": O.OPT_POST_CORESIDENT, "split": O.OPT_EARLY_SPLIT, "overlap": O.OPT_OVERLAP, "stages": O.OPT_MAX_STAGES,
         "strip": O.OPT_STRIP_TILES, "sb": O.OPT_SUPERBLOCK, "view": O.OPT_DICT_VIEW, "divd": O.OPT_DIV_DOUBLE, "dual": O.OPT_GEMM_DUAL,
         "cert": O.OPT_CERT_STRICT, "widen": O.OPT_CERT_WIDEN}
DEFAULTS = {"flags": 0, "groups": 0, "post": 0, "sms": 0, "serial": 0, "part": 0, "cores": 0, "split": 1, "overlap": 1, "stages": 0, "strip": 0, "sb": 0,
            "view": 1, "divd": 0, "dual": 0, "cert": 2, "widen": 0}
SETTINGS = os.environ.get(
    "SETTINGS",
    "split=1;split=0;split=0,groups=4;flags=1;overlap=0").split(";")

ctx = kb.default_context(0)
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(1)
exp = torch.randint(0, 256, (M,) + SIG, dtype=torch.uint8, device=dev, generator=g)
g.manual_seed(2)
dic = torch.rand((N,) + SIG, dtype=torch.float32, device=dev, generator=g)
idx = torch.empty((M, KEEP), dtype=torch.int64, device=dev)
sc = torch.empty((M, KEEP), dtype=torch.float32, device=dev)
torch.cuda.synchronize()


def apply(setting):
    vals = dict(DEFAULTS)
    for kv in setting.split(","):
        k, v = kv.split("=")
        vals[k] = int(v)
    for k in ("part",) + tuple(x for x in vals if x != "part"):  # partition first: it may be refused
        ctx.set_option(NAMES[k], vals[k])


ref = None
acc = {s: {"total": [], "gemm": [], "post": [], "wall": [], "same": True, "flagged": 0, "error": None} for s in SETTINGS}
for rnd in range(ROUNDS):
    for s in SETTINGS:
        a = acc[s]
        if a["error"]:
            continue
        try:
            apply(s)
        except (NotImplementedError, ValueError) as e:
            a["error"] = str(e)
            continue
        try:
          for r in range(REPS):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            st = torch.cuda.ExternalStream(ctx.stream_handle(), device=dev)
            e0.record(st)
            ctx.dictionary_indexing(exp, M, dic, N, _lib.KDI_NCC, KEEP, out=(idx, sc))
            e1.record(st)
            torch.cuda.synchronize()
            t = ctx.timings()
            if rnd > 0 or r > 0:  # first call of a setting: warm-up
                a["total"].append(t["total_ms"]); a["gemm"].append(t["gemm_topk_ms"]); a["post"].append(t["rescore_ms"])
                a["wall"].append(e0.elapsed_time(e1))
            a["flagged"] = t["flagged_rows"]
            a["model_rows"] = t.get("model_rows", -1)
        except _lib.KdiError as e:
            a["error"] = str(e)
            print(json.dumps({"setting": s, "round": rnd, "error": str(e)}), file=sys.stderr, flush=True)
            continue
        if ref is None:
            ref = idx.clone()
        a["same"] = a["same"] and bool(torch.equal(ref, idx))
for s, a in acc.items():
    if a["error"]:
        print(json.dumps({"setting": s, "error": a["error"]}), flush=True)
        continue
    print(json.dumps({"setting": s, "total_ms_mean": round(float(np.mean(a["total"])), 3),
                      "total_ms_min": round(min(a["total"]), 3), "event_span_ms_mean": round(float(np.mean(a["wall"])), 3),
                      "gemm_ms_mean": round(float(np.mean(a["gemm"])), 3), "post_tail_ms_mean": round(float(np.mean(a["post"])), 3),
                      "flagged": a["flagged"], "model_rows": a.get("model_rows"), "same_idx": a["same"], "n": len(a["total"])}), flush=True)
