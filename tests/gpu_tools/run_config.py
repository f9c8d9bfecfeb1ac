"""Run one BASELINE.json configuration at FULL size on the GPU(s), verify it, print one JSON line.

  python tests/gpu_tools/run_config.py --config 2|3                                   (one GPU)
  python -m torch.distributed.run --nproc-per-node 8 ... tests/gpu_tools/run_config.py --config 4|5|50
  (50 = config 5 with the dictionary generated on the device from rotations of a master pattern)

The run itself and the on-device verification (structure, planted best match, float64 evaluation of
a row sample) live in ``tools/di_configs.py``, which ``bench.py`` uses too; this script adds a
smaller sample against the CPU oracle (the reference's float32 NumPy arithmetic) on a host copy of
the dictionary - tie-tolerant index identity, scores within 1e-4 - and, for config 5, the
orientation similarity map against the oracle's.
"""

from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, required=True, choices=[2, 3, 4, 5, 50])
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--sample64", type=int, default=256, help="rows checked against the float64 evaluation")
    ap.add_argument("--sample-oracle", type=int, default=16, help="rows checked against the CPU oracle")
    ap.add_argument("--scale", type=float, default=1.0, help="shrink M and N (smoke runs)")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    import kikuchipy_b200 as kb
    from oracle import di_oracle as orc  # checker only
    from tools import di_configs as dc

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = kb.default_context(local)
    line = dc.run_config(args.config, ctx, rank, world, dev, steps=args.steps, warmup=args.warmup,
                         sample64=args.sample64, scale=args.scale, keep_result=True)
    idx, sc, exp, dictionary, smask = line.pop("_result")
    cfg = dc.CONFIGS[args.config]
    k, sig, M = cfg["k"], cfg["sig"], line["M"]
    if rank == 0 and args.sample_oracle > 0:
        t1 = time.perf_counter()
        rows = np.linspace(0, M - 1, min(args.sample_oracle, M)).astype(np.int64)
        sel = torch.from_numpy(rows).to(dev)
        e_h = exp.reshape(M, -1)[sel].cpu().numpy().reshape((-1,) + sig)
        ridx = np.zeros((rows.size, k), np.int64)
        rsc = np.full((rows.size, k), -1.0, np.float32)
        for r in range(world):  # the reference's chunk loop, one shard = one chunk
            s0, _ = dictionary.bounds[r]
            d_h = dictionary.shard(r).cpu().numpy().reshape((-1,) + sig)
            ci, cs = orc.dictionary_indexing(e_h, d_h, metric=cfg["metric"], keep_n=k, signal_mask=smask,
                                             n_per_iteration=20_000)
            alls = np.hstack((rsc, cs)); alli = np.hstack((ridx, ci + s0))
            best = np.argsort(-alls, axis=1, kind="stable")[:, :k]
            rsc, ridx = np.take_along_axis(alls, best, 1), np.take_along_axis(alli, best, 1)
            del d_h
        c = orc.compare_topk(ridx, rsc, idx[sel].cpu().numpy(), sc[sel].cpu().numpy(),
                             tie_tol=2e-5 if smask is not None else 1e-6)
        line["checks"].update({"oracle_rows": int(rows.size), "oracle_tie_ok": bool(c["tie_ok"]),
                               "oracle_scores_ok": bool(c["scores_ok"]), "oracle_max_dscore": float(c["max_dscore"]),
                               "oracle_exact_rows": float(c["exact_rows"]), "oracle_s": round(time.perf_counter() - t1, 1)})
    if rank == 0 and args.config in (5, 50) and line.get("osm"):
        nav = cfg["nav"]
        idx_h = idx.cpu().numpy()
        osm = ctx.orientation_similarity_map(idx_h, nav[0], nav[1], k, k, False, np.array([[0, 1, 0], [1, 1, 1], [0, 1, 0]]), 2)[..., 0]
        line["osm"]["equal_to_oracle"] = bool(np.array_equal(osm, orc.orientation_similarity_map(idx_h, nav)))
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
