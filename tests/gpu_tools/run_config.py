"""Run one BASELINE.json configuration at FULL size on the GPU(s), verify it, print one JSON line.

  python tests/gpu_tools/run_config.py --config 2|3                                   (one GPU)
  python -m torch.distributed.run --nproc-per-node 8 ... tests/gpu_tools/run_config.py --config 4|5

  2: 10 000 x 100 000, 60x60, NCC, keep_n 20            (1 GPU)
  3: 40 000 x 100 000, 120x120, circular mask, NDP      (1 GPU)
  4: 100 000 x 300 000, 60x60, NCC, keep_n 50           (dictionary sharded over the ranks)
  5: 200x200 map x 500 000, 80x80, bf16 candidates, host dictionary streamed, + OSM (sharded)

Inputs are synthetic and generated on the device: a uniform-random float32 dictionary (rank r
generates its shard with seed 100 + r) and "planted" uint8 patterns (pattern i is a noisy copy of
dictionary row j[i]; j is drawn globally so the planted rows live on all shards).

Verification (rank 0, after the timed steps):
  * planted row is the best match of every pattern; lists sorted, indices valid and unique
  * a sample of rows against an independent float64 evaluation of the whole dictionary
    (torch.matmul in float64, shard by shard): same index lists wherever float64 scores are
    separated by more than 1e-6, scores within 1e-5
  * a smaller sample against the CPU oracle (the reference's float32 NumPy arithmetic) on the
    full host copy of the dictionary: tie-tolerant index identity, scores within 1e-4
  * config 5: the orientation similarity map equals the oracle's on the returned indices
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

CONFIGS = {
    2: dict(M=10_000, nav=(100, 100), N=100_000, sig=(60, 60), metric="ncc", k=20, mask=False, bf16=False, host_dict=False),
    3: dict(M=40_000, nav=(200, 200), N=100_000, sig=(120, 120), metric="ndp", k=20, mask=True, bf16=False, host_dict=False),
    4: dict(M=100_000, nav=(250, 400), N=300_000, sig=(60, 60), metric="ncc", k=50, mask=False, bf16=False, host_dict=False),
    5: dict(M=40_000, nav=(200, 200), N=500_000, sig=(80, 80), metric="ncc", k=20, mask=False, bf16=True, host_dict=True),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, required=True, choices=sorted(CONFIGS))
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--sample64", type=int, default=256, help="rows checked against the float64 evaluation")
    ap.add_argument("--sample-oracle", type=int, default=16, help="rows checked against the CPU oracle")
    ap.add_argument("--scale", type=float, default=1.0, help="shrink M and N (smoke runs)")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    import kikuchipy_b200 as kb
    from kikuchipy_b200 import _lib
    from oracle import di_oracle as orc  # checker only

    cfg = dict(CONFIGS[args.config])
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    M, N, sig, k = int(cfg["M"] * args.scale), int(cfg["N"] * args.scale), cfg["sig"], cfg["k"]
    nav = cfg["nav"] if args.scale == 1.0 else (M,)
    S = sig[0] * sig[1]
    ctx = kb.default_context(local)
    if cfg["bf16"]:
        ctx.set_option(_lib.OPT_COMPUTE_DTYPE, 1)
    smask = orc.circular_signal_mask(sig) if cfg["mask"] else None
    start, end = kb.shard_bounds(N, world, rank)
    n_shard = end - start

    # ---- synthetic inputs ---------------------------------------------------------------------
    def shard_rows(r, rows=None):
        """rows (local indices; None = all) of rank r's shard, regenerated from its seed"""
        s0, s1 = kb.shard_bounds(N, world, r)
        g = torch.Generator(device=dev); g.manual_seed(100 + r)
        full = torch.rand((s1 - s0, S), dtype=torch.float32, device=dev, generator=g)
        return full if rows is None else full[rows]

    dic = shard_rows(rank)
    g = torch.Generator(device=dev); g.manual_seed(7)
    j = torch.randint(0, N, (M,), device=dev, generator=g)          # planted dictionary row per pattern
    noise_seed = 11
    exp = torch.zeros((M, S), dtype=torch.float32, device=dev)
    mine = (j >= start) & (j < end)
    exp[mine] = dic[(j[mine] - start)]
    if world > 1:
        dist.all_reduce(exp)                                        # every rank gets every planted row
    g.manual_seed(noise_seed)
    for a in range(0, M, 8192):
        b = min(a + 8192, M)
        nz = torch.rand((b - a, S), dtype=torch.float32, device=dev, generator=g)
        exp[a:b] = torch.clamp(torch.round(255.0 * (0.7 * exp[a:b] + 0.3 * nz)), 0, 255)
    exp = exp.to(torch.uint8).reshape((M,) + sig)
    dic = dic.reshape((n_shard,) + sig)
    dict_in = dic
    if cfg["host_dict"]:
        dict_in = ctx.pinned_empty((n_shard,) + sig, np.float32)
        dict_in[...] = dic.cpu().numpy()
    torch.cuda.synchronize()

    # ---- timed steps ----------------------------------------------------------------------------
    def step():
        if world == 1:
            idx = torch.empty((M, k), dtype=torch.int64, device=dev)
            sc = torch.empty((M, k), dtype=torch.float32, device=dev)
            ctx.set_signal_mask(smask)
            ctx.dictionary_indexing(exp, M, dict_in, n_shard, _lib.KDI_NCC if cfg["metric"] == "ncc" else _lib.KDI_NDP,
                                    k, out=(idx, sc))
            return idx, sc
        return kb.dictionary_indexing_sharded(exp, dict_in, N, metric=cfg["metric"], keep_n=k, signal_mask=smask,
                                              context=ctx)

    for _ in range(args.warmup):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    per_step = []
    for _ in range(args.steps):
        t_s = time.perf_counter()
        idx, sc = step()
        torch.cuda.synchronize()
        per_step.append(round((time.perf_counter() - t_s) * 1e3, 3))
    if world > 1:
        dist.barrier()
    ms = (time.perf_counter() - t0) * 1e3 / args.steps
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    tm = ctx.timings()
    s_eff = S if smask is None else int((~smask).sum())
    # (host-streamed dictionaries run the tensor-core pass in several launches under the upload;
    # the per-call timing then covers the last launch only)
    gemm_tflops = None
    if tm["gemm_topk_ms"] > 0 and not cfg["host_dict"]:
        gemm_tflops = 2.0 * M * n_shard * s_eff / (tm["gemm_topk_ms"] * 1e-3) / 1e12

    # ---- verification ---------------------------------------------------------------------------
    checks = {}
    checks["planted_best"] = bool(torch.equal(idx[:, 0], j))
    checks["sorted"] = bool((sc[:, :-1] >= sc[:, 1:]).all())
    checks["indices_valid"] = bool(int(idx.min()) >= 0 and int(idx.max()) < N)
    srt = torch.sort(idx, dim=1).values
    checks["indices_unique"] = bool((srt[:, 1:] != srt[:, :-1]).all())
    # float64 evaluation of sample rows against the whole dictionary (every rank can do it alone;
    # rank 0 reports)
    if rank == 0:
        rows = torch.linspace(0, M - 1, min(args.sample64, M), device=dev).long().unique()
        keep = torch.ones(S, dtype=torch.bool, device=dev) if smask is None else torch.from_numpy(~smask.ravel()).to(dev)
        e = exp.reshape(M, S)[rows].double()[:, keep]
        if cfg["metric"] == "ncc":
            e = e - e.mean(1, keepdim=True)
        e = e / e.norm(dim=1, keepdim=True)
        best_s = torch.full((rows.numel(), 0), 0.0, dtype=torch.float64, device=dev)
        best_i = torch.zeros((rows.numel(), 0), dtype=torch.int64, device=dev)
        for r in range(world):
            s0, _ = kb.shard_bounds(N, world, r)
            d_all = shard_rows(r)
            for a in range(0, d_all.shape[0], 16384):
                d = d_all[a:a + 16384].double()[:, keep]
                if cfg["metric"] == "ncc":
                    d = d - d.mean(1, keepdim=True)
                d = d / d.norm(dim=1, keepdim=True)
                s = e @ d.T
                kk = min(k + 8, s.shape[1])
                ts, ti = torch.topk(s, kk, dim=1)
                best_s = torch.cat([best_s, ts], 1)
                best_i = torch.cat([best_i, ti + s0 + a], 1)
                o = torch.argsort(best_s, dim=1, descending=True, stable=True)[:, : k + 8]
                best_s, best_i = torch.gather(best_s, 1, o), torch.gather(best_i, 1, o)
            del d_all
        ref_s, ref_i = best_s.cpu().numpy(), best_i.cpu().numpy()
        got_s, got_i = sc[rows].cpu().numpy(), idx[rows].cpu().numpy()
        checks["f64_max_dscore"] = float(np.abs(ref_s[:, :k] - got_s).max())
        # index lists must agree wherever the float64 scores are separated by more than 1e-6
        ok_rows, exact_rows = 0, 0
        for r in range(ref_i.shape[0]):
            exact_rows += int(np.array_equal(ref_i[r, :k], got_i[r]))
            good = True
            for p in range(k):
                if got_i[r, p] != ref_i[r, p]:
                    # acceptable only if the returned index has a float64 score within 1e-6 of the reference's p-th
                    where = np.nonzero(ref_i[r] == got_i[r, p])[0]
                    if where.size == 0 or abs(ref_s[r, where[0]] - ref_s[r, p]) > 1e-6:
                        good = False
                        break
            ok_rows += int(good)
        checks["f64_rows"] = int(ref_i.shape[0])
        checks["f64_rows_tie_ok"] = ok_rows
        checks["f64_rows_identical"] = exact_rows
    # CPU oracle (reference arithmetic, float32 NumPy) on a few rows against the full dictionary
    if args.sample_oracle > 0:
        if rank == 0:
            t1 = time.perf_counter()
            rows = np.linspace(0, M - 1, min(args.sample_oracle, M)).astype(np.int64)
            e_h = exp.reshape(M, S)[torch.from_numpy(rows).to(dev)].cpu().numpy().reshape((-1,) + sig)
            ridx = np.zeros((rows.size, k), np.int64)
            rsc = np.full((rows.size, k), -1.0, np.float32)
            for r in range(world):  # the reference's chunk loop, one shard = one chunk
                s0, _ = kb.shard_bounds(N, world, r)
                d_h = shard_rows(r).cpu().numpy().reshape((-1,) + sig)
                ci, cs = orc.dictionary_indexing(e_h, d_h, metric=cfg["metric"], keep_n=k, signal_mask=smask,
                                                 n_per_iteration=20_000)
                alls = np.hstack((rsc, cs)); alli = np.hstack((ridx, ci + s0))
                best = np.argsort(-alls, axis=1, kind="stable")[:, :k]
                rsc, ridx = np.take_along_axis(alls, best, 1), np.take_along_axis(alli, best, 1)
                del d_h
            got_i = idx[torch.from_numpy(rows).to(dev)].cpu().numpy()
            got_s = sc[torch.from_numpy(rows).to(dev)].cpu().numpy()
            c = orc.compare_topk(ridx, rsc, got_i, got_s, tie_tol=2e-5 if smask is not None else 1e-6)
            checks["oracle_rows"] = int(rows.size)
            checks["oracle_tie_ok"] = bool(c["tie_ok"])
            checks["oracle_scores_ok"] = bool(c["scores_ok"])
            checks["oracle_max_dscore"] = float(c["max_dscore"])
            checks["oracle_exact_rows"] = float(c["exact_rows"])
            checks["oracle_s"] = round(time.perf_counter() - t1, 1)
    osm_info = None
    if args.config == 5 and rank == 0 and len(nav) == 2:
        t1 = time.perf_counter()
        idx_h = idx.cpu().numpy()
        osm = ctx.orientation_similarity_map(idx_h, nav[0], nav[1], k, k, False, np.array([[0, 1, 0], [1, 1, 1], [0, 1, 0]]), 2)[..., 0]
        osm_ms = (time.perf_counter() - t1) * 1e3
        t1 = time.perf_counter()
        ref = orc.orientation_similarity_map(idx_h, nav)
        osm_info = {"shape": list(osm.shape), "equal_to_oracle": bool(np.array_equal(osm, ref)), "gpu_ms_incl_copies": round(osm_ms, 2),
                    "oracle_s": round(time.perf_counter() - t1, 2), "mean": float(osm.mean())}
    if rank == 0:
        line = {
            "config": args.config, "n_gpus": world, "M": M, "N": N, "signal": list(sig), "s_eff": s_eff, "metric": cfg["metric"],
            "keep_n": k, "compute_dtype": "bf16" if cfg["bf16"] else "fp16", "dictionary": "host (streamed)" if cfg["host_dict"] else "device",
            "ms_per_step": round(ms, 3), "rank0_ms_each_step": per_step, "patterns_per_s": round(M / (ms * 1e-3)), "comparisons_per_s": float(M) * N / (ms * 1e-3),
            "rank0_stage_ms": {kk: round(v, 3) for kk, v in tm.items() if kk.endswith("_ms")},
            "rank0_gemm_tflops_algorithmic": None if gemm_tflops is None else round(gemm_tflops, 1),
            "flagged_rows": int(tm["flagged_rows"]), "checks": checks, "osm": osm_info,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
