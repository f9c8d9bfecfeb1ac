"""Full-size runs of the BASELINE.json configurations with on-device verification (no oracle, no
reference: everything here runs on the GPU box).  Used by ``bench.py`` (parity block of the timed
objects, the ``extra`` configuration lines) and by ``tests/gpu_tools/run_config.py`` (which adds a
CPU-oracle sample on top).

  2: 10 000 x 100 000, 60x60, NCC, keep_n 20            (1 GPU)
  3: 40 000 x 100 000, 120x120, circular mask, NDP      (1 GPU)
  4: 100 000 x 300 000, 60x60, NCC, keep_n 50           (dictionary sharded over the ranks)
  5: 200x200 map x 500 000, 80x80, bf16 candidates, host dictionary streamed, + OSM (sharded)
 50: config 5 with the dictionary GENERATED on the device, rank by rank, from 500 000 rotations of a master
     pattern - what the reference's lazy dictionary does chunk by chunk on the CPU
     (``dictionary_chunk.compute()``, _dictionary_indexing.py:105-108) - instead of 12.8 GB of float32
     patterns crossing PCIe every step

Verification of a result ``(indices, scores)``:
  * structure: lists sorted best first, indices valid and unique per row;
  * planted inputs (pattern i = noisy copy of dictionary row j[i]): j[i] must be the best match;
  * a sample of rows against an independent FLOAT64 evaluation of the whole dictionary
    (``torch.matmul`` in float64, shard by shard, regenerated from the shard seeds): identical
    index lists wherever the float64 scores are separated by more than 1e-6, scores within 1e-5.
"""

from __future__ import annotations

import time

import numpy as np

CONFIGS = {
    2: dict(M=10_000, nav=(100, 100), N=100_000, sig=(60, 60), metric="ncc", k=20, mask=False, bf16=False, host_dict=False),
    3: dict(M=40_000, nav=(200, 200), N=100_000, sig=(120, 120), metric="ndp", k=20, mask=True, bf16=False, host_dict=False),
    4: dict(M=100_000, nav=(250, 400), N=300_000, sig=(60, 60), metric="ncc", k=50, mask=False, bf16=False, host_dict=False),
    5: dict(M=40_000, nav=(200, 200), N=500_000, sig=(80, 80), metric="ncc", k=20, mask=False, bf16=True, host_dict=True),
    50: dict(M=40_000, nav=(200, 200), N=500_000, sig=(80, 80), metric="ncc", k=20, mask=False, bf16=True, host_dict=False,
             generated=True),
}


def circular_mask(sig) -> np.ndarray:
    """True = excluded; pixels farther than max(sy//2, sx//2) from (sy//2, sx//2)."""
    sy, sx = sig
    r, c = np.mgrid[:sy, :sx]
    return np.sqrt((r - sy // 2) ** 2 + (c - sx // 2) ** 2) > max(sy // 2, sx // 2)


class ShardedDictionary:
    """Uniform-random float32 dictionary of N rows, split into ``world`` contiguous shards; shard r is
    ``torch.rand`` with seed ``seed + r`` on the device, so any rank can regenerate any shard."""

    def __init__(self, N, S, world, dev, seed=100, shard_bounds=None):
        self.N, self.S, self.world, self.dev, self.seed = int(N), int(S), int(world), dev, int(seed)
        if shard_bounds is None:
            from kikuchipy_b200 import shard_bounds
        self.bounds = [shard_bounds(self.N, self.world, r) for r in range(self.world)]

    def shard(self, r):
        import torch

        s0, s1 = self.bounds[r]
        g = torch.Generator(device=self.dev); g.manual_seed(self.seed + r)
        return torch.rand((s1 - s0, self.S), dtype=torch.float32, device=self.dev, generator=g)


class GeneratedShardedDictionary:
    """The same interface for a dictionary projected from a master pattern: N random rotations (seed 4) of a
    smooth synthetic two-hemisphere master pattern (seed 5) seen by a ``sig`` detector; shard r is projected
    on the device by ``kdi_project_patterns`` whenever it is asked for, so any rank can regenerate any shard.
    ``generated(rank)`` is the lazy dictionary (rotations + master pattern) the indexing call is given."""

    def __init__(self, N, sig, world, dev, ctx, shard_bounds=None, master_pattern_size=1001):
        import kikuchipy_b200 as kb
        from kikuchipy_b200 import synthetic as po

        self.N, self.sig, self.S, self.world, self.dev, self.ctx = int(N), tuple(sig), int(sig[0] * sig[1]), int(world), dev, ctx
        if shard_bounds is None:
            shard_bounds = kb.shard_bounds
        self.bounds = [shard_bounds(self.N, self.world, r) for r in range(self.world)]
        self.mu, self.ml = po.synthetic_master_pattern(master_pattern_size, seed=5)
        self.dc = kb.direction_cosines([-0.9, 0.85, -0.7, 0.95], 0.5, sig[0], sig[1], po.tilted_detector_matrix(70.0))
        self.rotations = po.random_rotations(self.N, seed=4)
        self._mp = ctx.master_pattern(self.mu, self.ml, self.dc)

    def shard(self, r):
        import torch

        s0, s1 = self.bounds[r]
        out = torch.empty((s1 - s0, self.S), dtype=torch.float32, device=self.dev)
        self.ctx.project_patterns(self._mp, torch.from_numpy(self.rotations[s0:s1]).to(self.dev), out=out)
        return out

    def generated(self, rank):
        import kikuchipy_b200 as kb

        s0, s1 = self.bounds[rank]
        return kb.get_patterns(self.mu, self.ml, self.rotations[s0:s1], direction_cosines=self.dc,
                               detector_shape=self.sig, context=self.ctx)

    def close(self):
        self._mp.close()


def planted_patterns(dictionary: ShardedDictionary, my_shard, rank, M, seed=7, noise_seed=11):
    """uint8 patterns (M, S): pattern i = clip(round(255 (0.7 dict[j[i]] + 0.3 noise))), j drawn globally
    (the planted rows live on all shards; an all-reduce assembles them).  Returns (patterns, j)."""
    import torch
    import torch.distributed as dist

    dev, S = dictionary.dev, dictionary.S
    g = torch.Generator(device=dev); g.manual_seed(seed)
    j = torch.randint(0, dictionary.N, (M,), device=dev, generator=g)
    start, end = dictionary.bounds[rank]
    exp = torch.zeros((M, S), dtype=torch.float32, device=dev)
    mine = (j >= start) & (j < end)
    exp[mine] = my_shard.reshape(-1, S)[(j[mine] - start)]
    if dictionary.world > 1:
        dist.all_reduce(exp)
    g.manual_seed(noise_seed)
    for a in range(0, M, 8192):  # bounded temporaries
        b = min(a + 8192, M)
        nz = torch.rand((b - a, S), dtype=torch.float32, device=dev, generator=g)
        exp[a:b] = torch.clamp(torch.round(255.0 * (0.7 * exp[a:b] + 0.3 * nz)), 0, 255)
    return exp.to(torch.uint8), j


def structural_checks(idx, sc, N, planted_j=None) -> dict:
    import torch

    out = {
        "sorted": bool((sc[:, :-1] >= sc[:, 1:]).all()),
        "indices_valid": bool(int(idx.min()) >= 0 and int(idx.max()) < N),
    }
    srt = torch.sort(idx, dim=1).values
    out["indices_unique"] = bool((srt[:, 1:] != srt[:, :-1]).all())
    if planted_j is not None:
        out["planted_hit_rate"] = float((idx[:, 0] == planted_j).double().mean())
    return out


def float64_check(exp_u8, rows, dictionary: ShardedDictionary, metric, k, keep_cols, got_idx, got_sc, margin=8) -> dict:
    """``rows`` of the result against a float64 evaluation of the whole dictionary."""
    import torch

    dev = dictionary.dev
    S = dictionary.S
    keep = torch.ones(S, dtype=torch.bool, device=dev) if keep_cols is None else keep_cols
    e = exp_u8.reshape(-1, S)[rows].double()[:, keep]
    if metric == "ncc":
        e = e - e.mean(1, keepdim=True)
    e = e / e.norm(dim=1, keepdim=True)
    best_s = torch.zeros((rows.numel(), 0), dtype=torch.float64, device=dev)
    best_i = torch.zeros((rows.numel(), 0), dtype=torch.int64, device=dev)
    for r in range(dictionary.world):
        s0, _ = dictionary.bounds[r]
        d_all = dictionary.shard(r)
        for a in range(0, d_all.shape[0], 16384):
            d = d_all[a:a + 16384].double()[:, keep]
            if metric == "ncc":
                d = d - d.mean(1, keepdim=True)
            d = d / d.norm(dim=1, keepdim=True)
            s = e @ d.T
            ts, ti = torch.topk(s, min(k + margin, s.shape[1]), dim=1)
            best_s = torch.cat([best_s, ts], 1)
            best_i = torch.cat([best_i, ti + s0 + a], 1)
            o = torch.argsort(best_s, dim=1, descending=True, stable=True)[:, : k + margin]
            best_s, best_i = torch.gather(best_s, 1, o), torch.gather(best_i, 1, o)
        del d_all
    ref_s, ref_i = best_s.cpu().numpy(), best_i.cpu().numpy()
    got_s, got_i = got_sc[rows].cpu().numpy(), got_idx[rows].cpu().numpy()
    ok_rows = exact_rows = 0
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
    return {"f64_rows": int(ref_i.shape[0]), "f64_rows_identical": exact_rows, "f64_rows_tie_ok": ok_rows,
            "f64_max_dscore": float(np.abs(ref_s[:, :k] - got_s).max())}


def run_config(number, ctx, rank, world, dev, steps=3, warmup=1, sample64=256, scale=1.0, keep_result=False) -> dict:
    """Time and verify BASELINE configuration ``number`` on planted inputs.  Collective over the
    process group when world > 1.  Returns a dict (meaningful on rank 0; every rank must call)."""
    import torch
    import torch.distributed as dist

    import kikuchipy_b200 as kb
    from kikuchipy_b200 import _lib

    cfg = dict(CONFIGS[number])
    M, N, sig, k = int(cfg["M"] * scale), int(cfg["N"] * scale), cfg["sig"], cfg["k"]
    nav = cfg["nav"] if scale == 1.0 else (M,)
    S = sig[0] * sig[1]
    ctx.set_option(_lib.OPT_COMPUTE_DTYPE, 1 if cfg["bf16"] else 0)
    smask = circular_mask(sig) if cfg["mask"] else None
    generated = bool(cfg.get("generated"))
    if generated:
        dictionary = GeneratedShardedDictionary(N, sig, world, dev, ctx, shard_bounds=kb.shard_bounds)
    else:
        dictionary = ShardedDictionary(N, S, world, dev, shard_bounds=kb.shard_bounds)
    start, end = dictionary.bounds[rank]
    n_shard = end - start
    dic = dictionary.shard(rank)
    exp, j = planted_patterns(dictionary, dic, rank, M)
    exp = exp.reshape((M,) + sig)
    dic = dic.reshape((n_shard,) + sig)
    dict_in, pinned = dic, None
    if cfg["host_dict"]:
        pinned = ctx.pinned_empty((n_shard,) + sig, np.float32)
        pinned[...] = dic.cpu().numpy()
        dict_in = pinned
    gen = None
    if generated:
        gen = dictionary.generated(rank)
        del dic, dict_in  # (the materialised shard was only needed for the planted patterns)
        dic = dict_in = None
    torch.cuda.synchronize()
    code = _lib.KDI_NCC if cfg["metric"] == "ncc" else _lib.KDI_NDP

    def step():
        if world == 1:
            idx = torch.empty((M, k), dtype=torch.int64, device=dev)
            sc = torch.empty((M, k), dtype=torch.float32, device=dev)
            ctx.set_signal_mask(smask)
            if generated:
                ctx.dictionary_indexing_projected(exp, M, gen.master_pattern, gen.rotations, code, k, out=(idx, sc))
            else:
                ctx.dictionary_indexing(exp, M, dict_in, n_shard, code, k, out=(idx, sc))
            return idx, sc
        return kb.dictionary_indexing_sharded(exp, gen if generated else dict_in, N, metric=cfg["metric"], keep_n=k,
                                              signal_mask=smask, context=ctx)

    try:
        for _ in range(warmup):
            step()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            idx, sc = step()
            torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = (time.perf_counter() - t0) * 1e3 / steps
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        tm = ctx.timings()
    finally:
        ctx.set_signal_mask(None)
        ctx.set_option(_lib.OPT_COMPUTE_DTYPE, 0)
        if pinned is not None:
            ctx.pinned_free(pinned)
    s_eff = S if smask is None else int((~smask).sum())
    gemm_tflops = None
    if tm["gemm_topk_ms"] > 0 and not cfg["host_dict"]:
        gemm_tflops = 2.0 * M * n_shard * s_eff / (tm["gemm_topk_ms"] * 1e-3) / 1e12
    checks = structural_checks(idx, sc, N, j)
    if rank == 0 and sample64 > 0:
        rows = torch.linspace(0, M - 1, min(sample64, M), device=dev).long().unique()
        keep = None if smask is None else torch.from_numpy(~smask.ravel()).to(dev)
        checks.update(float64_check(exp, rows, dictionary, cfg["metric"], k, keep, idx, sc))
    osm_info = None
    if number in (5, 50) and rank == 0 and len(nav) == 2:
        t1 = time.perf_counter()
        osm = ctx.orientation_similarity_map(idx.cpu().numpy(), nav[0], nav[1], k, k, False,
                                             np.array([[0, 1, 0], [1, 1, 1], [0, 1, 0]]), 2)[..., 0]
        osm_info = {"shape": list(osm.shape), "gpu_ms_incl_copies": round((time.perf_counter() - t1) * 1e3, 2),
                    "mean": float(osm.mean()), "max": float(osm.max())}
    out = {
        "config": number, "n_gpus": world, "M": M, "N": N, "signal": list(sig), "s_eff": s_eff, "metric": cfg["metric"],
        "keep_n": k, "compute_dtype": "bf16" if cfg["bf16"] else "fp16",
        "dictionary": ("generated on the device from rotations of a 1001x1001 master pattern" if generated
                       else "pinned host, streamed" if cfg["host_dict"] else "device-resident"),
        "ms_per_step": round(ms, 3), "patterns_per_s": round(M / (ms * 1e-3)), "steps": steps, "warmup": warmup,
        "rank0_stage_ms": {kk: round(v, 3) for kk, v in tm.items() if kk.endswith("_ms")},
        "rank0_gemm_tflops_algorithmic": None if gemm_tflops is None else round(gemm_tflops, 1),
        "flagged_rows_rank0": int(tm["flagged_rows"]), "checks": checks, "osm": osm_info,
    }
    if keep_result:
        out["_result"] = (idx, sc, exp, dictionary, smask)
    elif generated:
        dictionary.close()
    return out
