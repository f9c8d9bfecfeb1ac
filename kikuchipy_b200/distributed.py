"""Dictionary sharding across GPUs: one process per GPU, one all-gather of per-shard top-k.

The reference has no multi-device code; its serial loop over dictionary chunks with a running
top-k (/root/reference/src/kikuchipy/indexing/_dictionary_indexing.py:94-128) is the same
reduction this module spreads over ranks: every rank matches ALL experimental patterns against
its contiguous dictionary shard (global indices via ``index_offset``, the ``+= start`` of
``:118``), the per-shard ``(M, keep_n)`` scores + indices are all-gathered (NCCL over NVLink on
GPUs, gloo in the CPU tests) and every rank merges the ``world_size`` lists with the same merge
kernel that serves the chunk loop.
"""

from __future__ import annotations

import os
import time

import numpy as np


class _Trace:
    """Per-phase wall times of the sharded pipeline (KDI_TRACE=1, rank 0)."""

    def __init__(self, on: bool):
        self.on, self.t, self.marks = on, None, []
        if on:
            import torch

            torch.cuda.synchronize()
            self.t = time.perf_counter()

    def __call__(self, name: str):
        if self.on:
            import torch

            torch.cuda.synchronize()
            now = time.perf_counter()
            self.marks.append((name, (now - self.t) * 1e3))
            self.t = now

    def report(self, ctx):
        if self.on:
            tm = ctx.timings()
            inner = {k: round(tm[k], 3) for k in ("normalize_exp_ms", "normalize_dict_ms", "gemm_topk_ms", "rescore_ms")}
            print("[kdi trace] " + " ".join(f"{n}={v:.3f}ms" for n, v in self.marks) + f" | inside candidates: {inner}",
                  flush=True)


def shard_bounds(n: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous, balanced ``[start, end)`` of shard ``rank`` (first ``n % world_size`` shards
    get one extra row)."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError("bad rank / world size")
    base, extra = divmod(int(n), world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def gather_topk(scores, indices, group=None):
    """All-gather per-rank ``(M, k)`` score / index tensors into ``(world, M, k)`` tensors
    (list-major: what ``kdi_merge_topk`` expects).  Works on CUDA tensors (NCCL) and CPU
    tensors (gloo)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    s_all = torch.empty((world,) + tuple(scores.shape), dtype=scores.dtype, device=scores.device)
    i_all = torch.empty((world,) + tuple(indices.shape), dtype=indices.dtype, device=indices.device)
    # contiguous slices of one buffer: NCCL takes the flat all-gather path, gloo gathers per slice
    dist.all_gather(list(s_all.unbind(0)), scores.contiguous(), group=group)
    dist.all_gather(list(i_all.unbind(0)), indices.contiguous(), group=group)
    return s_all, i_all


def dictionary_indexing_sharded(
    experimental,
    dictionary_shard,
    dictionary_size: int,
    metric: str = "ncc",
    keep_n: int = 20,
    navigation_mask: np.ndarray | None = None,
    signal_mask: np.ndarray | None = None,
    *,
    context=None,
    group=None,
):
    """Index ``experimental`` against a dictionary whose rows ``shard_bounds(dictionary_size,
    world, rank)`` this rank holds in ``dictionary_shard``.

    Returns ``(simulation_indices, scores)`` as CUDA tensors ``(M, keep_n)`` (global dictionary
    indices, identical on every rank).  Needs an initialised ``torch.distributed`` process group
    with the NCCL backend.
    """
    import torch
    import torch.distributed as dist

    from . import _lib

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    start, end = shard_bounds(dictionary_size, world, rank)
    n_shard = end - start
    if int(dictionary_shard.shape[0]) != n_shard:
        raise ValueError(f"rank {rank} expects {n_shard} dictionary rows, got {dictionary_shard.shape[0]}")
    ctx = context if context is not None else _lib.default_context()
    code = {"ncc": _lib.KDI_NCC, "ndp": _lib.KDI_NDP}[metric]
    ctx.set_signal_mask(signal_mask)
    nav_shape = tuple(experimental.shape[:-2])
    n_exp_all = int(np.prod(nav_shape)) if nav_shape else 1
    kept = n_exp_all if navigation_mask is None else int((~navigation_mask).sum())
    keep_n = min(int(keep_n), int(dictionary_size))
    dev = torch.device("cuda", ctx.device)
    if world == 1:
        scores = torch.empty((kept, keep_n), dtype=torch.float32, device=dev)
        idx = torch.empty((kept, keep_n), dtype=torch.int64, device=dev)
        ctx.dictionary_indexing(experimental, n_exp_all, dictionary_shard, n_shard, code, keep_n,
                                nav_mask=navigation_mask, index_offset=start, out=(idx, scores))
        return idx, scores
    if ctx.candidate_capacity(keep_n) == 0:
        return _sharded_exact_lists(ctx, experimental, n_exp_all, dictionary_shard, n_shard, code, keep_n,
                                    navigation_mask, start, kept, dictionary_size, group)

    trace = _Trace(os.environ.get("KDI_TRACE") == "1" and rank == 0)
    # 1. this shard's candidates by tensor-core score (global indices)
    shard, approx, gidx = ctx.shard_candidates(experimental, n_exp_all, dictionary_shard, n_shard, code, keep_n,
                                               nav_mask=navigation_mask, index_offset=start)
    trace("candidates")
    try:
        # 2. the one exchange of the path: all-gather of per-shard top-kc lists, merged per row
        s_all, i_all = gather_topk(approx, gidx, group)
        trace("all_gather")
        g_idx, g_approx = ctx.merge_topk(s_all, i_all, shard.kc)
        trace("merge")
        # 3. every rank rescores exactly the candidates whose dictionary rows it holds ...
        exact = shard.rescore_owned(g_idx)
        trace("rescore_owned")
        # 4. ... and the exact scores are combined (each candidate has exactly one owner)
        dist.all_reduce(exact, op=dist.ReduceOp.MAX, group=group)
        trace("all_reduce")
        # 5. rank by exact score + certificate (identical on every rank)
        idx, scores, flags = shard.finalize(g_approx, g_idx, exact, keep_n, dictionary_size)
        trace("finalize")
        trace.report(ctx)
        if flags.numel():
            # rows whose certificate failed: exact top-k per shard, gathered and merged
            k_local = min(keep_n, n_shard)
            fi, fs = shard.exact_rows(flags, k_local)
            if k_local != keep_n:
                fs = torch.nn.functional.pad(fs, (0, keep_n - k_local), value=-float("inf"))
                fi = torch.nn.functional.pad(fi, (0, keep_n - k_local), value=-1)
            fs_all, fi_all = gather_topk(fs, fi, group)
            mi, ms = ctx.merge_topk(fs_all, fi_all, keep_n)
            rows = flags.long()
            idx[rows] = mi
            scores[rows] = ms
        return idx, scores
    finally:
        shard.close()


def _sharded_exact_lists(ctx, experimental, n_exp_all, dictionary_shard, n_shard, code, keep_n,
                         navigation_mask, start, kept, dictionary_size, group):
    """keep_n beyond the candidate pipeline: per-shard final top-k lists, gathered and merged."""
    import torch

    dev = torch.device("cuda", ctx.device)
    k_local = min(int(keep_n), n_shard)
    scores = torch.empty((kept, k_local), dtype=torch.float32, device=dev)
    idx = torch.empty((kept, k_local), dtype=torch.int64, device=dev)
    ctx.dictionary_indexing(
        experimental, n_exp_all, dictionary_shard, n_shard, code, k_local,
        nav_mask=navigation_mask, index_offset=start, out=(idx, scores),
    )
    if k_local != keep_n:  # ragged shards: pad so every rank contributes the same shape
        pad_s = torch.full((kept, keep_n), -float("inf"), dtype=torch.float32, device=dev)
        pad_i = torch.full((kept, keep_n), -1, dtype=torch.int64, device=dev)
        pad_s[:, :k_local] = scores
        pad_i[:, :k_local] = idx
        scores, idx = pad_s, pad_i
    s_all, i_all = gather_topk(scores, idx, group)
    return ctx.merge_topk(s_all, i_all, min(int(keep_n), int(dictionary_size)))
