"""Dictionary sharding across GPUs: one process per GPU, rank ``r`` holds
dictionary rows ``shard_bounds(N, world, r)`` and every experimental row.

The reduction is the one the reference runs serially over dictionary chunks
(/root/reference/src/kikuchipy/indexing/_dictionary_indexing.py:94-128: per-chunk top-k, chunk
offset added to the indices, running merge), spread over ranks.  The tensor-core pass (the
expensive part) runs on every rank against its own shard; everything that is per experimental
row afterwards is split by rows so that no rank repeats another rank's work:

  0. host input only: every rank uploads 1/world of the raw experimental rows over its own PCIe
     link and the raw bytes are all-gathered over NVLink
  1. candidates: this shard's ``kc`` best per row by tensor-core score, global indices
  2. the exchange of the path: all-to-all of the candidate lists by row slice, per-row merge
     of the ``world`` lists of this rank's slice
  3. all-gather of the merged candidate indices (every rank needs to know which of its
     dictionary rows were nominated)
  4. every rank rescores, in exact float32, the candidates whose dictionary rows it holds
  5. reduce-scatter(MAX) of the exact scores by row slice (each candidate has one owner)
  6. rank + certificate for this rank's slice
  7. all-gather of the finished slices (identical result on every rank)
  8. rows whose certificate failed anywhere: exact top-k per shard, gathered and merged

That is the COLLECTIVE-LIBRARY form of the exchange (``exchange="nccl"``): torch.distributed calls
around ``kdi_shard_*`` stages, written against a small ``stages`` interface so that it also runs under
``gloo`` on CPU tensors in the tests (with the oracle standing in for the GPU stages).

The default on a single NVLink / NVSwitch node is ``exchange="peer"``: the same pipeline inside
libkdi (``kdi_shard_run_peer``, csrc/kdi_comm.cu).  Every rank maps every other rank's "symmetric
block" through CUDA IPC once; the kernels then store their results straight into the block of the rank
that needs them (selection kernel -> slice owner, merge kernel -> request queue of the dictionary-row
owner, rescoring kernel -> slice owner's score table, finalize -> everybody) and the ranks meet at
device-side barriers.  No pack / unpack kernels, no host synchronisation between the steps, and the
owner rescoring walks a compact request list instead of every row.  torch.distributed is only used to
all-gather the 64-byte IPC handles when the blocks are (re)created, and for the rare flagged rows.
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


def row_slice(n_rows: int, world_size: int, rank: int) -> tuple[int, int, int]:
    """``(rows per slice, start, end)`` of the experimental rows rank ``rank`` post-processes:
    equal slices of ``ceil(n_rows / world_size)`` rows (the last ones may be short or empty), so
    the collectives exchange equal-sized blocks."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError("bad rank / world size")
    per = -(-int(n_rows) // world_size) if n_rows > 0 else 0
    start = min(rank * per, n_rows)
    return per, start, min(start + per, n_rows)


def gather_topk(scores, indices, group=None):
    """All-gather per-rank ``(M, k)`` score / index tensors into ``(world, M, k)`` tensors
    (list-major: what ``kdi_merge_topk`` expects).  Works on CUDA tensors (NCCL) and CPU
    tensors (gloo)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    scores, indices = scores.contiguous(), indices.contiguous()
    s_all = torch.empty((world,) + tuple(scores.shape), dtype=scores.dtype, device=scores.device)
    i_all = torch.empty((world,) + tuple(indices.shape), dtype=indices.dtype, device=indices.device)
    # (concatenation form along dim 0: the one form both NCCL and gloo accept)
    dist.all_gather_into_tensor(s_all.view((-1,) + tuple(scores.shape[1:])), scores, group=group)
    dist.all_gather_into_tensor(i_all.view((-1,) + tuple(indices.shape[1:])), indices, group=group)
    return s_all, i_all


def gather_experimental(experimental_host: np.ndarray, n_rows: int, device, group=None):
    """Step 0: upload rows ``row_slice(n_rows, world, rank)`` of a raw host pattern array and
    all-gather the raw bytes, so each PCIe link carries 1/world of the experimental set.
    Returns a ``(n_rows, S)`` device tensor of the input dtype."""
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    flat = np.ascontiguousarray(experimental_host).reshape(n_rows, -1)
    per, start, end = row_slice(n_rows, world, rank)
    mine = torch.zeros((per, flat.shape[1]), dtype=torch.from_numpy(flat[:0]).dtype, device=device)
    if end > start:
        mine[: end - start].copy_(torch.from_numpy(flat[start:end]), non_blocking=True)
    full = torch.empty((per * world, flat.shape[1]), dtype=mine.dtype, device=device)
    dist.all_gather_into_tensor(full.view(torch.uint8), mine.view(torch.uint8), group=group)
    return full[:n_rows]


def gather_rows(local: np.ndarray, counts, group=None) -> np.ndarray:
    """All-gather of per-rank row blocks of a float64 result array (``counts[r]`` rows from rank
    ``r``, concatenated in rank order on every rank).  Used where the path partitions into
    independent units with no exchange step (refinement: one pattern = one unit): each rank works
    on its own slice and only the finished rows travel.  NCCL moves device tensors, gloo (the CPU
    tests) host tensors."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    if len(counts) != world:
        raise ValueError("one row count per rank is needed")
    width = int(local.shape[1])
    per = max(int(c) for c in counts)
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    mine = torch.zeros((per, width), dtype=torch.float64, device=dev)
    mine[: local.shape[0]] = torch.from_numpy(np.ascontiguousarray(local, dtype=np.float64)).to(dev)
    out = torch.empty((world * per, width), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(out, mine, group=group)
    out = out.cpu().numpy().reshape(world, per, width)
    return np.concatenate([out[r, : int(counts[r])] for r in range(world)], axis=0)


_PEER_COMMS: dict = {}


def peer_comm(ctx, nbytes: int, group=None):
    """This process's :class:`kikuchipy_b200._lib.PeerComm` for ``(ctx, group)`` with at least ``nbytes``
    of symmetric memory; (re)created collectively - every rank must ask for the same size - with the
    IPC handles all-gathered through torch.distributed."""
    import torch
    import torch.distributed as dist

    from . import _lib

    key = (id(ctx), id(group))
    comm = _PEER_COMMS.get(key)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if comm is not None and (comm._ctx is not ctx or comm.rank != rank or comm.world != world):
        comm = None  # the key of a group object that no longer exists: not this job's mapping
    if comm is not None and comm._h and comm.nbytes >= nbytes:
        return comm
    if comm is not None:
        torch.cuda.synchronize()
        dist.barrier(group)  # nobody still stores into a block that is about to go away
        comm.close()
    nbytes = int(nbytes * 1.25) + (1 << 20)  # head room: slightly larger jobs reuse the mapping
    comm = _lib.PeerComm(ctx, rank, world, nbytes)
    nccl = dist.get_backend(group) == "nccl"
    dev = torch.device("cuda", ctx.device) if nccl else torch.device("cpu")
    mine = torch.frombuffer(bytearray(bytes(comm.handle)), dtype=torch.uint8).to(dev)
    allh = torch.empty((world * _lib.IPC_HANDLE_BYTES,), dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(allh, mine, group=group)
    comm.connect(bytes(allh.cpu().numpy().tobytes()))
    dist.barrier(group)  # every rank has mapped every block (and zeroed its own) before anyone stores
    _PEER_COMMS[key] = comm
    return comm


def _peer_pipeline(ctx, experimental, n_exp_all, dictionary_shard, n_shard, code, keep_n, navigation_mask, kept,
                   dictionary_size, group):
    """exchange="peer": one library call per rank + the rare flagged rows."""
    import torch
    import torch.distributed as dist

    from .master_pattern import GeneratedDictionary

    world = dist.get_world_size(group)
    comm = peer_comm(ctx, ctx.comm_bytes_needed(world, kept, keep_n), group)
    source = ((dictionary_shard.master_pattern, dictionary_shard.rotations)
              if isinstance(dictionary_shard, GeneratedDictionary) else dictionary_shard)
    shard, idx, scores, flags = ctx.shard_run_peer(comm, experimental, n_exp_all, source, n_shard, code, keep_n,
                                                   dictionary_size, nav_mask=navigation_mask)
    try:
        if flags.numel() > 0:  # the same list on every rank: exact top-k per shard, gathered and merged
            rows = torch.sort(flags).values
            k_local = min(keep_n, n_shard)
            fi, fs = shard.exact_rows(rows, k_local)
            if k_local != keep_n:
                fs = torch.nn.functional.pad(fs, (0, keep_n - k_local), value=-float("inf"))
                fi = torch.nn.functional.pad(fi, (0, keep_n - k_local), value=-1)
            fs_all, fi_all = gather_topk(fs, fi, group)
            mi, ms = ctx.merge_topk(fs_all, fi_all, keep_n)
            idx[rows.long()] = mi
            scores[rows.long()] = ms
    finally:
        shard.close()
    return idx, scores


def _pack(scores, indices):
    """(float32 scores, int64 indices) of equal shape -> one byte tensor ``(..., 12)`` so that both
    travel in ONE collective (one concatenation kernel; NCCL launch latency dominates here)."""
    import torch

    s = scores.contiguous()
    i = indices.contiguous()
    return torch.cat([s.view(torch.uint8).view(s.shape + (4,)), i.view(torch.uint8).view(i.shape + (8,))], dim=-1)


def _unpack(packed):
    import torch

    scores = packed[..., :4].contiguous().view(torch.float32).squeeze(-1)
    indices = packed[..., 4:].contiguous().view(torch.int64).squeeze(-1)
    return scores, indices


class _KdiStages:
    """The GPU stages of the pipeline: thin calls into libkdi (``kdi_shard_*``)."""

    def __init__(self, ctx, experimental, n_exp_all, dictionary_shard, n_shard, code, keep_n, navigation_mask,
                 start):
        self.ctx, self.args = ctx, (experimental, n_exp_all, dictionary_shard, n_shard, code, keep_n)
        self.nav, self.start, self.shard = navigation_mask, start, None

    def candidates(self, pad_rows):
        from .master_pattern import GeneratedDictionary

        experimental, n_exp_all, dictionary_shard, n_shard, code, keep_n = self.args
        if isinstance(dictionary_shard, GeneratedDictionary):  # this rank's rows are generated from its rotations
            self.shard, approx, gidx = self.ctx.shard_candidates_projected(
                experimental, n_exp_all, dictionary_shard.master_pattern, dictionary_shard.rotations, code, keep_n,
                nav_mask=self.nav, index_offset=self.start, pad_rows=pad_rows)
        else:
            self.shard, approx, gidx = self.ctx.shard_candidates(*self.args, nav_mask=self.nav,
                                                                 index_offset=self.start, pad_rows=pad_rows)
        return approx, gidx, self.shard.kc

    def merge(self, s_all, i_all, k):
        return self.ctx.merge_topk(s_all, i_all, k)

    def rescore_owned(self, gidx, approx, keep_n):
        return self.shard.rescore_owned(gidx, approx, keep_n)

    def finalize(self, approx, gidx, exact, keep_n, dict_total, row0, rows):
        return self.shard.finalize(approx, gidx, exact, keep_n, dict_total, row0=row0, rows=rows)

    def exact_rows(self, rows, k_local):
        return self.shard.exact_rows(rows, k_local)

    def close(self):
        if self.shard is not None:
            self.shard.close()


def run_sharded_pipeline(stages, n_rows: int, keep_n: int, dictionary_size: int, n_shard: int, group=None,
                         trace=None):
    """Steps 1-8 of the module docstring over ``stages`` (see ``_KdiStages``).  Returns
    ``(indices, scores)`` tensors ``(n_rows, keep_n)``, identical on every rank."""
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    trace = trace or (lambda name: None)
    per, r0, r1 = row_slice(n_rows, world, rank)
    padded = per * world
    # 1. this shard's candidates by tensor-core score (global indices), padded to equal slices
    approx, gidx, kc = stages.candidates(padded)
    trace("candidates")
    # 2. all-to-all by row slice: block j of the send buffer (rows of slice j) goes to rank j;
    #    the receive buffer is list-major (list l = rank l's candidates for MY rows).  Scores and
    #    indices travel as one 12-byte record per candidate (one collective instead of two).
    packed_in = torch.empty((world, per, kc, 12), dtype=torch.uint8, device=approx.device)
    dist.all_to_all_single(packed_in, _pack(approx, gidx), group=group)
    s_in, i_in = _unpack(packed_in)
    trace("all_to_all")
    my_idx, my_approx = stages.merge(s_in, i_in, kc)
    trace("merge")
    # 3. everyone learns every row's merged candidates (indices + tensor-core scores)
    g_packed = torch.empty((padded, kc, 12), dtype=torch.uint8, device=my_idx.device)
    dist.all_gather_into_tensor(g_packed, _pack(my_approx, my_idx), group=group)
    g_approx, g_idx = _unpack(g_packed)
    trace("all_gather_idx")
    # 4. exact scores of the candidates whose dictionary rows this rank holds (-inf elsewhere, and
    #    for candidates too far below the keep_n-th tensor-core score to matter)
    exact = stages.rescore_owned(g_idx, g_approx, keep_n)
    trace("rescore_owned")
    # 5. each candidate has exactly one owner: MAX combines, scattered back by row slice
    my_exact = torch.empty((per, kc), dtype=exact.dtype, device=exact.device)
    dist.reduce_scatter_tensor(my_exact, exact.contiguous(), op=dist.ReduceOp.MAX, group=group)
    trace("reduce_scatter")
    # 6. rank by exact score + certificate for this rank's rows
    idx_s, sc_s, flags = stages.finalize(my_approx, my_idx, my_exact, keep_n, dictionary_size, r0, r1 - r0)
    trace("finalize")
    # 7. finished slices to everyone (+ how many rows each rank flagged)
    idx = torch.empty((padded, keep_n), dtype=idx_s.dtype, device=idx_s.device)
    scores = torch.empty((padded, keep_n), dtype=sc_s.dtype, device=sc_s.device)
    dist.all_gather_into_tensor(idx, idx_s.contiguous(), group=group)
    dist.all_gather_into_tensor(scores, sc_s.contiguous(), group=group)
    counts = torch.zeros((world,), dtype=torch.int64, device=idx_s.device)
    dist.all_gather_into_tensor(counts, torch.tensor([int(flags.numel())], dtype=torch.int64, device=idx_s.device),
                                group=group)
    counts = counts.cpu()
    idx, scores = idx[:n_rows], scores[:n_rows]
    trace("all_gather_results")
    # 8. rows whose certificate failed on any rank: exact top-k per shard, gathered and merged
    n_max = int(counts.max())
    if n_max > 0:
        mine = torch.full((n_max,), -1, dtype=torch.int32, device=idx_s.device)
        mine[: flags.numel()] = flags.to(torch.int32)
        rows = torch.empty((world * n_max,), dtype=torch.int32, device=idx_s.device)
        dist.all_gather_into_tensor(rows, mine, group=group)
        rows = torch.sort(rows[rows >= 0]).values  # same order on every rank
        k_local = min(keep_n, n_shard)
        fi, fs = stages.exact_rows(rows, k_local)
        if k_local != keep_n:  # a shard smaller than keep_n contributes what it has
            fs = torch.nn.functional.pad(fs, (0, keep_n - k_local), value=-float("inf"))
            fi = torch.nn.functional.pad(fi, (0, keep_n - k_local), value=-1)
        fs_all, fi_all = gather_topk(fs, fi, group)
        mi, ms = stages.merge(fs_all, fi_all, keep_n)
        idx[rows.long()] = mi
        scores[rows.long()] = ms
        trace("flagged_rows")
    return idx, scores


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
    exchange: str | None = None,
):
    """Index ``experimental`` against a dictionary whose rows ``shard_bounds(dictionary_size,
    world, rank)`` this rank holds in ``dictionary_shard``.

    ``exchange``: ``"peer"`` (default; ``KDI_EXCHANGE`` overrides) - the exchange inside libkdi over
    peer-mapped memory, for up to 8 ranks on one node - or ``"nccl"`` - torch.distributed collectives
    around the library's stages (any number of ranks / nodes).

    ``experimental`` (the same array on every rank) and ``dictionary_shard`` may be host arrays
    or CUDA tensors; ``dictionary_shard`` may also be a :class:`GeneratedDictionary` holding this
    rank's rotations (the shard is then generated on the device).  Returns ``(simulation_indices, scores)`` as CUDA tensors ``(M, keep_n)``
    (global dictionary indices, identical on every rank).  Needs an initialised
    ``torch.distributed`` process group with the NCCL backend.
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
    from .master_pattern import GeneratedDictionary

    generated = isinstance(dictionary_shard, GeneratedDictionary)
    if world == 1:
        scores = torch.empty((kept, keep_n), dtype=torch.float32, device=dev)
        idx = torch.empty((kept, keep_n), dtype=torch.int64, device=dev)
        if generated:
            ctx.dictionary_indexing_projected(experimental, n_exp_all, dictionary_shard.master_pattern,
                                              dictionary_shard.rotations, code, keep_n, nav_mask=navigation_mask,
                                              index_offset=start, out=(idx, scores))
        else:
            ctx.dictionary_indexing(experimental, n_exp_all, dictionary_shard, n_shard, code, keep_n,
                                    nav_mask=navigation_mask, index_offset=start, out=(idx, scores))
        return idx, scores

    # torch ops, NCCL collectives and library kernels all on the context's stream: no host syncs
    with torch.cuda.stream(ctx.torch_stream()):
        trace = _Trace(os.environ.get("KDI_TRACE") == "1" and rank == 0)
        if not (hasattr(experimental, "is_cuda") and experimental.is_cuda):
            experimental = np.asarray(experimental)
            if experimental.dtype in (np.uint8, np.uint16, np.float32, np.float64):
                experimental = gather_experimental(experimental, n_exp_all, dev, group)
                trace("upload_gather_experimental")
        if ctx.candidate_capacity(keep_n) == 0:
            if generated:
                raise NotImplementedError(f"keep_n {keep_n} is too large for a generated, sharded dictionary")
            return _sharded_exact_lists(ctx, experimental, n_exp_all, dictionary_shard, n_shard, code, keep_n,
                                        navigation_mask, start, kept, dictionary_size, group)
        exchange = exchange or os.environ.get("KDI_EXCHANGE", "peer")
        if exchange not in ("peer", "nccl"):
            raise ValueError("exchange must be 'peer' or 'nccl'")
        if exchange == "peer" and world <= 8 and dictionary_size >= world and kept < (1 << 24):
            idx, scores = _peer_pipeline(ctx, experimental, n_exp_all, dictionary_shard, n_shard, code, keep_n,
                                         navigation_mask, kept, dictionary_size, group)
            trace("peer_pipeline")
            trace.report(ctx)
            torch.cuda.current_stream().synchronize()
            return idx, scores
        stages = _KdiStages(ctx, experimental, n_exp_all, dictionary_shard, n_shard, code, keep_n, navigation_mask,
                            start)
        try:
            idx, scores = run_sharded_pipeline(stages, kept, keep_n, dictionary_size, n_shard, group, trace)
            trace.report(ctx)
            torch.cuda.current_stream().synchronize()
            return idx, scores
        finally:
            stages.close()


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
    idx, scores = ctx.merge_topk(s_all, i_all, min(int(keep_n), int(dictionary_size)))
    torch.cuda.current_stream().synchronize()
    return idx, scores
