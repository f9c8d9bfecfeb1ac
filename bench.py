"""Benchmark of the dictionary-indexing hot path (BASELINE.json metric: EBSD patterns indexed
per second, 60x60 patterns, 100 000-entry dictionary, NCC, keep_n = 20).

  python bench.py --gpus N --steps K --warmup W          # this repo (CUDA, libkdi)
  python bench.py --impl reference --gpus N ...          # the reference's CPU arithmetic (oracle port)

One "step" = one full pass of the hot path over one batch of synthetic input:
normalise experimental + dictionary rows, tensor-core match with fused top-k, exact rescoring,
(N > 1: all-gather of the per-shard top-k + merge).  N = 1 is BASELINE.json configs[1]
(10 000 patterns vs 100 000 dictionary entries).  For N > 1 the dictionary (100 000 entries) is
sharded over the ranks and the pattern count grows with N (10 000 x N), so the work per GPU is
constant: weak scaling in patterns/s.

`value`   : inputs (raw uint8 patterns, raw float32 dictionary shard) resident in HBM, timed with
            CUDA events on the library's stream, max over ranks.
`e2e`     : the same job through the public API with pinned HOST buffers; H2D of both inputs
            and D2H of the result inside the timed region (wall clock between device syncs).
            N > 1: each rank uploads its dictionary shard and 1/N of the experimental rows (the
            raw rows are all-gathered over NVLink) and rank 0 reads the result back; byte counts
            are whole-job totals.
`e2e_pageable`: the same with ordinary (pageable) NumPy arrays, what a kikuchipy user passes.
`roofline`: the GEMM+top-k kernel; achieved = 2*M*N_shard*S / its CUDA-event duration; `frac` is
            against the measured BURST bf16 peak, `frac_of_sustained` against the sustained one.
`parity`  : checks of the objects the timed steps produced (structure, a 256-row sample against a
            float64 evaluation of the whole dictionary) and the planted-best hit rate of one extra
            step on planted patterns.
`extra`   : the other BASELINE.json configurations that fit this run (N = 1: configs[2];
            N = 8: configs[3] and configs[4]), timed and verified the same way on planted inputs.
`cpu_baseline`: the NumPy oracle (the reference's own arithmetic) on the host cores, on a
            bounded sample, extrapolated linearly in the number of patterns.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SIG = (60, 60)
S = SIG[0] * SIG[1]
M_PER_GPU = 10_000
N_DICT = 100_000
KEEP_N = 20
REF_CHUNK = 2083  # the reference's default dictionary chunk: 30 MB of float32 60x60 patterns


def workload_config(m_total: int) -> dict:
    """`config` of BOTH arms (identical strings: the driver compares them)."""
    return {
        "workload": f"{m_total} uint8 60x60 patterns vs {N_DICT}-entry float32 dictionary, NCC, keep_n={KEEP_N}",
        "l2": "inputs larger than L2 (raw dictionary %.0f MB; L2 126 MB)" % (N_DICT * S * 4 / 1e6),
    }


_BLAS_LIMIT = None


def blas_setup() -> dict:
    """Give the BLAS behind NumPy every core this process may use - torchrun exports
    OMP_NUM_THREADS=1 for nproc > 1, which would silently make the CPU arm single-threaded - and
    report what is in effect."""
    global _BLAS_LIMIT
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    info = {"numpy": np.__version__, "cores_available": cores,
            "env": {k: os.environ.get(k) for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS")}}
    try:
        from threadpoolctl import threadpool_info, threadpool_limits

        _BLAS_LIMIT = threadpool_limits(limits=cores)  # kept alive for the life of the process
        info["blas"] = [{k: lib.get(k) for k in ("internal_api", "version", "num_threads", "threading_layer", "architecture")}
                        for lib in threadpool_info() if lib.get("user_api") == "blas"]
        info["threads"] = max([lib.get("num_threads", 1) for lib in threadpool_info() if lib.get("user_api") == "blas"] or [1])
    except Exception as e:  # noqa: BLE001
        info["blas"] = f"threadpoolctl unavailable ({type(e).__name__}); BLAS thread count as inherited"
        info["threads"] = cores
    return info


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return p, "measured"
    except Exception:  # noqa: BLE001
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons, sampled in the background; ``window`` keeps the
    samples taken inside the timed region."""

    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int, period_ms: int = 20):
        self.index = index
        self.period_ms = max(5, int(period_ms))
        self.samples = []  # (host time the line was read, fields)
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", str(self.period_ms),
                 "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.time(), [x.strip() for x in line.strip().split(",")]))

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()

    def window(self, t0: float, t1: float) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, f in self.samples:
            if len(f) < 9 or not (t0 <= t <= t1 + 0.03):
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples in the timed region"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power)}


# --------------------------------------------------------------------------------------------
# CPU arm: the reference's arithmetic (oracle port), timed on the host cores
# --------------------------------------------------------------------------------------------

def cpu_sample(exp_u8: np.ndarray, dic_f32: np.ndarray, m_total: int):
    """Run the reference's chunk loop (prepare chunk -> einsum -> topk -> merge,
    _dictionary_indexing.py:94-128) for the sample patterns in ``exp_u8`` against the whole
    dictionary, timing dictionary preparation and matching separately; extrapolate the matching
    part linearly to ``m_total`` patterns.  Returns (patterns/s at m_total, detail dict)."""
    from oracle import di_oracle as orc  # the timed CPU arm (allowed use of oracle/)

    m_s = exp_u8.shape[0]
    n = dic_f32.shape[0]
    t0 = time.perf_counter()
    e = orc.prepare_experimental(exp_u8, "ncc", m_s)
    t_exp = time.perf_counter() - t0
    dic2 = dic_f32.reshape(n, -1)
    scores = np.full((m_s, KEEP_N), -1.0, dtype=np.float32)
    idx = np.zeros((m_s, KEEP_N), dtype=np.int32)
    t_prep = t_match = 0.0
    for start in range(0, n, REF_CHUNK):
        chunk = dic2[start:start + REF_CHUNK]
        t1 = time.perf_counter()
        d = orc.prepare_dictionary(chunk, "ncc")
        t2 = time.perf_counter()
        sim = orc.match(e, d)
        k = min(KEEP_N, chunk.shape[0])
        i_i = orc.argtopk(sim, k) + start
        s_i = orc.topk(sim, k)
        all_s = np.hstack((scores, s_i)); all_i = np.hstack((idx, i_i))
        best = np.argsort(-all_s, axis=1)[:, :KEEP_N]
        scores = np.take_along_axis(all_s, best, axis=1)
        idx = np.take_along_axis(all_i, best, axis=1)
        t3 = time.perf_counter()
        t_prep += t2 - t1
        t_match += t3 - t2
    scale = m_total / m_s
    t_full = t_prep + (t_exp + t_match) * scale
    detail = {"sample_patterns": m_s, "dictionary": n, "t_prepare_dictionary_s": round(t_prep, 3),
              "t_match_sample_s": round(t_match + t_exp, 3), "t_full_extrapolated_s": round(t_full, 3)}
    return m_total / t_full, detail


def host_inputs(m_sample: int, seed_exp=1, seed_dict=2):
    rng = np.random.default_rng(seed_exp)
    exp = rng.integers(0, 256, (m_sample,) + SIG, dtype=np.uint8)
    rng = np.random.default_rng(seed_dict)
    dic = rng.random((N_DICT,) + SIG, dtype=np.float32)
    return exp, dic


def run_reference(args, rank, world):
    if rank != 0:
        return
    blas = blas_setup()
    m_total = M_PER_GPU * args.gpus
    m_s = min(args.cpu_sample, m_total)
    exp, dic = host_inputs(m_s)
    vals, detail = [], None
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        v, detail = cpu_sample(exp, dic, m_total)
        if i >= args.warmup:
            vals.append((v, time.perf_counter() - t0))
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([t for _, t in vals])) * 1e3
    sample = (f"{m_s} of {m_total} patterns x full {N_DICT}-entry dictionary in {REF_CHUNK}-row chunks per step; "
              f"matching time scaled by {m_total}/{m_s}, dictionary preparation counted once")
    line = {
        "impl": "reference", "metric": "EBSD patterns indexed/sec (60x60, dict=100k, NCC, keep_n=20)",
        "value": value, "unit": "patterns/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(m_total),
        "note": "CPU: NumPy/BLAS restatement of the reference (kikuchipy is pure Python on dask+numpy; dask is not installable here)",
        "cpu_baseline": {"value": value, "unit": "patterns/s", "cores": blas["threads"], "kind": "port", "sample": sample,
                         "detail": detail, "blas": blas},
        "e2e": {"value": value, "unit": "patterns/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------

def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    import kikuchipy_b200 as kb
    from kikuchipy_b200 import _lib

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # one process per GPU: keep this rank's pinned buffers on the GPU's own NUMA node (N = 1 keeps
    # all cores for the CPU baseline)
    numa_cpus = kb.bind_to_gpu_numa_node(local_rank) if world > 1 and args.numa_bind else None
    ctx = kb.default_context(local_rank)
    if args.cta_group:
        ctx.set_option(_lib.OPT_CTA_GROUP, args.cta_group)
    if args.compute_dtype == "bf16":
        ctx.set_option(_lib.OPT_COMPUTE_DTYPE, 1)
    if args.strip_tiles:
        ctx.set_option(_lib.OPT_STRIP_TILES, args.strip_tiles)
    if args.superblock:
        ctx.set_option(_lib.OPT_SUPERBLOCK, args.superblock)
    if args.overlap is not None:
        ctx.set_option(_lib.OPT_OVERLAP, args.overlap)
    if args.cert_widen is not None:
        ctx.set_option(_lib.OPT_CERT_WIDEN, args.cert_widen)
    stream = torch.cuda.ExternalStream(ctx.stream_handle(), device=dev)

    m_total = M_PER_GPU * world
    start, end = kb.shard_bounds(N_DICT, world, rank)
    n_shard = end - start

    # synthetic inputs, generated on the device (identical experimental set on every rank)
    g = torch.Generator(device=dev); g.manual_seed(1)
    exp_dev = torch.randint(0, 256, (m_total,) + SIG, dtype=torch.uint8, device=dev, generator=g)
    g.manual_seed(2 + rank)
    dict_dev = torch.rand((n_shard,) + SIG, dtype=torch.float32, device=dev, generator=g)
    idx_dev = torch.empty((m_total, KEEP_N), dtype=torch.int64, device=dev)
    sc_dev = torch.empty((m_total, KEEP_N), dtype=torch.float32, device=dev)
    torch.cuda.synchronize()

    result = {}

    def step_device(exp=None):
        exp = exp_dev if exp is None else exp
        if world == 1:
            ctx.dictionary_indexing(exp, m_total, dict_dev, n_shard, _lib.KDI_NCC, KEEP_N,
                                    index_offset=start, out=(idx_dev, sc_dev))
            result["idx"], result["sc"] = idx_dev, sc_dev
        else:
            # candidates per shard -> exchange by row slice -> owner rescoring -> finalize -> gather
            result["idx"], result["sc"] = kb.dictionary_indexing_sharded(exp, dict_dev, N_DICT, metric="ncc",
                                                                     keep_n=KEEP_N, context=ctx)
        return ctx.timings()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank, args.clock_sample_ms)
    if rank == 0:
        sampler.start()  # nvidia-smi needs a few hundred ms before its first sample: start early
    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            step_device()
        barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        t_start = time.time()
        e0.record(stream)
        gemm_ms, launches, tms = [], 0, []
        for _ in range(args.steps):
            tm = step_device()
            gemm_ms.append(tm["gemm_topk_ms"]); launches += tm["kernel_launches"]
            tms.append(tm)
        e1.record(stream)
        barrier()
        t_end = time.time()
        clocks = sampler.window(t_start, t_end) if rank == 0 else None
        sampler.stop()
        dev_ms = e0.elapsed_time(e1)
    t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    ms_per_step = dev_ms / args.steps
    value = m_total / (ms_per_step * 1e-3)

    # ---- parity of the TIMED objects (outside the timed region) ----------------------------------
    from tools import di_configs as cfgs

    dictionary = cfgs.ShardedDictionary(N_DICT, S, world, dev, seed=2, shard_bounds=kb.shard_bounds)
    parity = cfgs.structural_checks(result["idx"], result["sc"], N_DICT)
    if rank == 0:
        rows = torch.linspace(0, m_total - 1, 256, device=dev).long().unique()
        parity.update(cfgs.float64_check(exp_dev, rows, dictionary, "ncc", KEEP_N, None, result["idx"], result["sc"]))
    parity["flagged_rows_last_step"] = int(tms[-1]["flagged_rows"])
    if world == 1:
        # how the rows of the last timed step were certified (KDI_OPT_CERT_STRICT = 2, the default): by the
        # worst-case bound on the tensor-core error (a proof), on the measured error model, or not at all
        # (exact path)
        parity["rows_certified"] = {"by_bound": m_total - int(tms[-1]["model_rows"]) - int(tms[-1]["flagged_rows"]),
                                    "by_model": int(tms[-1]["model_rows"]), "exact_path": int(tms[-1]["flagged_rows"]),
                                    "bound": ctx.certificate_bound(S, 1 if args.compute_dtype == "bf16" else 0)}
    # the same step with the STRICT certificate (KDI_OPT_CERT_STRICT = 1: the bound only, rows it cannot
    # decide go to the exact path) and with the model only (0, 32-entry lists: round 1's certificate)
    if world == 1 and not args.no_extras and args.compute_dtype != "bf16":
        modes = {}
        ref_idx, ref_sc = idx_dev.clone(), sc_dev.clone()
        n_alt = max(3, min(args.steps, 10))
        for label, mode in (("strict", 1), ("model_only", 0)):
            try:
                ctx.set_option(_lib.OPT_CERT_STRICT, mode)
                with torch.cuda.stream(stream):
                    for _ in range(3):
                        step_device()
                    torch.cuda.synchronize()
                    s0 = torch.cuda.Event(enable_timing=True); s1 = torch.cuda.Event(enable_timing=True)
                    s0.record(stream)
                    for _ in range(n_alt):
                        tm_s = step_device()
                    s1.record(stream)
                    torch.cuda.synchronize()
                modes[label] = {"ms_per_step": round(s0.elapsed_time(s1) / n_alt, 4), "steps": n_alt,
                                "exact_path_rows": int(tm_s["flagged_rows"]), "by_model_rows": int(tm_s["model_rows"]),
                                "identical_to_default": bool(torch.equal(ref_idx, idx_dev) and torch.equal(ref_sc, sc_dev))}
            except Exception as e:  # noqa: BLE001 - never lose the main line to an extra
                modes[label] = {"error": f"{type(e).__name__}: {e}"}
            finally:
                ctx.set_option(_lib.OPT_CERT_STRICT, 2)
        del ref_idx, ref_sc
        parity["certificate_modes"] = modes
    # one extra (untimed) step on PLANTED patterns of the same shape: the planted dictionary row must
    # be the best match of every pattern
    with torch.cuda.stream(stream):
        planted, j = cfgs.planted_patterns(dictionary, dict_dev, rank, m_total)
        step_device(planted.reshape((m_total,) + SIG))
        torch.cuda.synchronize()
        parity["planted_hit_rate"] = float((result["idx"][:, 0] == j).double().mean())
        del planted

    # ---- end to end: pinned host inputs -> public API -> host result ----------------------
    exp_host = ctx.pinned_empty((m_total,) + SIG, np.uint8)
    dict_host = ctx.pinned_empty((n_shard,) + SIG, np.float32)
    exp_host[...] = exp_dev.cpu().numpy()
    dict_host[...] = dict_dev.cpu().numpy()
    # whole-job bytes per step: N = 1 uploads everything over one link; N > 1: every rank uploads
    # its dictionary shard and 1/N of the experimental rows (all-gathered over NVLink), and rank 0
    # reads the result back
    h2d = exp_host.nbytes + N_DICT * S * 4
    d2h = m_total * KEEP_N * 12  # the result (identical on every rank) is read back by rank 0

    def step_e2e():
        if world == 1:
            res = kb.dictionary_indexing(exp_host, dict_host, metric="ncc", keep_n=KEEP_N, verbose=False)
            return res.scores
        i, s = kb.dictionary_indexing_sharded(exp_host, dict_host, N_DICT, metric="ncc", keep_n=KEEP_N, context=ctx)
        return (i.cpu(), s.cpu()) if rank == 0 else (i, s)

    with torch.cuda.stream(stream):
        for _ in range(max(1, args.warmup // 2)):
            step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            step_e2e()
        barrier()
        e2e_ms = (time.perf_counter() - t0) * 1e3 / args.e2e_steps
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())

    # the same from ordinary (pageable) NumPy arrays - what a kikuchipy user passes
    exp_page, dict_page = np.array(exp_host), np.array(dict_host)

    def step_pageable():
        if world == 1:
            return kb.dictionary_indexing(exp_page, dict_page, metric="ncc", keep_n=KEEP_N, verbose=False).scores
        i, s_ = kb.dictionary_indexing_sharded(exp_page, dict_page, N_DICT, metric="ncc", keep_n=KEEP_N, context=ctx)
        return (i.cpu(), s_.cpu()) if rank == 0 else (i, s_)

    with torch.cuda.stream(stream):
        step_pageable()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            step_pageable()
        barrier()
        page_ms = (time.perf_counter() - t0) * 1e3 / args.e2e_steps
    t = torch.tensor([page_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    page_ms = float(t.item())
    del exp_page, dict_page

    # ---- end to end with the dictionary GENERATED on the device: the
    # workflow the reference runs with a lazy dictionary (get_patterns(compute=False) ->
    # dictionary_indexing projects each chunk on the CPU inside the loop); inputs per step are the
    # pinned host patterns and this rank's rotations, no float32 dictionary crosses PCIe
    gen_ms = None
    if not args.no_generated:
        from kikuchipy_b200 import synthetic as po  # synthetic master pattern / rotations

        mu, ml = po.synthetic_master_pattern(args.master_pattern_size, seed=5)
        dc = kb.direction_cosines([-0.9, 0.85, -0.7, 0.95], 0.5, SIG[0], SIG[1], po.tilted_detector_matrix(70.0))
        rot = po.random_rotations(N_DICT, seed=4)[start:end]
        gen = kb.get_patterns(mu, ml, rot, direction_cosines=dc, detector_shape=SIG, context=ctx)

        def step_gen():
            if world == 1:
                res = kb.dictionary_indexing(exp_host, gen, metric="ncc", keep_n=KEEP_N, verbose=False, context=ctx)
                return res.scores
            i, s = kb.dictionary_indexing_sharded(exp_host, gen, N_DICT, metric="ncc", keep_n=KEEP_N, context=ctx)
            return (i.cpu(), s.cpu()) if rank == 0 else (i, s)

        with torch.cuda.stream(stream):
            for _ in range(2):
                step_gen()
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.e2e_steps):
                step_gen()
            barrier()
            gen_ms = (time.perf_counter() - t0) * 1e3 / args.e2e_steps
        gen_tm = ctx.timings()
        t = torch.tensor([gen_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        gen_ms = float(t.item())

    # ---- the other BASELINE.json configurations that fit this run (every rank takes part) --------
    extra = {}
    if not args.no_extras:
        # (50 = config 5 with the dictionary generated on the device instead of streamed from host memory)
        numbers = [int(x) for x in args.extras.split(",") if x] if args.extras else {1: [3], 8: [4, 5, 50]}.get(world, [])
        for number in numbers:
            label = "config5_generated" if number == 50 else f"config{number}"
            try:
                r = cfgs.run_config(number, ctx, rank, world, dev, steps=3, warmup=2, sample64=256)
                extra[label] = r
            except Exception as e:  # noqa: BLE001 - never lose the main line to an extra
                extra[label] = {"error": f"{type(e).__name__}: {e}"}
            torch.cuda.empty_cache()
    if rank != 0:
        return
    pk, pk_src = peaks()
    # the tensor-core pass of one step is a few launches of the same kernel (first dictionary
    # quarter, then one per row-block group) that together compute every (pattern, dictionary row)
    # pair once; gemm_topk_ms is the CUDA-event span from the first launch's start to the last
    # one's end, so achieved = the step's algorithmic flops / that span
    g_ms = float(np.mean(gemm_ms))
    g_launches = float(np.mean([x["gemm_launches"] for x in tms]))
    flops = 2.0 * m_total * n_shard * S
    achieved = flops / (g_ms * 1e-3) / 1e12
    peak = float(pk["bf16_tflops"])  # burst: the kernel runs for a few ms inside a short step
    peak_sustained = float(pk.get("bf16_tflops_sustained", pk["bf16_tflops"]))
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "gemm_traffic.json")) as f:
            traffic = json.load(f).get("dram_bytes_per_step")
    except Exception:  # noqa: BLE001
        pass
    last = tms[-1]
    line = {
        "metric": "EBSD patterns indexed/sec (60x60, dict=100k, NCC, keep_n=20)",
        "value": value, "unit": "patterns/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "fp16" if args.compute_dtype != "bf16" else "bf16", "data": "synthetic",
        "config": workload_config(m_total),
        "detail": {
            "dictionary_rows_per_gpu": n_shard,
            "operands": "16-bit tensor-core candidates (fp32 accumulate) + exact fp32 rescoring of every reported score",
            "cta_group": args.cta_group or 2,
            "numa_bound_cpus_rank0": (len(numa_cpus) if numa_cpus else None),
            "stage_ms": {k: round(float(np.mean([x[k] for x in tms])), 4)
                         for k in ("normalize_exp_ms", "normalize_dict_ms", "gemm_topk_ms", "rescore_ms",
                                   "fallback_ms", "total_ms")},
            "flagged_rows_last_step": int(last["flagged_rows"]),
        },
        "clocks": clocks,
        "parity": parity,
        "e2e": {"value": m_total / (e2e_ms * 1e-3), "unit": "patterns/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": args.e2e_steps},
        "e2e_pageable": {"value": m_total / (page_ms * 1e-3), "unit": "patterns/s", "ms_per_step": page_ms,
                         "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": args.e2e_steps,
                         "note": "inputs are ordinary NumPy arrays; the library stages them through its own pinned buffers"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                     "frac": achieved / peak, "frac_of_sustained": achieved / peak_sustained, "peak_sustained": peak_sustained,
                     "traffic": traffic, "kernel": "kdi_gemm_kernel (GEMM + fused top-k)",
                     "ms_per_launch": g_ms / max(g_launches, 1.0), "launches_per_step": g_launches,
                     "ms_per_step": g_ms, "flops_per_step": flops,
                     "traffic_note": "DRAM bytes of the step's GEMM launches at N=1 (ncu, profiles/gemm_traffic.json)",
                     "peak_source": f"{pk_src}: burst bf16 peak (cuBLAS, best of 10) as the denominator of frac; sustained alongside"},
    }
    if gen_ms is not None:
        line["e2e_generated"] = {
            "value": m_total / (gen_ms * 1e-3), "unit": "patterns/s", "ms_per_step": gen_ms,
            "h2d_bytes_per_step": int(exp_host.nbytes + N_DICT * 32), "d2h_bytes_per_step": int(d2h),
            "steps": args.e2e_steps,
            "note": f"dictionary generated on the device from {N_DICT} rotations of a {args.master_pattern_size}x"
                    f"{args.master_pattern_size} two-hemisphere master pattern (get_patterns fused into the prepare "
                    "step); host inputs per step: patterns + rotations",
            "rank0_stage_ms": {k: round(float(gen_tm[k]), 4) for k in ("normalize_exp_ms", "normalize_dict_ms",
                                                                        "gemm_topk_ms", "rescore_ms", "total_ms")},
        }
    if extra:
        line["extra"] = extra
    if world == 1 and not args.no_cpu:
        blas = blas_setup()
        exp_s = exp_host[: args.cpu_sample]
        v, detail = cpu_sample(np.array(exp_s), np.asarray(dict_host), m_total)
        line["cpu_baseline"] = {
            "value": v, "unit": "patterns/s", "cores": blas["threads"], "kind": "port",
            "sample": f"{args.cpu_sample} of {m_total} patterns x full dictionary in {REF_CHUNK}-row chunks; matching "
                      f"time scaled linearly to {m_total} patterns, dictionary preparation counted once",
            "detail": detail, "blas": blas,
        }
    if world == 1 and not args.no_extras:
        try:
            line["neighbouring_rows"] = neighbouring_rows(ctx)
        except Exception as e:  # noqa: BLE001 - never lose the main line to an extra
            line["neighbouring_rows"] = {"error": f"{type(e).__name__}: {e}"}
    print(json.dumps(line), flush=True)


def cpu_preprocess_sample(pats, bg):
    """patterns/s of the oracle port (SciPy / NumPy, one core) for static + dynamic background removal."""
    from oracle import preprocess_oracle as pp  # checker / CPU baseline only

    t0 = time.perf_counter()
    pp.remove_dynamic_background(pp.remove_static_background(pats, bg))
    return round(len(pats) / (time.perf_counter() - t0), 1)


def cpu_refine_sample(mu, ml, dc, pats, x0):
    """patterns/s of the oracle port (NumPy + the restated Nelder-Mead, one core) for orientation refinement."""
    from oracle import refinement_oracle as ro  # checker / CPU baseline only

    prob = ro.Problem(mu, ml, SIG[0], SIG[1], direction_cosines=dc)
    t0 = time.perf_counter()
    ro.refine_orientation(prob, pats.reshape(len(pats), -1), x0, False)
    return round(len(pats) / (time.perf_counter() - t0), 2)


def neighbouring_rows(ctx):
    """Device timings (CUDA events inside the library) of the rows either side of the path
    (SURVEY.md section 8f): preprocessing of the measured patterns and orientation refinement of
    the indexed ones, on synthetic inputs of the benchmark's pattern size.  Not part of `value`."""
    import torch

    import kikuchipy_b200 as kb
    from kikuchipy_b200 import _lib
    from kikuchipy_b200 import synthetic as syn

    out = {}
    rng = np.random.default_rng(7)
    pats = rng.integers(0, 256, (M_PER_GPU, SIG[0], SIG[1]), dtype=np.uint8)
    bg = pats[:16].mean(axis=0).astype(np.uint8)
    dev = torch.from_numpy(pats).cuda()
    for _ in range(2):
        kb.preprocess(dev, static_bg=bg)
    ms = ctx.timings()["total_ms"]
    out["preprocess_static+dynamic"] = {"patterns": M_PER_GPU, "kernel_ms": round(ms, 3),
                                        "patterns_per_s": round(M_PER_GPU / ms * 1e3),
                                        "cpu_port_patterns_per_s_1core": cpu_preprocess_sample(pats[:200], bg)}
    n_ref, mp_n = 4096, 501
    mu, ml = syn.synthetic_master_pattern(mp_n, seed=5)
    dc = kb.direction_cosines([-0.9, 0.85, -0.7, 0.95], 0.5, SIG[0], SIG[1], syn.tilted_detector_matrix(70.0))
    eu = np.stack([rng.uniform(0.2, 6.0, n_ref), rng.uniform(0.2, 2.9, n_ref), rng.uniform(0.2, 6.0, n_ref)], axis=1)
    quat = kb.refinement.euler_to_quaternion(eu)
    mp = ctx.master_pattern(mu, ml, dc)
    sim = ctx.project_patterns(mp, quat)
    lo, hi = sim.min(axis=1, keepdims=True), sim.max(axis=1, keepdims=True)
    pats8 = np.round((sim - lo) / (hi - lo) * 255).astype(np.uint8)
    x0 = (eu + np.deg2rad(rng.uniform(-1, 1, eu.shape)))[:, None, :]
    for _ in range(2):
        res = ctx.refine(mp, _lib.REFINE_ORI, pats8, SIG[0], SIG[1], False, x0)
    ms = ctx.timings()["total_ms"]
    out["refine_orientation"] = {"patterns": n_ref, "kernel_ms": round(ms, 3), "patterns_per_s": round(n_ref / ms * 1e3),
                                 "cpu_port_patterns_per_s_1core": cpu_refine_sample(mu, ml, dc, pats8[:8], x0[:8]),
                                 "mean_evaluations": round(float(res[:, 1].mean()), 1),
                                 "mean_score": round(float(res[:, 0].mean()), 5),
                                 "median_misorientation_to_truth_deg": round(float(np.degrees(np.median(
                                     2 * np.arccos(np.clip(np.abs(np.sum(kb.refinement.euler_to_quaternion(res[:, 2:5]) * quat, axis=1)), 0, 1))))), 4)}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-sample", type=int, default=2000)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the other BASELINE configurations and the preprocessing / refinement timings")
    ap.add_argument("--extras", default="", help="BASELINE configurations to run as extras (default: 3 at N = 1; 4,5 at N = 8)")
    ap.add_argument("--no-generated", action="store_true", help="skip the generated-dictionary end-to-end leg")
    ap.add_argument("--numa-bind", action="store_true",
                    help="bind each rank to its GPU's NUMA node (no effect on single-node-affinity boxes like this pool's)")
    ap.add_argument("--master-pattern-size", type=int, default=1001)
    ap.add_argument("--cta-group", type=int, default=0)
    ap.add_argument("--compute-dtype", default="fp16", choices=["fp16", "bf16"])
    ap.add_argument("--strip-tiles", type=int, default=0)
    ap.add_argument("--superblock", type=int, default=0)
    ap.add_argument("--cert-widen", type=int, default=None, help="KDI_OPT_CERT_WIDEN (A/B runs; default: the library's)")
    ap.add_argument("--clock-sample-ms", type=int, default=20, help="nvidia-smi sampling period during the timed region")
    ap.add_argument("--overlap", type=int, default=None, help="0: one kernel at a time; 1 (default): overlapped schedule")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if world != args.gpus and rank == 0:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}; using {world}", file=sys.stderr)
    run_ours(args, rank, world, local_rank)
    if world > 1:
        import torch.distributed as dist

        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
