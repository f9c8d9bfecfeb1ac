"""Parity of the CUDA path (through the C ABI / the Python mirror of the reference surface)
against the CPU oracle and the committed golden vectors.  Needs a B200.

Gates: scores within 1e-4 of the reference arithmetic (observed ~5e-8);
index lists identical wherever the reference's own scores are separated by more than ``tie_tol``
(1e-6; 2e-5 with a signal mask, where the reference's float32 row sums over the NumPy
fancy-indexed, non-contiguous array are themselves ~1e-5 noisy); strictly identical on the
reference fixtures and on planted inputs.
"""

import numpy as np
import pytest

import kikuchipy_b200 as kb
from kikuchipy_b200 import _lib
from oracle import di_oracle as orc

pytestmark = pytest.mark.gpu

SCORE_TOL = 1e-4


@pytest.fixture(scope="module")
def ctx():
    c = kb.default_context(0)
    yield c
    for opt in (_lib.OPT_COMPUTE_DTYPE, _lib.OPT_FORCE_EXACT, _lib.OPT_STRIP_TILES, _lib.OPT_SUPERBLOCK,
                _lib.OPT_MAX_STAGES, _lib.OPT_GEMM_SMS, _lib.OPT_MIN_GROUPS, _lib.OPT_POST_PER_GROUP,
                _lib.OPT_GEMM_SERIAL, _lib.OPT_SM_PARTITION):
        c.set_option(opt, 0)
    c.set_option(_lib.OPT_DEP_FLAGS, 0)
    c.set_option(_lib.OPT_CERT_STRICT, 2)
    c.set_option(_lib.OPT_CERT_WIDEN, 0)
    c.set_option(_lib.OPT_SPLIT_SELECT, 1)
    c.set_option(_lib.OPT_CTA_GROUP, 2)
    c.set_option(_lib.OPT_OVERLAP, 1)
    c.set_signal_mask(None)


def _check(ridx, rsc, idx, sc, tie_tol=1e-6, strict=False):
    assert idx.dtype == np.int64 and sc.dtype == np.float32
    assert idx.shape == ridx.shape and sc.shape == rsc.shape
    r = orc.compare_topk(ridx, rsc, idx, sc, tie_tol=tie_tol, score_tol=SCORE_TOL)
    assert r["scores_ok"], r
    assert r["tie_ok"], r
    if strict:
        assert r["exact_rows"] == 1.0, r
    # best first
    assert np.all(np.diff(sc, axis=1) <= 0)
    return r


# ---- prepare_* -------------------------------------------------------------------------------

@pytest.mark.parametrize("src_dtype", [np.uint8, np.uint16, np.float32, np.float64])
@pytest.mark.parametrize("masked, nav", [(True, True), (True, False), (False, True), (False, False)])
def test_bulk_staged_normalise_is_bit_identical(ctx, src_dtype, masked, nav):
    """The kernel that stages raw rows with cp.async.bulk and compacts them run by run must produce
    exactly the rows of the scattered-load kernel - for every source type, with and without the signal
    mask / the navigation mask, float32 rows and 16-bit operands (checked through the scores) alike."""
    sig = (36, 40)  # 1440 pixels: rows of 1440 / 2880 / 5760 / 11520 bytes, all multiples of 16
    rng = np.random.default_rng(5)
    raw = (rng.random((700,) + sig) * (60000 if src_dtype == np.uint16 else 255)).astype(src_dtype)
    dic = orc.synthetic_dictionary(1500, sig, seed=6)
    smask = orc.circular_signal_mask(sig) if masked else None
    row_mask = (rng.random(700) < 0.3) if nav else None
    out = {}
    try:
        for bulk in (1, 0):
            ctx.set_option(_lib.OPT_BULK_NORMALIZE, bulk)
            ctx.set_signal_mask(smask)
            p = ctx.patterns(raw, 700, _lib.KDI_NCC, row_mask)
            rows = np.asarray(p)
            p.close()
            idx, sc = ctx.dictionary_indexing(raw, 700, dic, 1500, _lib.KDI_NDP, 7, nav_mask=row_mask)
            out[bulk] = (rows, idx, sc)
    finally:
        ctx.set_option(_lib.OPT_BULK_NORMALIZE, 0)
        ctx.set_signal_mask(None)
    for a, b in zip(out[1], out[0]):
        assert np.array_equal(a, b)
    want = orc.prepare_experimental(raw, "ncc", 700, navigation_mask=row_mask, signal_mask=smask)
    assert out[1][0].shape == want.shape and np.max(np.abs(out[1][0] - want)) < 2e-6


@pytest.mark.parametrize("src_dtype", [np.uint8, np.uint16, np.float32, np.float64])
def test_masked_prepare_large_row_count_matches_small(ctx, src_dtype):
    """Masked prepare of a larger set, every source type: same bits whatever the number of rows
    per call, and the reference's values."""
    rng = np.random.default_rng(3)
    sig = (40, 40)
    data = (rng.random((1500,) + sig) * 250).astype(src_dtype)
    smask = orc.circular_signal_mask(sig)
    ctx.set_signal_mask(smask)
    try:
        for code, name in ((_lib.KDI_NCC, "ncc"), (_lib.KDI_NDP, "ndp")):
            big = np.asarray(ctx.patterns(data, 1500, code))
            small = np.concatenate([np.asarray(ctx.patterns(data[a:a + 500], 500, code)) for a in (0, 500, 1000)])
            assert np.array_equal(big, small)
            ref = orc.prepare_dictionary(data.reshape(1500, -1).astype(np.float32), name, smask)
            assert big.shape == ref.shape and np.abs(big - ref).max() < 3e-7
    finally:
        ctx.set_signal_mask(None)


@pytest.mark.parametrize("metric", ["ncc", "ndp"])
def test_prepared_rows_match_reference(ctx, golden, metric):
    g = golden("config1_nickel_x_1000.npz")
    code = _lib.KDI_NCC if metric == "ncc" else _lib.KDI_NDP
    ctx.set_signal_mask(None)
    got = np.asarray(ctx.patterns(g["nickel"], 9, code))
    assert got.shape == (9, 3600)
    assert np.max(np.abs(got - g[f"{metric}_prepared_exp"])) < 2e-8
    # navigation mask drops rows, signal mask compacts columns (False = keep)
    ctx.set_signal_mask(g["signal_mask"])
    got = np.asarray(ctx.patterns(g["nickel"], 9, code, row_mask=g["nav_mask"]))
    ref = orc.prepare_experimental(g["nickel"], metric, 9, g["nav_mask"], g["signal_mask"])
    assert got.shape == ref.shape
    assert np.max(np.abs(got - ref)) < 5e-7
    ctx.set_signal_mask(None)


@pytest.mark.parametrize("metric", ["ncc", "ndp"])
def test_metric_call_matches_golden_block(golden, metric):
    """metric(exp, dict) - reference use at tests/test_signals/test_ebsd_master_pattern.py:237-243."""
    g = golden("config1_nickel_x_1000.npz")
    dic = orc.synthetic_dictionary(1000, (60, 60), seed=2)
    cls = kb.NormalizedCrossCorrelationMetric if metric == "ncc" else kb.NormalizedDotProductMetric
    sim = cls(9, 1000)(g["nickel"], dic)
    assert sim.shape == (9, 1000) and sim.dtype == np.float32
    assert np.max(np.abs(sim - g[f"{metric}_sim_f32"])) < 1e-6
    assert np.max(np.abs(sim - g[f"{metric}_sim_f64"])) < 1e-6
    m = cls(9, 1000, signal_mask=g["signal_mask"])
    assert np.max(np.abs(m(g["nickel"], dic) - g[f"{metric}_sim_masked_f32"])) < 2e-5
    m = cls(9, 9)
    self_ = m(g["nickel"], g["nickel"].reshape(9, 60, 60))
    assert np.max(np.abs(self_ - g[f"{metric}_self_f32"])) < 1e-6
    assert np.allclose(np.diag(self_), 1, atol=1e-6)


def test_plugin_hooks_topk(golden):
    """The three hooks the reference driver calls (_dictionary_indexing.py:70,193-201)."""
    g = golden("config1_nickel_x_1000.npz")
    dic = orc.synthetic_dictionary(1000, (60, 60), seed=2)
    m = kb.NormalizedCrossCorrelationMetric(9, 1000)
    e = m.prepare_experimental(g["nickel"])
    d = m.prepare_dictionary(dic.reshape(1000, -1))
    sim = m.match(e, d)
    idx = sim.argtopk(5, axis=-1).reshape((-1, 5))
    sc = sim.topk(5, axis=-1).reshape((-1, 5))
    ref = g["ncc_sim_f32"]
    ridx = orc.argtopk(ref, 5)
    _check(ridx, orc.topk(ref, 5), idx, sc, strict=True)
    assert np.max(np.abs(np.asarray(sim) - ref)) < 1e-6


# ---- BASELINE config 1 and the golden driver vectors ------------------------------------------------

@pytest.mark.parametrize("cta_group", [1, 2])
def test_config1_nickel(ctx, golden, cta_group):
    g = golden("config1_nickel_x_1000.npz")
    dic = orc.synthetic_dictionary(1000, (60, 60), seed=2)
    ctx.set_option(_lib.OPT_CTA_GROUP, cta_group)
    res = kb.dictionary_indexing(g["nickel"], dic, metric="ncc", keep_n=5, verbose=False)
    ridx, rsc = orc.dictionary_indexing(g["nickel"], dic, metric="ncc", keep_n=5)
    _check(ridx, rsc, res.simulation_indices, res.scores, strict=True)
    assert res.shape == (3, 3) and res.size == 9 and res.rotations_per_point == 5
    assert ctx.timings()["gemm_launches"] == 1  # the tensor-core kernel ran
    ctx.set_option(_lib.OPT_CTA_GROUP, 2)


@pytest.mark.parametrize("cta_group", [1, 2])
def test_golden_driver_vectors(ctx, golden, cta_group):
    g = golden("driver_64x4096.npz")
    ctx.set_option(_lib.OPT_CTA_GROUP, cta_group)
    exp = orc.synthetic_experimental(64, (60, 60), seed=1)
    dic = orc.synthetic_dictionary(4096, (60, 60), seed=2)
    idx, sc = ctx.dictionary_indexing(exp, 64, dic, 4096, _lib.KDI_NCC, 20)
    _check(g["idx"], g["scores"], idx, sc)
    pexp, j = orc.planted_experimental(dic, 64, seed=3)
    idx, sc = ctx.dictionary_indexing(pexp, 64, dic, 4096, _lib.KDI_NCC, 20)
    assert np.array_equal(idx[:, 0], j)
    _check(g["planted_idx"], orc.dictionary_indexing(pexp, dic, keep_n=20)[1], idx, sc)
    ctx.set_option(_lib.OPT_CTA_GROUP, 2)


# ---- random workloads: metrics, masks, keep_n, operand types, schedules ----------------------------

CASES = [
    # M, N, sig, keep_n, metric, signal mask, nav mask, options
    (256, 4096, (60, 60), 20, "ncc", False, False, {}),
    (256, 4096, (60, 60), 20, "ndp", False, False, {}),
    (300, 5000, (60, 60), 50, "ncc", False, False, {}),
    (200, 3000, (60, 60), 1, "ncc", True, True, {}),
    (130, 2500, (48, 40), 24, "ndp", True, False, {}),
    (200, 3000, (60, 60), 20, "ncc", False, False, {"bf16": 1}),
    (300, 5000, (60, 60), 50, "ncc", False, False, {"cg": 2}),
    (257, 4097, (31, 29), 7, "ncc", False, True, {"cg": 2}),
    (700, 2100, (20, 20), 20, "ncc", False, False, {"strip": 1, "sb": 1}),
    (700, 2100, (20, 20), 20, "ncc", False, False, {"strip": 3, "sb": 2, "cg": 2}),
    (9, 20, (60, 60), 20, "ncc", False, False, {}),          # dictionary smaller than the candidate list
    (33, 100, (12, 12), 60, "ncc", False, False, {}),        # keep_n beyond the fused path -> exact path
    (64, 4096, (60, 60), 20, "ncc", False, False, {"exact": 1}),
    (1, 1, (3, 3), 1, "ncc", False, False, {}),               # one pattern, one dictionary entry
    (3, 2, (5, 7), 2, "ndp", False, False, {}),               # keep_n equal to the dictionary size
    (7, 300, (240, 240), 10, "ncc", True, False, {"cg": 2}),  # 57 600-pixel detector, masked (generic normalise path)
    (300, 5000, (60, 60), 20, "ncc", True, True, {"cg": 2, "split": 0}),  # selection inside the rescoring kernel
]


@pytest.mark.parametrize("case", CASES, ids=[str(c[:5]) + str(c[7]) for c in CASES])
def test_random_workloads(ctx, case):
    M, N, sig, k, metric, smask, nmask, opt = case
    ctx.set_option(_lib.OPT_CTA_GROUP, opt.get("cg", 1))  # cases without "cg" exercise the single-CTA kernel
    ctx.set_option(_lib.OPT_COMPUTE_DTYPE, opt.get("bf16", 0))
    ctx.set_option(_lib.OPT_FORCE_EXACT, opt.get("exact", 0))
    ctx.set_option(_lib.OPT_STRIP_TILES, opt.get("strip", 0))
    ctx.set_option(_lib.OPT_SUPERBLOCK, opt.get("sb", 0))
    ctx.set_option(_lib.OPT_SPLIT_SELECT, opt.get("split", 1))
    try:
        exp = orc.synthetic_experimental(M, sig, seed=1)
        dic = orc.synthetic_dictionary(N, sig, seed=2)
        before = exp.copy(), dic.copy()
        sm = orc.circular_signal_mask(sig) if smask else None
        nm = (np.random.default_rng(5).random(M) < 0.3) if nmask else None
        ctx.set_signal_mask(sm)
        idx, sc = ctx.dictionary_indexing(exp, M, dic, N, _lib.KDI_NCC if metric == "ncc" else _lib.KDI_NDP,
                                          k, nav_mask=nm)
        ridx, rsc = orc.dictionary_indexing(exp, dic, metric=metric, keep_n=k, navigation_mask=nm, signal_mask=sm,
                                            n_experimental_patterns=M)
        _check(ridx, rsc, idx, sc, tie_tol=2e-5 if smask else 1e-6)
        # inputs are never modified (reference tests/test_indexing/test_dictionary_indexing.py:41-43)
        assert np.array_equal(exp, before[0]) and np.array_equal(dic, before[1])
    finally:
        for o in (_lib.OPT_COMPUTE_DTYPE, _lib.OPT_FORCE_EXACT, _lib.OPT_STRIP_TILES, _lib.OPT_SUPERBLOCK):
            ctx.set_option(o, 0)
        ctx.set_option(_lib.OPT_CTA_GROUP, 2)
        ctx.set_option(_lib.OPT_SPLIT_SELECT, 1)
        ctx.set_signal_mask(None)


@pytest.mark.parametrize("src_dtype", [np.uint16, np.float32, np.float64])
def test_source_dtypes_host_and_device(ctx, src_dtype):
    """Experimental patterns of every accepted source type, from host and from device memory, give
    the result of the reference's cast-to-float32-first arithmetic."""
    import torch

    rng = np.random.default_rng(8)
    exp = (rng.random((70, 24, 24)) * 1000).astype(src_dtype)
    dic = orc.synthetic_dictionary(1500, (24, 24), seed=2)
    ridx, rsc = orc.dictionary_indexing(exp.astype(np.float32), dic, keep_n=12, n_experimental_patterns=70)
    idx, sc = ctx.dictionary_indexing(exp, 70, dic, 1500, _lib.KDI_NCC, 12)
    _check(ridx, rsc, idx, sc)
    if src_dtype != np.uint16:
        i2 = torch.empty((70, 12), dtype=torch.int64, device="cuda"); s2 = torch.empty((70, 12), dtype=torch.float32, device="cuda")
        ctx.dictionary_indexing(torch.from_numpy(exp).cuda(), 70, torch.from_numpy(dic).cuda(), 1500, _lib.KDI_NCC, 12, out=(i2, s2))
        assert np.array_equal(i2.cpu().numpy(), idx) and np.array_equal(s2.cpu().numpy(), sc)


def test_constant_pattern_gives_nan_row_without_disturbing_others(ctx):
    """A constant experimental pattern has zero variance: the reference divides 0 by 0 and that
    row's scores are NaN (its index order is then arbitrary).  The other rows must be unaffected
    and the call must not hang."""
    exp = orc.synthetic_experimental(40, (20, 20), seed=1)
    exp[7] = 128
    dic = orc.synthetic_dictionary(900, (20, 20), seed=2)
    idx, sc = ctx.dictionary_indexing(exp, 40, dic, 900, _lib.KDI_NCC, 10)
    assert np.all(np.isnan(sc[7]))
    keep = np.arange(40) != 7
    with np.errstate(invalid="ignore", divide="ignore"):
        ridx, rsc = orc.dictionary_indexing(exp, dic, keep_n=10, n_experimental_patterns=40)
    _check(ridx[keep], rsc[keep], idx[keep], sc[keep])
    assert idx.min() >= 0 and idx.max() < 900


def test_fused_and_exact_paths_agree_bit_for_bit(ctx):
    exp = orc.synthetic_experimental(128, (40, 40), seed=1)
    dic = orc.synthetic_dictionary(6000, (40, 40), seed=2)
    i1, s1 = ctx.dictionary_indexing(exp, 128, dic, 6000, _lib.KDI_NCC, 20)
    ctx.set_option(_lib.OPT_FORCE_EXACT, 1)
    i2, s2 = ctx.dictionary_indexing(exp, 128, dic, 6000, _lib.KDI_NCC, 20)
    ctx.set_option(_lib.OPT_FORCE_EXACT, 0)
    assert np.array_equal(i1, i2) and np.array_equal(s1, s2)


@pytest.mark.parametrize("cta_group", [1, 2])
def test_overlapped_schedule_equals_serial(ctx, cta_group):
    """The overlapped schedule (dictionary normalisation on the auxiliary stream, one tensor-core
    launch per row-block group on alternating streams, rescoring of finished groups beside the
    next launch) must give bit-identical results to one-kernel-at-a-time execution - for the full
    pipeline, for pre-normalised pattern sets (kdi_match_topk) and for the candidate stage of the
    sharded pipeline."""
    import torch

    M, N, sig, k = 1500, 7000, (30, 30), 20
    exp = torch.from_numpy(orc.synthetic_experimental(M, sig, seed=5)).cuda()
    dic = torch.from_numpy(orc.synthetic_dictionary(N, sig, seed=6)).cuda()
    ctx.set_option(_lib.OPT_CTA_GROUP, cta_group)
    ctx.set_option(_lib.OPT_SUPERBLOCK, 2)
    ctx.set_option(_lib.OPT_STRIP_TILES, 3)
    out = {}
    try:
        for mode in (0, 2):
            ctx.set_option(_lib.OPT_OVERLAP, mode)
            idx = torch.empty((M, k), dtype=torch.int64, device="cuda")
            sc = torch.empty((M, k), dtype=torch.float32, device="cuda")
            ctx.dictionary_indexing(exp, M, dic, N, _lib.KDI_NCC, k, out=(idx, sc))
            launches = ctx.timings()["gemm_launches"]
            e = ctx.patterns(exp, M, _lib.KDI_NCC)
            d = ctx.patterns(dic, N, _lib.KDI_NCC)
            i2, s2 = ctx.match_topk(e, d, k)
            shard, approx, gidx = ctx.shard_candidates(exp, M, dic, N, _lib.KDI_NCC, k, index_offset=100)
            out[mode] = (idx.cpu().numpy(), sc.cpu().numpy(), i2, s2, approx.cpu().numpy(), gidx.cpu().numpy(), launches)
            shard.close(); e.close(); d.close()
    finally:
        ctx.set_option(_lib.OPT_OVERLAP, 1)
        ctx.set_option(_lib.OPT_SUPERBLOCK, 0)
        ctx.set_option(_lib.OPT_STRIP_TILES, 0)
        ctx.set_option(_lib.OPT_CTA_GROUP, 2)
    assert out[0][6] == 1 and out[2][6] >= 3  # one launch vs one per row-block group (+ a first slice)
    for a, b in zip(out[0][:4], out[2][:4]):
        assert np.array_equal(a, b)
    assert np.array_equal(out[0][0], out[0][2]) and np.array_equal(out[0][1], out[0][3])
    # candidate lists: the tensor-core scores do not depend on the schedule (an exact tie at the
    # kc-th place may keep either index)
    assert np.array_equal(out[0][4], out[2][4]) and np.mean(out[0][5] == out[2][5]) > 0.999
    ridx, rsc = orc.dictionary_indexing(exp.cpu().numpy(), dic.cpu().numpy(), keep_n=k)
    _check(ridx, rsc, out[2][0], out[2][1])


SCHEDULES = {
    # option settings of the overlapped schedule; every one must reproduce serial execution bit for bit
    "events_only": {_lib.OPT_DEP_FLAGS: 0},
    "flags": {_lib.OPT_DEP_FLAGS: 1},
    "flags_groups4_post": {_lib.OPT_DEP_FLAGS: 1, _lib.OPT_MIN_GROUPS: 4, _lib.OPT_POST_PER_GROUP: 1},
    "flags_sms_serial_post": {_lib.OPT_DEP_FLAGS: 1, _lib.OPT_MIN_GROUPS: 3, _lib.OPT_POST_PER_GROUP: 1,
                              _lib.OPT_GEMM_SMS: 132, _lib.OPT_GEMM_SERIAL: 1},
    "partition16_post": {_lib.OPT_DEP_FLAGS: 1, _lib.OPT_MIN_GROUPS: 4, _lib.OPT_POST_PER_GROUP: 1,
                         _lib.OPT_SM_PARTITION: 16},
}


@pytest.mark.parametrize("name", list(SCHEDULES))
@pytest.mark.parametrize("src", ["f32", "u8dict", "masked"])
def test_schedule_variants_equal_serial(ctx, name, src):
    """Device-side readiness counters (the tensor-core kernel's TMA producer waits per 256-row tile for
    the dictionary normalise running beside it), more row-block groups, per-group post-processing on
    the post stream, a reduced GEMM grid and an SM partition change WHEN things run, never WHAT comes
    out: every variant must equal one-kernel-at-a-time execution bit for bit.  ``masked`` and the
    shared-memory normalise kernels take the event-ordered schedule instead of the counters."""
    import torch

    M, N, sig, k = 1800, 9000, (32, 32), 20
    exp = torch.from_numpy(orc.synthetic_experimental(M, sig, seed=15)).cuda()
    dic_np = orc.synthetic_dictionary(N, sig, seed=16)
    if src == "u8dict":
        dic_np = np.round(dic_np * 255).astype(np.uint8)
    dic = torch.from_numpy(dic_np).cuda()
    smask = orc.circular_signal_mask(sig) if src == "masked" else None
    ctx.set_signal_mask(smask)
    ctx.set_option(_lib.OPT_SUPERBLOCK, 2)
    ctx.set_option(_lib.OPT_STRIP_TILES, 3)
    res = {}
    try:
        for mode in ("serial", name):
            if mode == "serial":
                ctx.set_option(_lib.OPT_OVERLAP, 0)
            else:
                ctx.set_option(_lib.OPT_OVERLAP, 2)
                try:
                    for o, v in SCHEDULES[name].items():
                        ctx.set_option(o, v)
                except NotImplementedError as e:  # no green-context support in this driver
                    pytest.skip(str(e))
            idx = torch.empty((M, k), dtype=torch.int64, device="cuda")
            sc = torch.empty((M, k), dtype=torch.float32, device="cuda")
            for _ in range(3):  # repeated: a race would not show up every time
                idx.fill_(-7); sc.fill_(-7)
                ctx.dictionary_indexing(exp, M, dic, N, _lib.KDI_NCC, k, out=(idx, sc))
                got = (idx.cpu().numpy(), sc.cpu().numpy())
                if mode in res:
                    assert np.array_equal(res[mode][0], got[0]) and np.array_equal(res[mode][1], got[1])
                res[mode] = got
    finally:
        for o in (_lib.OPT_SUPERBLOCK, _lib.OPT_STRIP_TILES, _lib.OPT_GEMM_SMS, _lib.OPT_MIN_GROUPS,
                  _lib.OPT_POST_PER_GROUP, _lib.OPT_GEMM_SERIAL, _lib.OPT_SM_PARTITION):
            ctx.set_option(o, 0)
        ctx.set_option(_lib.OPT_DEP_FLAGS, 0)
        ctx.set_option(_lib.OPT_OVERLAP, 1)
        ctx.set_signal_mask(None)
    assert np.array_equal(res["serial"][0], res[name][0]) and np.array_equal(res["serial"][1], res[name][1])
    ridx, rsc = orc.dictionary_indexing(exp.cpu().numpy(), dic_np, keep_n=k, signal_mask=smask)
    _check(ridx, rsc, res[name][0], res[name][1], tie_tol=2e-5 if smask is not None else 1e-6)


@pytest.mark.parametrize("metric", ["ncc", "ndp"])
@pytest.mark.parametrize("compute", ["fp16", "bf16"])
def test_random_full_dictionary_fused_equals_exact(ctx, metric, compute):
    """RANDOM (not planted) patterns against the full 100 000-entry dictionary of BASELINE configs[1]:
    the fused pipeline (16-bit tensor-core candidates -> exact rescoring -> certificate, flagged rows
    through the exact path) must return exactly what the exact float32 path returns for every row -
    indices and scores bit for bit, i.e. zero unflagged differences.  Near-ties are as frequent here
    as random data makes them (the 49th-50th gap has a median of 6.7e-5)."""
    import torch

    M, N, sig, k = 2000, 100_000, (60, 60), 20
    g = torch.Generator(device="cuda"); g.manual_seed(21)
    exp = torch.randint(0, 256, (M,) + sig, dtype=torch.uint8, device="cuda", generator=g)
    dic = torch.rand((N,) + sig, dtype=torch.float32, device="cuda", generator=g)
    code = _lib.KDI_NCC if metric == "ncc" else _lib.KDI_NDP
    ctx.set_option(_lib.OPT_COMPUTE_DTYPE, 1 if compute == "bf16" else 0)
    out = {}
    try:
        for force in (0, 1):
            ctx.set_option(_lib.OPT_FORCE_EXACT, force)
            idx = torch.empty((M, k), dtype=torch.int64, device="cuda")
            sc = torch.empty((M, k), dtype=torch.float32, device="cuda")
            ctx.dictionary_indexing(exp, M, dic, N, code, k, out=(idx, sc))
            out[force] = (idx.cpu().numpy(), sc.cpu().numpy(), ctx.timings())
    finally:
        ctx.set_option(_lib.OPT_FORCE_EXACT, 0)
        ctx.set_option(_lib.OPT_COMPUTE_DTYPE, 0)
    assert out[0][2]["gemm_launches"] >= 1 and out[1][2]["gemm_launches"] == 0
    diff_rows = np.flatnonzero((out[0][0] != out[1][0]).any(axis=1) | (out[0][1] != out[1][1]).any(axis=1))
    assert diff_rows.size == 0, (diff_rows[:10], out[0][2]["flagged_rows"])
    # the certificate should rarely fire on this data (it costs an exact pass per flagged row)
    assert out[0][2]["flagged_rows"] <= (M // 4 if compute == "bf16" else M // 20), out[0][2]["flagged_rows"]
    # and a sample of rows against the CPU oracle (the reference's arithmetic)
    rows = np.arange(0, M, 125)
    ridx, rsc = orc.dictionary_indexing(exp[rows].cpu().numpy(), dic.cpu().numpy(), metric=metric, keep_n=k)
    _check(ridx, rsc, out[0][0][rows], out[0][1][rows])


def test_duplicate_dictionary_rows_fall_back_to_exact(ctx):
    """Exact ties across the candidate boundary: the certificate must flag the rows and the exact
    path must order ties by ascending index."""
    dic = orc.synthetic_dictionary(400, (20, 20), seed=2)
    dic[100:180] = dic[7]  # 81 identical rows
    exp = np.clip(np.rint(dic[[7, 9]] * 255), 0, 255).astype(np.uint8)
    idx, sc = ctx.dictionary_indexing(exp, 2, dic, 400, _lib.KDI_NCC, 20)
    assert ctx.timings()["flagged_rows"] >= 1
    assert list(idx[0]) == [7] + list(range(100, 119))
    assert np.all(sc[0] == sc[0, 0])
    assert idx[1, 0] == 9


def test_index_offset_and_device_outputs(ctx):
    import torch

    exp = orc.synthetic_experimental(50, (20, 20), seed=1)
    dic = orc.synthetic_dictionary(1500, (20, 20), seed=2)
    ridx, rsc = orc.dictionary_indexing(exp, dic, keep_n=10)
    idx = torch.empty((50, 10), dtype=torch.int64, device="cuda")
    sc = torch.empty((50, 10), dtype=torch.float32, device="cuda")
    ctx.dictionary_indexing(torch.from_numpy(exp).cuda(), 50, torch.from_numpy(dic).cuda(), 1500, _lib.KDI_NCC, 10,
                            index_offset=1000, out=(idx, sc))
    _check(ridx + 1000, rsc, idx.cpu().numpy(), sc.cpu().numpy())


# ---- reference test-suite semantics (tests/test_indexing/test_dictionary_indexing.py) ---------------

def test_self_dictionary_identity(dummy_array):
    dic = dummy_array.reshape(-1, 3, 3)
    before = dummy_array.copy()
    res = kb.dictionary_indexing(dummy_array, dic, metric="ndp", rechunk=True, verbose=False)  # :27-43
    assert np.allclose(res.scores[:, 0], 1)
    assert np.array_equal(res.simulation_indices[:, 0], np.arange(9))
    assert res.scores.shape == (9, 9)  # keep_n clipped to the dictionary size (:67)
    assert np.array_equal(dummy_array, before)
    smask = np.array([[0, 0, 0], [0, 1, 0], [0, 0, 0]], dtype=bool)
    res = kb.dictionary_indexing(dummy_array, dic, n_per_iteration=2, signal_mask=smask, rechunk=True, verbose=False)
    assert np.allclose(res.scores[:, 0], 1) and res.simulation_indices.dtype == np.int64  # :45-66
    ridx, rsc = orc.dictionary_indexing(dummy_array, dic, signal_mask=smask)
    assert np.max(np.abs(res.scores - rsc)) < 1e-5


@pytest.mark.parametrize("nav_slice, nav_shape", [((0, 0), ()), ((0, slice(0, 1)), (1,)), ((0, slice(0, 3)), (3,)),
                                                   ((slice(0, 3), slice(0, 2)), (3, 2))])
def test_navigation_shapes(dummy_array, nav_slice, nav_shape):
    """0-D / 1-D / 2-D navigation (reference :119-145)."""
    exp = dummy_array[nav_slice]
    dic = dummy_array.reshape(-1, 3, 3)
    res = kb.dictionary_indexing(exp, dic, keep_n=3, verbose=False)
    assert res.shape == nav_shape
    n = int(np.prod(nav_shape)) if nav_shape else 1
    assert res.scores.shape == (n, 3) and res.size == n and res.scan_unit == "px"
    ridx, rsc = orc.dictionary_indexing(exp, dic, keep_n=3)
    assert np.max(np.abs(res.scores - rsc)) < 1e-5
    assert np.array_equal(res.simulation_indices[:, 0], ridx[:, 0])


def test_navigation_mask_semantics(dummy_array):
    """reference :166-180 and _dictionary_indexing.py:142-163."""
    dic = dummy_array.reshape(-1, 3, 3)
    nav = np.array([[0, 0, 0], [0, 1, 0], [0, 0, 0]], dtype=bool)
    res = kb.dictionary_indexing(dummy_array, dic, keep_n=1, navigation_mask=nav, verbose=False)
    assert res.size == 8 and res.rotations_per_point == 1 and res.shape == (3, 3)
    assert res.prop["scores"].shape == (9,)  # squeezed because keep_n == 1 with a navigation mask
    assert np.array_equal(res.is_in_data, ~nav.ravel())
    assert np.allclose(res.scores, 1)
    res = kb.dictionary_indexing(dummy_array, dic, metric="ndp", navigation_mask=~nav, verbose=False,
                                 dictionary_rotations=np.tile([1.0, 0, 0, 0], (9, 1)))
    assert res.size == 1 and res.scores.shape == (1, 9) and res.rotations.shape == (9, 9, 4)
    res = kb.dictionary_indexing(dummy_array, dic, keep_n=1, verbose=False)
    assert res.scores.shape == (9, 1)  # no squeeze without a navigation mask (:164-166)


def test_info_and_speed_lines(dummy_array, capsys):
    dic = dummy_array.reshape(-1, 3, 3)
    kb.dictionary_indexing(dummy_array, dic, phase_name="ni")
    out = capsys.readouterr().out
    assert "Dictionary indexing information:\n  Phase name: ni\n  Matching 9 experimental pattern(s) to 9 dictionary pattern(s)\n" in out
    assert "NormalizedCrossCorrelationMetric: float32, greater is better" in out
    assert "  Indexing speed: " in out and "comparisons/s" in out


# ---- merge and orientation similarity map ------------------------------------------------------------

def test_merge_topk_matches_numpy(ctx):
    rng = np.random.default_rng(3)
    L, R, K = 8, 500, 50
    sc = np.sort(rng.random((L, R, K), dtype=np.float32), axis=2)[:, :, ::-1].copy()
    idx = rng.permutation(L * R * K).reshape(L, R, K).astype(np.int64)
    io, so = ctx.merge_topk(sc, idx, 50)
    alls = np.concatenate(list(sc), axis=1)
    alli = np.concatenate(list(idx), axis=1)
    best = np.argsort(-alls, axis=1, kind="stable")[:, :50]
    assert np.array_equal(so, np.take_along_axis(alls, best, 1))
    r = orc.compare_topk(np.take_along_axis(alli, best, 1), np.take_along_axis(alls, best, 1), io, so, tie_tol=0)
    assert r["tie_ok"]
    # ties are ordered by ascending index
    sc2 = np.zeros((2, 3, 4), np.float32)
    idx2 = np.arange(24, dtype=np.int64)[::-1].reshape(2, 3, 4).copy()
    io, so = ctx.merge_topk(sc2, idx2, 5)
    for r_ in range(3):
        pool = np.sort(np.concatenate([idx2[0, r_], idx2[1, r_]]))
        assert list(io[r_]) == list(pool[:5])


def test_osm_goldens(ctx, golden):
    g = golden("osm.npz")

    class X:
        def __init__(self, idx, shape):
            self.prop = {"simulation_indices": idx}
            self.shape = shape

    assert np.array_equal(kb.orientation_similarity_map(X(g["idx34"], (3, 4))), g["osm34"])
    assert np.array_equal(kb.orientation_similarity_map(X(g["idx34"], (3, 4)), n_best=2, normalize=True),
                          g["osm34_norm_n2"])
    o = kb.orientation_similarity_map(X(g["idx34"], (3, 4)), from_n_best=1)
    assert o.shape == (3, 4, 3) and np.array_equal(o, g["osm34_from1"])
    idx = g["idx_17x23"]
    assert np.array_equal(kb.orientation_similarity_map(X(idx, (17, 23))), g["osm_17x23"])
    assert np.array_equal(kb.orientation_similarity_map(X(idx, (17, 23)), n_best=7, normalize=True),
                          g["osm_17x23_n7_norm"])
    assert np.array_equal(kb.orientation_similarity_map(X(idx, (17, 23)), footprint=g["footprint8"], center_index=4),
                          g["osm_17x23_fp8"])
    # reference tests/test_indexing/test_orientation_similarity_map.py:27-64
    assert np.allclose(kb.orientation_similarity_map(X(g["idx_tiled"], (10, 10))), 5)
    assert np.allclose(kb.orientation_similarity_map(X(g["idx_tiled"], (10, 10)), normalize=True), 1)
    with pytest.raises(ValueError, match="n_best 6 cannot be greater than keep_n 5"):
        kb.orientation_similarity_map(X(g["idx_tiled"], (10, 10)), n_best=6)
    assert kb.orientation_similarity_map(X(g["idx_tiled"], (10, 10)), n_best=5, from_n_best=2).shape == (10, 10, 4)
    assert np.isnan(kb.orientation_similarity_map(X(g["idx34"][:1], (1, 1))))


# ---- BASELINE sizes through size-independent properties --------------------------------------------

def test_config2_full_size_properties(ctx):
    """10 000 x 100 000 (BASELINE configs[1]) on device-generated planted data: the planted row
    must be the best match of every pattern, lists are sorted, indices are valid and unique, and
    the best score equals an independent exact evaluation of that pair."""
    import torch

    M, N, sig, k = 10_000, 100_000, (60, 60), 20
    g = torch.Generator(device="cuda"); g.manual_seed(7)
    dic = torch.rand((N,) + sig, device="cuda", generator=g)
    j = torch.randint(0, N, (M,), device="cuda", generator=g)
    noise = torch.rand((M,) + sig, device="cuda", generator=g)
    exp = torch.clamp(torch.round(255.0 * (0.7 * dic[j] + 0.3 * noise)), 0, 255).to(torch.uint8)
    idx = torch.empty((M, k), dtype=torch.int64, device="cuda")
    sc = torch.empty((M, k), dtype=torch.float32, device="cuda")
    for cg in (1, 2):
        ctx.set_option(_lib.OPT_CTA_GROUP, cg)
        ctx.dictionary_indexing(exp, M, dic, N, _lib.KDI_NCC, k, out=(idx, sc))
        assert ctx.timings()["gemm_launches"] >= 1
        assert torch.equal(idx[:, 0], j)
        assert bool((sc[:, :-1] >= sc[:, 1:]).all())
        assert int(idx.min()) >= 0 and int(idx.max()) < N
        srt = torch.sort(idx, dim=1).values
        assert bool((srt[:, 1:] != srt[:, :-1]).all())
        # independent check of the best score on a sample of rows (torch float64)
        rows = torch.arange(0, M, 97, device="cuda")
        e = exp[rows].double().flatten(1); e = e - e.mean(1, keepdim=True); e = e / e.norm(dim=1, keepdim=True)
        d = dic[j[rows]].double().flatten(1); d = d - d.mean(1, keepdim=True); d = d / d.norm(dim=1, keepdim=True)
        assert float(((e * d).sum(1) - sc[rows, 0].double()).abs().max()) < 1e-5
    ctx.set_option(_lib.OPT_CTA_GROUP, 2)


# ---- the other BASELINE configurations ---------------------------------------------------------------

def test_config3_shape_small(ctx):
    """configs[2] at oracle size: 120x120 patterns, circular signal mask (11 287 of 14 400 pixels
    kept -> K padded to 11 328), NormalizedDotProductMetric, keep_n 20."""
    sig = (120, 120)
    sm = orc.circular_signal_mask(sig)
    assert int((~sm).sum()) == 11287
    exp = orc.synthetic_experimental(64, sig, seed=1)
    dic = orc.synthetic_dictionary(3000, sig, seed=2)
    res = kb.dictionary_indexing(exp, dic, metric="ndp", keep_n=20, signal_mask=sm, verbose=False)
    ridx, rsc = orc.dictionary_indexing(exp, dic, metric="ndp", keep_n=20, signal_mask=sm)
    _check(ridx, rsc, res.simulation_indices, res.scores, tie_tol=2e-5)
    pexp, j = orc.planted_experimental(dic, 64, seed=3)
    res = kb.dictionary_indexing(pexp, dic, metric="ndp", keep_n=20, signal_mask=sm, verbose=False)
    assert np.array_equal(res.simulation_indices[:, 0], j)


def test_config5_shape_small_with_osm(ctx):
    """configs[4] at oracle size: 80x80 patterns (K = 6 400, no padding), bf16 operands, a 2-D map
    and its orientation similarity map."""
    sig = (80, 80)
    dic = orc.synthetic_dictionary(5000, sig, seed=2)
    # a smooth map: neighbouring points are noisy copies of nearby dictionary rows
    rng = np.random.default_rng(9)
    base = (np.add.outer(np.arange(20), np.arange(20)) * 7) % 5000
    noise = rng.random((20, 20) + sig, dtype=np.float32)
    exp = np.clip(np.rint(255 * (0.6 * dic[base] + 0.4 * noise)), 0, 255).astype(np.uint8)
    ctx.set_option(_lib.OPT_COMPUTE_DTYPE, 1)
    try:
        res = kb.dictionary_indexing(exp, dic, metric="ncc", keep_n=20, verbose=False)
    finally:
        ctx.set_option(_lib.OPT_COMPUTE_DTYPE, 0)
    ridx, rsc = orc.dictionary_indexing(exp, dic, metric="ncc", keep_n=20)
    r = _check(ridx, rsc, res.simulation_indices, res.scores)
    assert np.array_equal(res.simulation_indices[:, 0], base.ravel())
    osm = kb.orientation_similarity_map(res)
    assert osm.shape == (20, 20) and osm.dtype == np.float32
    assert np.array_equal(osm, orc.orientation_similarity_map(res.simulation_indices, (20, 20)))
    if r["exact_rows"] == 1.0:
        assert np.array_equal(osm, orc.orientation_similarity_map(ridx, (20, 20)))
    osm_n = kb.orientation_similarity_map(res, n_best=10, normalize=True)
    assert np.array_equal(osm_n, orc.orientation_similarity_map(res.simulation_indices, (20, 20), n_best=10, normalize=True))


def _planted_device(M, N, sig, seed):
    import torch

    g = torch.Generator(device="cuda"); g.manual_seed(seed)
    dic = torch.rand((N,) + sig, device="cuda", generator=g)
    j = torch.randint(0, N, (M,), device="cuda", generator=g)
    exp = torch.empty((M,) + sig, dtype=torch.uint8, device="cuda")
    for a in range(0, M, 4096):  # bounded temporaries
        b = min(a + 4096, M)
        noise = torch.rand((b - a,) + sig, device="cuda", generator=g)
        exp[a:b] = torch.clamp(torch.round(255.0 * (0.7 * dic[j[a:b]] + 0.3 * noise)), 0, 255).to(torch.uint8)
    return exp, dic, j


def _full_size_properties(ctx, M, N, sig, k, metric, smask):
    import torch

    exp, dic, j = _planted_device(M, N, sig, seed=11)
    idx = torch.empty((M, k), dtype=torch.int64, device="cuda")
    sc = torch.empty((M, k), dtype=torch.float32, device="cuda")
    ctx.set_signal_mask(smask)
    try:
        ctx.dictionary_indexing(exp, M, dic, N, _lib.KDI_NCC if metric == "ncc" else _lib.KDI_NDP, k, out=(idx, sc))
    finally:
        ctx.set_signal_mask(None)
    tm = ctx.timings()
    assert tm["gemm_launches"] >= 1
    assert torch.equal(idx[:, 0], j)
    assert bool((sc[:, :-1] >= sc[:, 1:]).all())
    assert int(idx.min()) >= 0 and int(idx.max()) < N
    srt = torch.sort(idx, dim=1).values
    assert bool((srt[:, 1:] != srt[:, :-1]).all())
    rows = torch.arange(0, M, 211, device="cuda")
    keep = torch.ones(sig[0] * sig[1], dtype=torch.bool, device="cuda") if smask is None else \
        torch.from_numpy(~smask.ravel()).cuda()
    e = exp[rows].double().flatten(1)[:, keep]
    d = dic[j[rows]].double().flatten(1)[:, keep]
    if metric == "ncc":
        e = e - e.mean(1, keepdim=True); d = d - d.mean(1, keepdim=True)
    e = e / e.norm(dim=1, keepdim=True); d = d / d.norm(dim=1, keepdim=True)
    assert float(((e * d).sum(1) - sc[rows, 0].double()).abs().max()) < 1e-5
    return tm


def test_config3_full_size_properties(ctx):
    """40 000 x 100 000, 120x120, circular mask, NDP (BASELINE configs[2])."""
    tm = _full_size_properties(ctx, 40_000, 100_000, (120, 120), 20, "ndp", orc.circular_signal_mask((120, 120)))
    assert tm["flagged_rows"] < 400


def test_config4_shard_full_size_properties(ctx):
    """One rank's share of BASELINE configs[3]: 100 000 patterns vs a 37 500-row shard, keep_n 50."""
    tm = _full_size_properties(ctx, 100_000, 37_500, (60, 60), 50, "ncc", None)
    assert tm["flagged_rows"] < 1000


def test_reference_chunk_loop_over_plugin_hooks(ctx):
    """Level-1 integration: the reference's own chunk loop (_dictionary_indexing.py:94-128),
    written here exactly as the reference has it, driving the GPU metric through the three plugin
    hooks + argtopk/topk - what an unmodified kikuchipy does with ``metric=<GPU metric>``."""
    exp = orc.synthetic_experimental(96, (40, 40), seed=1).reshape(8, 12, 40, 40)
    dic = orc.synthetic_dictionary(2500, (40, 40), seed=2)
    nav = np.random.default_rng(2).random((8, 12)) < 0.25
    sm = orc.circular_signal_mask((40, 40))
    metric = kb.NormalizedCrossCorrelationMetric(96, 2500, navigation_mask=nav, signal_mask=sm)
    keep_n, n_per_iteration = 20, 700
    experimental = metric.prepare_experimental(exp)
    dictionary = dic.reshape((2500, -1))
    n_experimental = experimental.shape[0]
    n_iterations = int(np.ceil(2500 / n_per_iteration))
    negative_sign = -metric.sign
    simulation_indices = np.zeros((n_experimental, keep_n), dtype=np.int32)
    scores = np.full((n_experimental, keep_n), negative_sign, dtype=metric.dtype)
    chunk_starts = np.cumsum([0] + [n_per_iteration] * (n_iterations - 1))
    chunk_ends = np.cumsum([n_per_iteration] * n_iterations)
    chunk_ends[-1] = max(chunk_ends[-1], 2500)
    for start, end in zip(chunk_starts, chunk_ends):
        simulated = metric.prepare_dictionary(dictionary[start:end])
        similarities = metric.match(experimental, simulated)
        k = min(keep_n, end - start)
        idx_i = similarities.argtopk(k, axis=-1).reshape((-1, k)) + start
        sc_i = similarities.topk(k, axis=-1).reshape((-1, k))
        all_scores = np.hstack((scores, sc_i))
        all_idx = np.hstack((simulation_indices, idx_i))
        best = np.argsort(negative_sign * all_scores, axis=1)[:, :keep_n]
        scores = np.take_along_axis(all_scores, best, axis=1)
        simulation_indices = np.take_along_axis(all_idx, best, axis=1)
    ridx, rsc = orc.dictionary_indexing(exp, dic, keep_n=keep_n, n_per_iteration=n_per_iteration,
                                        navigation_mask=nav, signal_mask=sm)
    _check(ridx, rsc, simulation_indices.astype(np.int64), scores, tie_tol=2e-5)
    # and the one-call driver gives the same answer
    res = kb.dictionary_indexing(exp, dic, keep_n=keep_n, n_per_iteration=n_per_iteration, navigation_mask=nav,
                                 signal_mask=sm, verbose=False)
    assert np.array_equal(res.scores, scores) and np.array_equal(res.simulation_indices, simulation_indices)


class _LazyDictionary:
    """Dask-like stand-in: slices stay lazy, ``compute()`` materialises and logs the chunk length."""

    def __init__(self, a, log):
        self._a, self.log = a, log
        self.shape, self.chunksize = a.shape, (1000,) + a.shape[1:]

    def __getitem__(self, key):
        return _LazyDictionary(self._a[key], self.log)

    def compute(self):
        self.log.append(self._a.shape[0])
        return self._a


@pytest.mark.parametrize("n_per_iteration", [None, 700, 4096])
def test_lazy_dictionary_is_streamed_chunk_by_chunk(ctx, n_per_iteration):
    """A lazy dictionary is computed one chunk at a time inside the loop (the reference:
    _dictionary_indexing.py:105-108) and fed to the appendable job (kdi_job_*); it is never
    materialised as a whole, and the result equals the one-call driver's on the full array."""
    exp = orc.synthetic_experimental(300, (24, 24), seed=3)
    dic = orc.synthetic_dictionary(5000, (24, 24), seed=4)
    log = []
    res = kb.dictionary_indexing(exp.reshape(15, 20, 24, 24), _LazyDictionary(dic, log), keep_n=12,
                                 n_per_iteration=n_per_iteration, verbose=False)
    per = 1000 if n_per_iteration is None else n_per_iteration
    assert sum(log) == 5000 and max(log) <= per and len(log) == -(-5000 // per)
    idx, sc = ctx.dictionary_indexing(exp, 300, dic, 5000, _lib.KDI_NCC, 12)
    assert np.array_equal(res.simulation_indices, idx) and np.array_equal(res.scores, sc)
    ridx, rsc = orc.dictionary_indexing(exp, dic, keep_n=12, n_per_iteration=per)
    _check(ridx, rsc, res.simulation_indices, res.scores)


def test_job_api_chunks_devices_and_errors(ctx):
    """kdi_job_begin / _append / _finish: ragged chunks from host (pageable and pinned) and device
    memory, device outputs, and the error paths (too many rows, finish before the last chunk)."""
    import torch

    exp = orc.synthetic_experimental(200, (20, 20), seed=5)
    dic = orc.synthetic_dictionary(3000, (20, 20), seed=6)
    want_i, want_s = ctx.dictionary_indexing(exp, 200, dic, 3000, _lib.KDI_NCC, 20, index_offset=7)
    pinned = ctx.pinned_empty((900, 20, 20), np.float32)
    pinned[...] = dic[1100:2000]
    job = ctx.indexing_job(exp, 200, 3000, _lib.KDI_NCC, 20, index_offset=7)
    job.append(dic[:100])                               # pageable
    job.append(torch.from_numpy(dic[100:1100]).cuda())  # device
    job.append(pinned)                                  # pinned
    job.append(dic[2000:])
    io = torch.empty((200, 20), dtype=torch.int64, device="cuda")
    so = torch.empty((200, 20), dtype=torch.float32, device="cuda")
    job.finish(out=(io, so))
    ctx.pinned_free(pinned)
    assert np.array_equal(io.cpu().numpy(), want_i) and np.array_equal(so.cpu().numpy(), want_s)
    job = ctx.indexing_job(exp, 200, 3000, _lib.KDI_NCC, 20)
    job.append(dic[:2000])
    with pytest.raises(ValueError, match="more dictionary rows appended"):
        job.append(dic[:2000])
    job = ctx.indexing_job(exp, 200, 3000, _lib.KDI_NCC, 20)
    job.append(dic[:2000])
    with pytest.raises(ValueError, match="only 2000 of the 3000 announced"):
        job.finish()
    # the context is still healthy
    i2, s2 = ctx.dictionary_indexing(exp, 200, dic, 3000, _lib.KDI_NCC, 20, index_offset=7)
    assert np.array_equal(i2, want_i) and np.array_equal(s2, want_s)


def test_pageable_and_pinned_host_dictionaries_agree(ctx):
    """Pageable host rows go through the library's pinned ring (host threads + DMA), pinned rows
    straight to the DMA engine; several pieces per call (> 64 MB) and identical results."""
    M, N, sig = 64, 12_000, (40, 40)
    exp = orc.synthetic_experimental(M, sig, seed=7)
    dic = orc.synthetic_dictionary(N, sig, seed=8)  # 77 MB: two pieces
    pinned = ctx.pinned_empty(dic.shape, np.float32)
    pinned[...] = dic
    a = ctx.dictionary_indexing(exp, M, dic, N, _lib.KDI_NDP, 5)
    b = ctx.dictionary_indexing(exp, M, pinned, N, _lib.KDI_NDP, 5)
    ctx.pinned_free(pinned)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert ctx.timings()["h2d_bytes"] >= dic.nbytes
    ridx, rsc = orc.dictionary_indexing(exp, dic, metric="ndp", keep_n=5)
    _check(ridx, rsc, a[0], a[1])


@pytest.mark.parametrize("keep_n", [53, 100, 104])
def test_large_keep_n_stays_on_the_tensor_core_path(ctx, keep_n):
    """keep_n up to 104 is nominated by the tensor-core kernel (128-entry candidate lists) and must equal
    the exact float32 path bit for bit; beyond that the exact path serves the call (the reference
    accepts any keep_n <= N, _dictionary_indexing.py:67)."""
    exp = orc.synthetic_experimental(300, (32, 32), seed=31)
    dic = orc.synthetic_dictionary(20_000, (32, 32), seed=32)
    assert ctx.candidate_capacity(keep_n) == 128 and ctx.candidate_capacity(105) == 0
    i1, s1 = ctx.dictionary_indexing(exp, 300, dic, 20_000, _lib.KDI_NCC, keep_n)
    tm = ctx.timings()
    assert tm["gemm_launches"] >= 1
    ctx.set_option(_lib.OPT_FORCE_EXACT, 1)
    try:
        i2, s2 = ctx.dictionary_indexing(exp, 300, dic, 20_000, _lib.KDI_NCC, keep_n)
    finally:
        ctx.set_option(_lib.OPT_FORCE_EXACT, 0)
    assert np.array_equal(i1, i2) and np.array_equal(s1, s2)
    assert tm["flagged_rows"] <= 60, tm["flagged_rows"]
    ridx, rsc = orc.dictionary_indexing(exp[:40], dic, keep_n=keep_n, n_experimental_patterns=40)
    _check(ridx, rsc, i1[:40], s1[:40])


@pytest.mark.parametrize("compute", ["fp16", "bf16"])
def test_adversarial_near_ties_are_caught_by_the_certificate(ctx, compute):
    """300 dictionary rows that differ from one another by ~1e-4 relative noise have exact
    scores ~1e-6 apart - far below what 16-bit operands resolve - and there are more of them than
    candidate slots.  The certificate (noise level floored at the a-priori rounding noise) must send
    every affected row through the exact path: results equal the forced-exact ones bit for bit."""
    rng = np.random.default_rng(41)
    dic = orc.synthetic_dictionary(6000, (30, 30), seed=42)
    base = dic[17].copy()
    dic[1000:1300] = base[None] * (1.0 + 1e-4 * rng.standard_normal((300, 30, 30)).astype(np.float32))
    exp = orc.synthetic_experimental(64, (30, 30), seed=43)
    exp[:32] = np.clip(np.rint(255 * (0.8 * base[None] + 0.2 * rng.random((32, 30, 30)))), 0, 255).astype(np.uint8)
    ctx.set_option(_lib.OPT_COMPUTE_DTYPE, 1 if compute == "bf16" else 0)
    try:
        i1, s1 = ctx.dictionary_indexing(exp, 64, dic, 6000, _lib.KDI_NCC, 20)
        flagged = ctx.timings()["flagged_rows"]
        ctx.set_option(_lib.OPT_FORCE_EXACT, 1)
        i2, s2 = ctx.dictionary_indexing(exp, 64, dic, 6000, _lib.KDI_NCC, 20)
    finally:
        ctx.set_option(_lib.OPT_FORCE_EXACT, 0)
        ctx.set_option(_lib.OPT_COMPUTE_DTYPE, 0)
    assert np.array_equal(i1, i2) and np.array_equal(s1, s2)
    assert flagged >= 32  # the planted rows cannot be certified from 32 candidates
    assert np.all((i1[:32] >= 1000) & (i1[:32] < 1300) | (i1[:32] == 17))


@pytest.mark.parametrize("metric", ["ncc", "ndp"])
def test_strict_certificate_is_a_bound_and_changes_nothing(ctx, metric):
    """KDI_OPT_CERT_STRICT: the certificate uses a worst-case bound on |tensor-core score - float32 score|
    (operand rounding by Cauchy-Schwarz, truncation in every accumulation step) instead of the measured
    error model.  On RANDOM patterns against the 100 000-entry dictionary of BASELINE configs[1]:
    (1) the bound holds on every one of the 128 000 measured (candidate, exact) pairs, fp16 and bf16;
    (2) strict results == default results bit for bit, with few rows sent to the exact path (the lists are
    one size larger; NCC - the scores of uncentred random rows under NDP lie within +-4e-3 of one another, less
    than the bound resolves even over 64 places, so most of those rows take the exact path); (3) the shard stages (pruned owner rescoring with the 2 E margin + finalize) agree
    with them on every row they certify."""
    import torch

    M, N, sig, k = 2000, 100_000, (60, 60), 20
    g = torch.Generator(device="cuda"); g.manual_seed(33)
    exp = torch.randint(0, 256, (M,) + sig, dtype=torch.uint8, device="cuda", generator=g)
    dic = torch.rand((N,) + sig, dtype=torch.float32, device="cuda", generator=g)
    code = _lib.KDI_NCC if metric == "ncc" else _lib.KDI_NDP
    assert ctx.candidate_capacity(k) == 32
    few = M // 50 if metric == "ncc" else M
    idx0 = torch.empty((M, k), dtype=torch.int64, device="cuda")
    sc0 = torch.empty((M, k), dtype=torch.float32, device="cuda")
    # the default: rows proven with the bound where their scores allow it, on the model elsewhere (counted);
    # with 32-entry lists most random NCC rows are proven, with 64-entry lists (KDI_OPT_CERT_WIDEN) all of
    # them; the model alone (mode 0) counts every row.  Same results in all of them
    counts = {}
    for mode, widen in ((0, 1), (2, 1), (2, 0)):
        ctx.set_option(_lib.OPT_CERT_STRICT, mode)
        ctx.set_option(_lib.OPT_CERT_WIDEN, widen)
        i_m = torch.empty_like(idx0); s_m = torch.empty_like(sc0)
        ctx.dictionary_indexing(exp, M, dic, N, code, k, out=(i_m, s_m))
        tm = ctx.timings()
        counts[(mode, widen)] = (int(tm["model_rows"]), int(tm["flagged_rows"]))
        if mode == 0:
            idx0, sc0 = i_m, s_m
        else:
            assert torch.equal(idx0, i_m) and torch.equal(sc0, s_m), (mode, widen)
    assert counts[(0, 1)][0] + counts[(0, 1)][1] == M
    if metric == "ncc":
        assert counts[(2, 1)] == (0, 0), counts           # every row proven
        assert 0 < counts[(2, 0)][0] < M // 2, counts      # 32-entry lists: most rows proven
    else:
        assert counts[(2, 1)] == counts[(2, 0)], counts   # NDP lists are not widened
    print(f"certificate counts (model rows, flagged rows) by (mode, widen), {metric}: {counts}")
    ratios = {}
    try:
        for compute in (1, 0):
            # (1) measured error of the tensor-core scores against the bound, default list size
            ctx.set_option(_lib.OPT_COMPUTE_DTYPE, compute)
            shard, approx, gidx = ctx.shard_candidates(exp, M, dic, N, code, k)
            exact = shard.rescore_owned(gidx)
            ok = gidx >= 0
            err = float((approx - exact)[ok].abs().max())
            bound = ctx.certificate_bound(sig[0] * sig[1], compute)
            ratios[compute] = err / bound
            assert err <= bound, (compute, err, bound)
            shard.close()
        assert 9e-4 < ctx.certificate_bound(3600, 0) < 2e-3 and ctx.certificate_bound(3600, 1) > 7e-3
        ctx.set_option(_lib.OPT_CERT_STRICT, 1)
        assert ctx.candidate_capacity(k) == 64 and ctx.candidate_capacity(30) == 128
        assert ctx.candidate_capacity(104) == 128 and ctx.candidate_capacity(105) == 0
        # (2) the whole pipeline
        idx1 = torch.empty((M, k), dtype=torch.int64, device="cuda")
        sc1 = torch.empty((M, k), dtype=torch.float32, device="cuda")
        ctx.dictionary_indexing(exp, M, dic, N, code, k, out=(idx1, sc1))
        tm = ctx.timings()
        assert tm["gemm_launches"] >= 1
        assert torch.equal(idx0, idx1) and torch.equal(sc0, sc1)
        assert tm["flagged_rows"] <= few, tm["flagged_rows"]
        flagged_strict = tm["flagged_rows"]
        # (3) the stages of a sharded job on one GPU
        shard, approx, gidx = ctx.shard_candidates(exp, M, dic, N, code, k)
        assert approx.shape[1] == 64
        exact_all = shard.rescore_owned(gidx)
        exact = shard.rescore_owned(gidx, approx, k)
        pruned = torch.isinf(exact) & (gidx >= 0)
        assert bool(pruned.any()) or metric == "ndp"  # the margin does prune something out of 64 candidates
        kth = torch.sort(exact_all, dim=1, descending=True).values[:, k - 1:k]
        assert bool((exact_all[pruned] < kth.expand_as(exact_all)[pruned]).all())  # nothing pruned belonged to the top k
        i2, s2, flags = shard.finalize(approx, gidx, exact, k, N)
        good = torch.ones(M, dtype=torch.bool, device="cuda")
        good[flags.long()] = False
        assert int(flags.numel()) <= few
        assert torch.equal(i2[good], idx0[good]) and torch.equal(s2[good], sc0[good])
        shard.close()
    finally:
        ctx.set_option(_lib.OPT_CERT_STRICT, 2)
        ctx.set_option(_lib.OPT_CERT_WIDEN, 0)
        ctx.set_option(_lib.OPT_COMPUTE_DTYPE, 0)
    print(f"strict certificate ({metric}): measured max error / bound = {ratios[0]:.4f} (fp16), {ratios[1]:.4f} (bf16); "
          f"{flagged_strict} of {M} rows through the exact path")


@pytest.mark.parametrize("keep_n", [20, 50])
def test_strict_certificate_near_ties_and_large_keep_n(ctx, keep_n):
    """Strict certificate on the adversarial dictionary (300 rows ~1e-6 apart in score, more than any list
    holds) and with the 128-entry lists: equal to the forced-exact results bit for bit."""
    rng = np.random.default_rng(41)
    dic = orc.synthetic_dictionary(6000, (30, 30), seed=42)
    base = dic[17].copy()
    dic[1000:1300] = base[None] * (1.0 + 1e-4 * rng.standard_normal((300, 30, 30)).astype(np.float32))
    exp = orc.synthetic_experimental(64, (30, 30), seed=43)
    exp[:32] = np.clip(np.rint(255 * (0.8 * base[None] + 0.2 * rng.random((32, 30, 30)))), 0, 255).astype(np.uint8)
    ctx.set_option(_lib.OPT_CERT_STRICT, 1)
    try:
        i1, s1 = ctx.dictionary_indexing(exp, 64, dic, 6000, _lib.KDI_NCC, keep_n)
        tm = ctx.timings()
        ctx.set_option(_lib.OPT_FORCE_EXACT, 1)
        i2, s2 = ctx.dictionary_indexing(exp, 64, dic, 6000, _lib.KDI_NCC, keep_n)
    finally:
        ctx.set_option(_lib.OPT_FORCE_EXACT, 0)
        ctx.set_option(_lib.OPT_CERT_STRICT, 2)
    assert tm["gemm_launches"] >= 1
    assert np.array_equal(i1, i2) and np.array_equal(s1, s2)
    assert tm["flagged_rows"] >= 32  # the planted rows cannot be certified


# ---- division route of the prepare kernels, view-mode dictionaries ------------------------------------

def _awkward_rows(rng, n, s):
    """float32 rows that stress the division by the row norm: ordinary, scaled far up and down, nearly
    and exactly constant, already centred, with zeros and with tiny (down to subnormal) values."""
    x = rng.random((n, s), dtype=np.float32)
    x[1] *= 1e-20
    x[2] *= 1e20
    x[3] = 0.5
    x[4] = 0.5
    x[4, 7] = np.nextafter(np.float32(0.5), np.float32(1))
    x[5] -= x[5].mean()
    x[6, ::3] = 0.0
    x[7, ::5] = 1e-39          # subnormal elements among ordinary ones
    x[8, ::2] = 1e-33
    x[9] = 0.0
    x[10] *= 1e-36
    x[11] = (x[11] - 0.5) * 1e-12
    x[12] *= 3e9               # norm beyond 2^30
    x[13] *= 1e-11             # norm below 2^-30
    x[14, 0] = np.float32(np.inf)
    return x


@pytest.mark.parametrize("src_dtype", [np.float32, np.uint8, np.uint16, np.float64])
@pytest.mark.parametrize("masked", [False, True])
def test_fma_division_route_is_bit_identical(ctx, src_dtype, masked):
    """The prepare kernels divide by the row norm with a float32 FMA sequence wherever that is exactly
    rounded and through the double reciprocal elsewhere (kdi_internal.cuh: kdi_div_fma / kdi_rowdiv):
    every kernel variant, both metrics and awkward rows must give the bits of the all-double route -
    float32 rows directly, the 16-bit operands through the tensor-core products."""
    rng = np.random.default_rng(77)
    sig = (36, 40)
    n = 300
    if np.issubdtype(src_dtype, np.floating):
        raw = _awkward_rows(rng, n, sig[0] * sig[1]).astype(src_dtype).reshape((n,) + sig)
    else:
        raw = (rng.random((n,) + sig) * (60000 if src_dtype == np.uint16 else 255)).astype(src_dtype)
        raw[3] = 7
        raw[9] = 0
    dic = orc.synthetic_dictionary(512, sig, seed=6)
    smask = orc.circular_signal_mask(sig) if masked else None
    out = {}
    try:
        ctx.set_signal_mask(smask)
        for route in (0, 1):
            ctx.set_option(_lib.OPT_DIV_DOUBLE, route)
            got = []
            for code in (_lib.KDI_NCC, _lib.KDI_NDP):
                with ctx.patterns(raw, n, code) as p, ctx.patterns(dic, 512, code) as d:
                    got.append(np.asarray(p).view(np.uint32))
                    got.append(ctx.debug_gemm16(p, d).view(np.uint32))
            out[route] = got
    finally:
        ctx.set_option(_lib.OPT_DIV_DOUBLE, 0)
        ctx.set_signal_mask(None)
    for a, b in zip(out[0], out[1]):
        assert np.array_equal(a, b)
    # and the values are the reference's (ordinary rows)
    want = orc.prepare_experimental(raw[20:], "ncc", n - 20, signal_mask=smask)
    got = out[0][0].view(np.float32)[20:]
    assert got.shape == want.shape and np.nanmax(np.abs(got - want)) < 2e-6


@pytest.mark.parametrize("metric", ["ncc", "ndp"])
@pytest.mark.parametrize("compute", ["fp16", "bf16"])
def test_view_mode_dictionary_is_bit_identical(ctx, metric, compute):
    """A device-resident float32 dictionary is kept as a VIEW by the driver (no normalised float32 copy;
    exact scores recomputed from the caller's rows with the prepare kernel's arithmetic): indices and
    scores must equal those of the copying mode bit for bit - ordinary rows, awkward rows (double
    route, NaN rows), near-ties that send rows through the exact path (which materialises the copy)."""
    import torch

    rng = np.random.default_rng(91)
    sig = (40, 40)
    M, N, k = 700, 9000, 20
    dic = rng.random((N, sig[0] * sig[1]), dtype=np.float32)
    dic[:15] = _awkward_rows(rng, 15, sig[0] * sig[1])[:15]
    # (rows whose scores are NaN - constant, all-zero, infinite - are left to the division-route test: NaN
    # scores have no defined rank, and which NaN payload survives a sum is not part of the contract)
    dic[[3, 9, 14]] = rng.random((3, sig[0] * sig[1]), dtype=np.float32)
    base = dic[4000].copy()
    dic[5000:5100] = base[None] * (1.0 + 1e-4 * rng.standard_normal((100, sig[0] * sig[1])).astype(np.float32))
    exp = orc.synthetic_experimental(M, sig, seed=92)
    exp[:16] = np.clip(np.rint(255 * (0.8 * base.reshape(sig)[None] + 0.2 * rng.random((16,) + sig))), 0, 255).astype(np.uint8)
    code = _lib.KDI_NCC if metric == "ncc" else _lib.KDI_NDP
    d_exp, d_dic = torch.from_numpy(exp).cuda(), torch.from_numpy(dic).cuda()
    out = {}
    try:
        ctx.set_option(_lib.OPT_COMPUTE_DTYPE, 1 if compute == "bf16" else 0)
        for view in (2, 0):  # (2: wherever possible - the default only where it pays)
            ctx.set_option(_lib.OPT_DICT_VIEW, view)
            idx = torch.empty((M, k), dtype=torch.int64, device="cuda")
            sc = torch.empty((M, k), dtype=torch.float32, device="cuda")
            ctx.dictionary_indexing(d_exp, M, d_dic, N, code, k, out=(idx, sc))
            out[min(view, 1)] = (idx.cpu().numpy(), sc.cpu().numpy().view(np.uint32), ctx.timings()["flagged_rows"])
    finally:
        ctx.set_option(_lib.OPT_DICT_VIEW, 1)
        ctx.set_option(_lib.OPT_COMPUTE_DTYPE, 0)
    bad = np.flatnonzero((out[1][0] != out[0][0]).any(axis=1) | (out[1][1] != out[0][1]).any(axis=1))
    assert bad.size == 0, (bad[:8], out[1][0][bad[:2]], out[0][0][bad[:2]], out[1][1][bad[:2]], out[0][1][bad[:2]])
    assert out[1][2] == out[0][2] and out[1][2] >= 16
    # the host path (always copying; what the oracle-parity tests above exercise) agrees too
    i_h, s_h = ctx.dictionary_indexing(exp, M, dic, N, code, k)
    assert np.array_equal(i_h, out[1][0]) and np.array_equal(s_h.view(np.uint32), out[1][1])
    # and so does the oracle on a sample, with the awkward dictionary rows left out (NumPy's float32
    # sums overflow / underflow on them)
    rows = np.arange(16, M, 57)
    d_ok = torch.from_numpy(dic[15:]).cuda()
    idx = torch.empty((rows.size, k), dtype=torch.int64, device="cuda")
    sc = torch.empty((rows.size, k), dtype=torch.float32, device="cuda")
    ctx.dictionary_indexing(torch.from_numpy(exp[rows]).cuda(), rows.size, d_ok, N - 15, code, k, out=(idx, sc))
    ridx, rsc = orc.dictionary_indexing(exp[rows], dic[15:].reshape((N - 15,) + sig), metric=metric, keep_n=k)
    _check(ridx, rsc, idx.cpu().numpy(), sc.cpu().numpy())


def test_view_mode_is_used_and_saves_the_float32_copy(ctx):
    """The view really is what runs for a device-resident float32 dictionary (fewer bytes written by the
    prepare step: its time drops), and it is not used for host or masked dictionaries."""
    import torch

    M, N, sig, k = 2048, 40_000, (60, 60), 20
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    exp = torch.randint(0, 256, (M,) + sig, dtype=torch.uint8, device="cuda", generator=g)
    dic = torch.rand((N,) + sig, dtype=torch.float32, device="cuda", generator=g)
    idx = torch.empty((M, k), dtype=torch.int64, device="cuda")
    sc = torch.empty((M, k), dtype=torch.float32, device="cuda")
    res = {}
    try:
        ctx.set_option(_lib.OPT_EARLY_SPLIT, 0)  # (the whole dictionary in one prepare launch, both ways)
        for view in (1, 0):
            ctx.set_option(_lib.OPT_DICT_VIEW, view)
            best = 1e9
            for _ in range(4):
                ctx.dictionary_indexing(exp, M, dic, N, _lib.KDI_NCC, k, out=(idx, sc))
                best = min(best, ctx.timings()["normalize_dict_ms"])
            res[view] = (best, idx.cpu().numpy().copy(), sc.cpu().numpy().copy())
    finally:
        ctx.set_option(_lib.OPT_DICT_VIEW, 1)
        ctx.set_option(_lib.OPT_EARLY_SPLIT, 1)
    assert np.array_equal(res[1][1], res[0][1]) and np.array_equal(res[1][2], res[0][2])
    assert res[1][0] < res[0][0], (res[1][0], res[0][0])


# ---- dtype=float64 (the reference's metrics accept it; tests/test_indexing/test_dictionary_indexing.py:45-59) ----

def test_float64_reference_test_case(dummy_array):
    """``test_dictionary_indexing_signal_mask``: 64-bit floats, a signal mask, n_per_iteration=2, rechunk."""
    dic = dummy_array.reshape(-1, 3, 3)
    smask = np.array([[0, 0, 0], [0, 1, 0], [0, 0, 0]], dtype=bool)
    res = kb.dictionary_indexing(dummy_array, dic, dtype=np.float64, n_per_iteration=2, signal_mask=smask,
                                 rechunk=True, verbose=False)
    assert res.scores.dtype == np.float64 and res.simulation_indices.dtype == np.int64
    assert np.allclose(res.scores[:, 0], 1, rtol=0, atol=1e-14)
    assert np.array_equal(res.simulation_indices[:, 0], np.arange(9))
    ridx, rsc = orc.dictionary_indexing(dummy_array, dic, signal_mask=smask, dtype=np.float64)
    assert np.max(np.abs(res.scores - rsc)) < 1e-13


@pytest.mark.parametrize("metric", ["ncc", "ndp"])
@pytest.mark.parametrize("masked, nav", [(False, False), (True, True)])
def test_float64_mode_matches_the_float64_oracle(ctx, metric, masked, nav):
    """dtype=float64 on the GPU metrics: float64 scores within 1e-12 of the reference arithmetic in
    float64, indices identical (the ranking is decided by the float64 scores), for host arrays, CUDA
    tensors, chunked dictionaries, masks, and near-duplicate dictionary rows that force the candidate
    list to grow."""
    import torch

    rng = np.random.default_rng(5)
    sig = (30, 30)
    M, N, k = 96, 5000, 12
    exp = orc.synthetic_experimental(M, sig, seed=61)
    dic = orc.synthetic_dictionary(N, sig, seed=62)
    # 40 rows that differ by ~1e-7 relative: float32 scores tie or swap, float64 scores do not
    dic[300:340] = dic[7][None] * (1.0 + 1e-7 * rng.standard_normal((40,) + sig)).astype(np.float32)
    exp[:8] = np.clip(np.rint(255 * (0.8 * dic[7][None] + 0.2 * rng.random((8,) + sig))), 0, 255).astype(np.uint8)
    smask = orc.circular_signal_mask(sig) if masked else None
    nmask = (rng.random(M) < 0.3).reshape(8, 12) if nav else None
    exp4 = exp.reshape((8, 12) + sig)
    ridx, rsc = orc.dictionary_indexing(exp4 if nav else exp, dic, metric=metric, keep_n=k, signal_mask=smask,
                                        navigation_mask=nmask, dtype=np.float64,
                                        n_experimental_patterns=M)
    for source in ("host", "cuda", "chunks"):
        e = torch.from_numpy(exp4).cuda() if source == "cuda" else exp4
        d = torch.from_numpy(dic).cuda() if source == "cuda" else dic
        res = kb.dictionary_indexing(e, d, metric=metric, keep_n=k, dtype=np.float64, signal_mask=smask,
                                     navigation_mask=nmask, n_per_iteration=1700 if source == "chunks" else None,
                                     context=ctx, verbose=False)
        sc, idx = res.scores, res.simulation_indices  # (the indexed points only)
        assert sc.dtype == np.float64 and sc.shape == rsc.shape
        assert np.max(np.abs(sc - rsc)) < 1e-12, source
        # identical order wherever the reference's float64 scores are separated at all (1e-13)
        r = orc.compare_topk(ridx, rsc, idx, sc, tie_tol=1e-13, score_tol=1e-12)
        assert r["scores_ok"] and r["tie_ok"], (source, r)
    ctx.set_signal_mask(None)


def test_float64_metric_call_and_plugin_hooks(golden):
    """``metric(exp, dict)`` in float64 (the full block) and the hooks the reference driver calls."""
    g = golden("config1_nickel_x_1000.npz")
    dic = orc.synthetic_dictionary(1000, (60, 60), seed=2)
    for name, cls in (("ncc", kb.NormalizedCrossCorrelationMetric), ("ndp", kb.NormalizedDotProductMetric)):
        m = cls(9, 1000, dtype=np.float64)
        block = m(g["nickel"], dic)
        assert block.dtype == np.float64 and block.shape == (9, 1000)
        assert np.max(np.abs(block - g[f"{name}_sim_f64"])) < 1e-13
        sim = m.match(m.prepare_experimental(g["nickel"]), m.prepare_dictionary(dic.reshape(1000, -1)))
        idx, sc = sim.argtopk(5, axis=-1), sim.topk(5, axis=-1)
        want = np.argsort(-g[f"{name}_sim_f64"], axis=1, kind="stable")[:, :5]
        assert np.array_equal(idx, want) and np.allclose(sc, np.take_along_axis(g[f"{name}_sim_f64"], want, 1), atol=1e-13, rtol=0)


def test_float64_scores_of_rows_beyond_the_shared_memory_staging(ctx):
    """``kdi_scores_f64`` stages rows of more than 25 600 values in global memory."""
    import torch

    rng = np.random.default_rng(8)
    exp = rng.integers(0, 256, (5, 170, 170), dtype=np.uint8)
    dic = rng.random((40, 170, 170), dtype=np.float32)
    cand = np.stack([rng.permutation(40)[:6] for _ in range(5)]).astype(np.int64)
    ctx.set_signal_mask(None)
    got = ctx.scores_f64(torch.from_numpy(exp.reshape(5, -1)).cuda(), None, torch.from_numpy(dic.reshape(40, -1)).cuda(),
                         _lib.KDI_NCC, cand)
    e = orc.prepare_experimental(exp, "ncc", 5, dtype=np.float64)
    d = orc.prepare_dictionary(dic.reshape(40, -1), "ncc", None, np.float64)
    want = np.take_along_axis(e @ d.T, cand, axis=1)
    assert np.max(np.abs(got - want)) < 1e-13


@pytest.mark.parametrize("M, N, sig, metric", [(2048, 6000, (60, 60), "ncc"), (2100, 30001, (60, 60), "ncc"),
                                               (5000, 9000, (40, 40), "ndp"), (2600, 4000, (57, 43), "ncc")])
def test_dual_row_block_tile_equals_default(ctx, M, N, sig, metric):
    """The 512 x 256 pair tile of the tensor-core kernel (KDI_OPT_GEMM_DUAL = 2: two row blocks per CTA, both
    halves of TMEM as accumulators) against the 256 x 256 tile: identical indices and scores - ragged row and
    dictionary counts, row counts that leave the second row block of the last CTA pair empty, K that is not a
    multiple of 64, both operand types, device-resident and host dictionaries."""
    import torch

    code = _lib.KDI_NCC if metric == "ncc" else _lib.KDI_NDP
    exp = orc.synthetic_experimental(M, sig, seed=21)
    dic = orc.synthetic_dictionary(N, sig, seed=22)
    exp_d, dic_d = torch.from_numpy(exp).cuda(), torch.from_numpy(dic).cuda()
    out = {}
    try:
        for dual in (0, 2):
            ctx.set_option(_lib.OPT_GEMM_DUAL, dual)
            res = []
            for dt in (0, 1):
                ctx.set_option(_lib.OPT_COMPUTE_DTYPE, dt)
                idx = torch.empty((M, 20), dtype=torch.int64, device="cuda")
                sc = torch.empty((M, 20), dtype=torch.float32, device="cuda")
                ctx.dictionary_indexing(exp_d, M, dic_d, N, code, 20, out=(idx, sc))
                res += [idx.cpu().numpy(), sc.cpu().numpy()]
            ctx.set_option(_lib.OPT_COMPUTE_DTYPE, 0)
            res += list(ctx.dictionary_indexing(exp, M, dic, N, code, 20))
            out[dual] = res
            assert ctx.timings()["gemm_launches"] >= 1
    finally:
        ctx.set_option(_lib.OPT_GEMM_DUAL, 1)
        ctx.set_option(_lib.OPT_COMPUTE_DTYPE, 0)
    for a, b in zip(out[0], out[2]):
        assert np.array_equal(a, b)
    ridx, rsc = orc.dictionary_indexing(exp[:64], dic, metric=metric, keep_n=20)
    _check(ridx, rsc, out[2][0][:64], out[2][1][:64])
