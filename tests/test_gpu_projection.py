"""Dictionary generation on the device (include/kdi.h: kdi_project_patterns,
kdi_patterns_create_projected, kdi_dictionary_indexing_projected) against the golden vectors
produced by the reference's own Numba kernels and against the NumPy restatement.

Tolerance: the reference computes the projection in float64 (compiled with fastmath) and casts to
float32; the kernel does the same arithmetic in float64 with CUDA's libm, so after the cast the
patterns agree to one float32 ulp (nearly all values bit-identical)."""

import numpy as np
import pytest

import kikuchipy_b200 as kb
from kikuchipy_b200 import _lib
from oracle import di_oracle as orc
from oracle import projection_oracle as po

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = kb.default_context(0)
    yield c
    c.set_signal_mask(None)
    c.set_option(_lib.OPT_OVERLAP, 1)


def _close(a, b, min_identical=0.99):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    assert a.shape == b.shape
    ulp = np.spacing(np.maximum(np.abs(a), np.abs(b)).astype(np.float32))
    assert np.all(np.abs(a - b) <= 1.01 * ulp), float(np.abs(a - b).max())
    assert np.mean(a == b) >= min_identical, float(np.mean(a == b))


def test_projection_matches_reference_golden(ctx, golden):
    z = golden("projection.npz")
    n = z["mu32"].shape[0]
    mp = ctx.master_pattern(z["mu32"], z["ml32"], z["dc_all"], scale=float(z["scale"]))
    _close(ctx.project_patterns(mp, z["rotations"]), z["out_f32"])
    mp8 = ctx.master_pattern(z["mu8"], z["ml8"], z["dc_all"], scale=(n - 1) / 2, rescale=True, out_min=-1.0, out_max=1.0)
    _close(ctx.project_patterns(mp8, z["rotations"]), z["out_u8"])
    mp64 = ctx.master_pattern(z["mu32"].astype(np.float64), z["ml32"].astype(np.float64), z["dc_all"])
    _close(ctx.project_patterns(mp64, z["rotations"]), z["out_f32"])
    for m in (mp, mp8, mp64):
        m.close()


def test_projection_matches_oracle_detector_sized(ctx):
    import torch

    mu, ml = po.synthetic_master_pattern(401, seed=5)
    dc = po.direction_cosines_fixed_pc([-0.9, 0.85, -0.7, 0.95], 0.5, 60, 60, po.tilted_detector_matrix(70.0))
    rot = po.random_rotations(150, seed=4)
    ref = po.project_patterns(rot, dc, mu, ml)
    mp = ctx.master_pattern(mu, ml, dc)
    _close(ctx.project_patterns(mp, rot), ref)
    # device rotations, device output
    out = torch.empty((150, 3600), dtype=torch.float32, device="cuda")
    ctx.project_patterns(mp, torch.from_numpy(rot).cuda(), out=out)
    _close(out.cpu().numpy(), ref)
    # rescaled variant
    mpr = ctx.master_pattern(mu, ml, dc, rescale=True, out_min=-1.0, out_max=1.0)
    refr = po.project_patterns(rot, dc, mu, ml, rescale=True, out_min=-1.0, out_max=1.0)
    got = ctx.project_patterns(mpr, rot)
    _close(got, refr, min_identical=0.97)
    assert got.min() == -1.0 and got.max() == 1.0
    mp.close(); mpr.close()


@pytest.mark.parametrize("metric", ["ncc", "ndp"])
def test_fused_projection_and_prepare_equal_two_steps(ctx, metric):
    """get_patterns -> prepare_dictionary in one kernel == the two steps one after the other."""
    code = _lib.KDI_NCC if metric == "ncc" else _lib.KDI_NDP
    mu, ml = po.synthetic_master_pattern(301, seed=7)
    dc = po.direction_cosines_fixed_pc([-0.9, 0.85, -0.7, 0.95], 0.5, 40, 40, po.tilted_detector_matrix(70.0))
    rot = po.random_rotations(300, seed=9)
    mp = ctx.master_pattern(mu, ml, dc)
    for smask in (None, orc.circular_signal_mask((40, 40))):
        ctx.set_signal_mask(smask)
        fused = np.asarray(ctx.patterns_projected(mp, rot, code))
        two = np.asarray(ctx.patterns(ctx.project_patterns(mp, rot), 300, code))
        assert np.array_equal(fused, two)
        ref = orc.prepare_dictionary(po.project_patterns(rot, dc, mu, ml), metric, smask)
        assert np.abs(fused - ref).max() < 2e-6
    ctx.set_signal_mask(None)
    mp.close()


def test_dictionary_indexing_with_generated_dictionary(ctx):
    """The public API with a GeneratedDictionary (lazy get_patterns result) == the same call on
    the materialised patterns == the oracle on the oracle's projection."""
    mu, ml = po.synthetic_master_pattern(301, seed=7)
    sig = (40, 40)
    dc = po.direction_cosines_fixed_pc([-0.9, 0.85, -0.7, 0.95], 0.5, sig[0], sig[1], po.tilted_detector_matrix(70.0))
    rot = po.random_rotations(5000, seed=9)
    gen = kb.get_patterns(mu, ml, rot, direction_cosines=dc, detector_shape=sig, context=ctx)
    assert isinstance(gen, kb.GeneratedDictionary) and gen.shape == (5000, 40, 40) and gen.dtype == np.float32
    patterns = gen.compute()
    assert patterns.shape == (5000, 40, 40) and patterns.dtype == np.float32
    # experimental patterns: noisy 8-bit versions of some dictionary entries
    rng = np.random.default_rng(1)
    j = rng.integers(0, 5000, 12 * 9)
    p = patterns[j]
    p = (p - p.min()) / (p.max() - p.min())
    exp = np.clip(np.rint(255 * (0.8 * p + 0.2 * rng.random(p.shape))), 0, 255).astype(np.uint8).reshape(12, 9, 40, 40)
    nav = rng.random((12, 9)) < 0.15
    smask = orc.circular_signal_mask(sig)
    for kwargs in ({}, {"navigation_mask": nav, "signal_mask": smask}):
        r_gen = kb.dictionary_indexing(exp, gen, keep_n=10, verbose=False, context=ctx, **kwargs)
        r_mat = kb.dictionary_indexing(exp, patterns, keep_n=10, verbose=False, context=ctx, **kwargs)
        assert np.array_equal(r_gen.simulation_indices, r_mat.simulation_indices)
        assert np.array_equal(r_gen.scores, r_mat.scores)
        assert r_gen.rotations is not None and r_gen.rotations.shape[-1] == 4
        ref_patterns = po.project_patterns(rot, dc, mu, ml).reshape(5000, 40, 40)
        ridx, rsc = orc.dictionary_indexing(exp, ref_patterns, keep_n=10, **kwargs)
        c = orc.compare_topk(ridx, rsc, r_gen.simulation_indices, r_gen.scores, tie_tol=2e-5)
        assert c["tie_ok"] and c["scores_ok"], c
    keep = ~nav.ravel()
    assert np.array_equal(r_gen.simulation_indices[:, 0], j[keep])
    gen.master_pattern.close()


def test_generated_dictionary_overlapped_schedule_and_errors(ctx):
    import torch

    ctx.set_signal_mask(None)
    mu, ml = po.synthetic_master_pattern(201, seed=3)
    dc = po.direction_cosines_fixed_pc([-0.9, 0.85, -0.7, 0.95], 0.5, 30, 30, po.tilted_detector_matrix(70.0))
    rot = po.random_rotations(7000, seed=2)
    mp = ctx.master_pattern(mu, ml, dc)
    exp = torch.from_numpy(orc.synthetic_experimental(1500, (30, 30), seed=5)).cuda()
    res = {}
    try:
        for mode in (0, 2):
            ctx.set_option(_lib.OPT_OVERLAP, mode)
            ctx.set_option(_lib.OPT_SUPERBLOCK, 2)
            idx = torch.empty((1500, 20), dtype=torch.int64, device="cuda")
            sc = torch.empty((1500, 20), dtype=torch.float32, device="cuda")
            ctx.dictionary_indexing_projected(exp, 1500, mp, rot, _lib.KDI_NCC, 20, out=(idx, sc))
            res[mode] = (idx.cpu().numpy(), sc.cpu().numpy())
    finally:
        ctx.set_option(_lib.OPT_OVERLAP, 1)
        ctx.set_option(_lib.OPT_SUPERBLOCK, 0)
    assert np.array_equal(res[0][0], res[2][0]) and np.array_equal(res[0][1], res[2][1])
    with pytest.raises(ValueError, match="signal sizes must be identical"):
        ctx.dictionary_indexing_projected(np.zeros((4, 31, 31), np.uint8), 4, mp, rot, _lib.KDI_NCC, 5)
    with pytest.raises(NotImplementedError):
        kb.get_patterns(mu, ml, rot, direction_cosines=dc, dtype_out="uint8")
    with pytest.raises(ValueError, match="either a detector or direction cosines"):
        kb.get_patterns(mu, ml, rot)
    mp.close()


def test_projection_with_one_pc_per_rotation(ctx, golden):
    """_project_patterns_from_master_pattern_with_varying_pc: direction cosines computed on the
    device from each rotation's own projection centre, against the reference's golden outputs."""
    z = golden("projection_varying_pc.npz")
    nrows, ncols = int(z["nrows"]), int(z["ncols"])
    mu32, ml32 = po.synthetic_master_pattern(101, seed=5, dtype=np.float32)
    mu8, ml8 = po.synthetic_master_pattern(101, seed=6, dtype=np.uint8)
    got = kb.get_patterns(mu32, ml32, z["rotations"], pcs=z["pcs"], om_detector_to_sample=z["om"], detector_shape=(nrows, ncols))
    assert got.shape == (7, nrows, ncols)
    _close(got.reshape(7, -1), z["out_f32"])
    got8 = kb.get_patterns(mu8, ml8, z["rotations"], pcs=z["pcs"], om_detector_to_sample=z["om"], detector_shape=(nrows, ncols))
    _close(got8.reshape(7, -1), z["out_u8"])
    # a detector object with one PC per rotation takes the same path; a mismatch is refused
    det = kb.Detector((nrows, ncols), pc=z["pcs"])
    det_om = z["om"]

    class Det:
        shape, pc, om_detector_to_sample = det.shape, det.pc, det_om

    assert np.array_equal(kb.get_patterns(mu32, ml32, z["rotations"], Det), got)
    with pytest.raises(ValueError, match="navigation_shape"):
        kb.get_patterns(mu32, ml32, z["rotations"][:3], Det)
    # many rotations: batches, every pattern equal to its single-PC projection
    rot = po.random_rotations(300, seed=3)
    pcs = np.array([0.5, 0.3, 0.6]) + np.random.default_rng(1).normal(scale=0.03, size=(300, 3))
    many = kb.get_patterns(mu32, ml32, rot, pcs=pcs, om_detector_to_sample=z["om"], detector_shape=(nrows, ncols))
    _close(many.reshape(300, -1), po.project_patterns_varying_pc(rot, pcs, nrows, ncols, z["om"], mu32, ml32))


def test_own_arithmetic_equals_math_library_version(ctx):
    """The projection kernel's own Newton / polynomial sequences (default) against the CUDA math
    library version of the same pixel (KDI_OPT_PROJECT_LIBM): float32 patterns within one ulp and
    almost everywhere identical - for random rotations, for the symmetric rotations whose terms
    cancel exactly (poles, hemisphere boundary, |x| == |y| diagonals), for a float64 master pattern
    (32-byte tap elements), for quaternions that are not unit (the normalisation's long way) and
    with the per-pattern rescale."""
    mu, ml = po.synthetic_master_pattern(301, seed=7)
    dc = po.direction_cosines_fixed_pc([-0.9, 0.85, -0.7, 0.95], 0.5, 48, 52, po.tilted_detector_matrix(70.0))
    # direction cosines that hit the poles and the diagonals exactly
    special = np.array([[0, 0, 1], [0, 0, -1], [1, 0, 0], [0, 1, 0], [-1, 0, 0], [0, -1, 0],
                        [np.sqrt(0.5), np.sqrt(0.5), 0], [np.sqrt(0.5), -np.sqrt(0.5), 0],
                        [1 / np.sqrt(3)] * 3, [0.6, 0.0, 0.8], [0.0, 0.6, -0.8]], dtype=np.float64)
    dc = np.concatenate([dc, special])
    s = np.sqrt(0.5)
    sym = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1], [s, s, 0, 0], [s, 0, s, 0], [s, 0, 0, s],
                    [0.5, 0.5, 0.5, 0.5], [0.5, -0.5, 0.5, -0.5], [0, s, s, 0]], dtype=np.float64)
    rot = np.concatenate([sym, po.random_rotations(400, seed=11)])
    scaled = rot * np.linspace(0.7, 1.4, rot.shape[0])[:, None]  # not unit: v / |v| has to do the work
    cases = [
        (ctx.master_pattern(mu, ml, dc), rot, 0.995),
        (ctx.master_pattern(mu.astype(np.float64), ml.astype(np.float64), dc), rot, 0.995),
        (ctx.master_pattern(mu, ml, dc), scaled, 0.995),
        (ctx.master_pattern(mu, ml, dc, rescale=True, out_min=-1.0, out_max=1.0), rot, 0.97),
    ]
    try:
        for mp, r, min_identical in cases:
            ctx.set_option(_lib.OPT_PROJECT_LIBM, 1)
            lib = ctx.project_patterns(mp, r)
            ctx.set_option(_lib.OPT_PROJECT_LIBM, 0)
            own = ctx.project_patterns(mp, r)
            assert np.isfinite(own).all()
            _close(own, lib, min_identical=min_identical)
            mp.close()
        # and against the oracle (NumPy float64) on the symmetric rotations
        mp = ctx.master_pattern(mu, ml, dc)
        _close(ctx.project_patterns(mp, sym), po.project_patterns(sym, dc, mu, ml), min_identical=0.99)
        mp.close()
    finally:
        ctx.set_option(_lib.OPT_PROJECT_LIBM, 0)


def test_generated_dictionary_schedules_agree(ctx):
    """A generated dictionary projected in one piece on the second stream (default), split into a
    quarter + three quarters (KDI_OPT_EARLY_SPLIT = 2) and prepared before anything else
    (KDI_OPT_EARLY_SPLIT = 0): identical results."""
    import torch

    ctx.set_signal_mask(None)
    mu, ml = po.synthetic_master_pattern(201, seed=3)
    dc = po.direction_cosines_fixed_pc([-0.9, 0.85, -0.7, 0.95], 0.5, 30, 30, po.tilted_detector_matrix(70.0))
    rot = po.random_rotations(20000, seed=2)
    mp = ctx.master_pattern(mu, ml, dc)
    exp = orc.synthetic_experimental(2500, (30, 30), seed=5)  # host rows: uploaded beside the projection
    res = {}
    try:
        for mode in (1, 2, 0):
            ctx.set_option(_lib.OPT_EARLY_SPLIT, mode)
            res[mode] = ctx.dictionary_indexing_projected(exp, 2500, mp, rot, _lib.KDI_NCC, 20)
            res[(mode, "dev")] = ctx.dictionary_indexing_projected(torch.from_numpy(exp).cuda(), 2500, mp, rot, _lib.KDI_NCC, 20)
    finally:
        ctx.set_option(_lib.OPT_EARLY_SPLIT, 1)
    for key in res:
        assert np.array_equal(res[key][0], res[1][0]) and np.array_equal(res[key][1], res[1][1]), key
    mp.close()
