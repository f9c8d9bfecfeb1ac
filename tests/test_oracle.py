"""The oracle (NumPy restatement) against the reference's golden vectors.

CPU only.  Golden vectors were produced by the reference's own modules
(tests/golden/make_golden.py); when /root/reference is mounted the live modules
are exercised as well.
"""

import numpy as np
import pytest

from oracle import di_oracle as orc
from oracle import ref_loader


def _sim(exp, dic, metric, nexp, signal_mask=None, navigation_mask=None, dtype=np.float32):
    e = orc.prepare_experimental(exp, metric, nexp, navigation_mask, signal_mask, dtype)
    d = orc.prepare_dictionary(dic.reshape(dic.shape[0], -1), metric, signal_mask, dtype)
    return orc.match(e, d, dtype)


@pytest.mark.parametrize("metric", ["ncc", "ndp"])
def test_config1_similarity_matches_reference(golden, metric):
    g = golden("config1_nickel_x_1000.npz")
    dic = orc.synthetic_dictionary(1000, (60, 60), seed=2)
    sim = _sim(g["nickel"], dic, metric, 9)
    # same NumPy/BLAS calls as the reference -> bit-identical here
    assert np.array_equal(sim, g[f"{metric}_sim_f32"])
    sim64 = _sim(g["nickel"], dic, metric, 9, dtype=np.float64)
    assert np.array_equal(sim64, g[f"{metric}_sim_f64"])
    simm = _sim(g["nickel"], dic, metric, 9, signal_mask=g["signal_mask"])
    assert np.array_equal(simm, g[f"{metric}_sim_masked_f32"])
    self_ = _sim(g["nickel"], g["nickel"].reshape(9, 60, 60), metric, 9)
    assert np.array_equal(self_, g[f"{metric}_self_f32"])


def test_config1_known_answers(golden):
    """Values quoted in SURVEY.md 8c for the nickel patterns."""
    g = golden("config1_nickel_x_1000.npz")
    s = g["ncc_self_f32"]
    assert np.allclose(s[0, 1:4], [0.993888, 0.991777, 0.997470], atol=1e-6)
    assert np.isclose(g["ndp_self_f32"][0, 1], 0.999525, atol=1e-6)
    assert np.max(np.abs(g["ncc_sim_f32"] - g["ncc_sim_f64"])) < 1e-6
    assert g["nickel"].min() == 23 and g["nickel"].max() == 246


def test_navigation_mask_rows(golden):
    g = golden("config1_nickel_x_1000.npz")
    dic = orc.synthetic_dictionary(1000, (60, 60), seed=2)
    sim = _sim(g["nickel"], dic, "ncc", 9, navigation_mask=g["nav_mask"])
    assert sim.shape == (7, 1000)
    assert np.array_equal(sim, g["ncc_sim_navmask_f32"])


def test_prepared_rows(golden):
    g = golden("config1_nickel_x_1000.npz")
    e = orc.prepare_experimental(g["nickel"], "ncc", 9)
    assert np.array_equal(e, g["ncc_prepared_exp"])
    assert np.allclose(e.mean(axis=1), 0, atol=1e-7)
    assert np.allclose(np.linalg.norm(e, axis=1), 1, atol=1e-6)
    e = orc.prepare_experimental(g["nickel"], "ndp", 9)
    assert np.array_equal(e, g["ndp_prepared_exp"])
    assert not np.allclose(e.mean(axis=1), 0, atol=1e-3)  # NDP does not centre


@pytest.mark.parametrize("metric", ["ncc", "ndp"])
def test_dummy_signal(golden, metric):
    g = golden("dummy_signal.npz")
    ds = g["dummy"]
    dsd = ds.reshape(-1, 3, 3)
    assert np.array_equal(_sim(ds, dsd, metric, 9), g[f"{metric}_sim"])
    assert np.array_equal(
        _sim(ds, dsd, metric, 9, signal_mask=g["signal_mask"]), g[f"{metric}_sim_masked"]
    )
    assert np.array_equal(
        _sim(ds, dsd, metric, 9, signal_mask=g["signal_mask"], dtype=np.float64),
        g[f"{metric}_sim_masked_f64"],
    )


def test_driver_self_dictionary_identity(dummy_array):
    """reference tests/test_indexing/test_dictionary_indexing.py:27-88."""
    dic = dummy_array.reshape(-1, 3, 3)
    before = dummy_array.copy()
    idx, sc = orc.dictionary_indexing(dummy_array, dic, metric="ndp")
    assert np.allclose(sc[:, 0], 1)
    assert np.array_equal(dummy_array, before)
    smask = np.array([[0, 0, 0], [0, 1, 0], [0, 0, 0]], dtype=bool)
    idx, sc = orc.dictionary_indexing(
        dummy_array, dic, dtype=np.float64, n_per_iteration=2, signal_mask=smask
    )
    assert sc.dtype == np.float64 and idx.dtype == np.int64
    assert np.allclose(sc[:, 0], 1)
    assert idx.shape == (9, 9)


def test_driver_chunked_equals_single_shot():
    exp = orc.synthetic_experimental(32, (12, 12), seed=1)
    dic = orc.synthetic_dictionary(500, (12, 12), seed=2)
    i1, s1 = orc.dictionary_indexing(exp, dic, keep_n=7)
    i2, s2 = orc.dictionary_indexing(exp, dic, keep_n=7, n_per_iteration=64)
    r = orc.compare_topk(i1, s1, i2, s2)
    assert r["tie_ok"] and r["max_dscore"] < 1e-6


def test_driver_golden(golden):
    g = golden("driver_64x4096.npz")
    exp = orc.synthetic_experimental(64, (60, 60), seed=1)
    dic = orc.synthetic_dictionary(4096, (60, 60), seed=2)
    idx, sc = orc.dictionary_indexing(exp, dic, keep_n=20)
    assert np.array_equal(idx, g["idx"]) and np.array_equal(sc, g["scores"])
    pexp, j = orc.planted_experimental(dic, 64, seed=3)
    assert np.array_equal(j, g["planted_j"])
    idx, sc = orc.dictionary_indexing(pexp, dic, keep_n=20)
    assert np.array_equal(idx[:, 0], j)
    assert np.array_equal(idx, g["planted_idx"])


def test_navigation_mask_assembly(dummy_array):
    """reference tests/test_indexing/test_dictionary_indexing.py:166-180."""
    dic = dummy_array.reshape(-1, 3, 3)
    nav = np.array([[0, 0, 0], [0, 1, 0], [0, 0, 0]], dtype=bool)
    idx, sc = orc.dictionary_indexing(dummy_array, dic, keep_n=1, navigation_mask=nav)
    assert idx.shape == (8, 1)
    s_all, i_all, in_data = orc.assemble_result(idx, sc, (3, 3), nav, 1)
    assert s_all.shape == (9,) and in_data.sum() == 8
    idx, sc = orc.dictionary_indexing(dummy_array, dic, metric="ndp", navigation_mask=~nav)
    assert idx.shape == (1, 9)


def test_osm_goldens(golden):
    g = golden("osm.npz")
    assert np.array_equal(orc.orientation_similarity_map(g["idx34"], (3, 4)), g["osm34"])
    assert np.array_equal(
        orc.orientation_similarity_map(g["idx34"], (3, 4), n_best=2, normalize=True),
        g["osm34_norm_n2"],
    )
    o = orc.orientation_similarity_map(g["idx34"], (3, 4), from_n_best=1)
    assert o.shape == (3, 4, 3) and np.array_equal(o, g["osm34_from1"])
    idx = g["idx_17x23"]
    assert np.array_equal(orc.orientation_similarity_map(idx, (17, 23)), g["osm_17x23"])
    assert np.array_equal(
        orc.orientation_similarity_map(idx, (17, 23), n_best=7, normalize=True),
        g["osm_17x23_n7_norm"],
    )
    assert np.array_equal(
        orc.orientation_similarity_map(idx, (17, 23), footprint=g["footprint8"], center_index=4),
        g["osm_17x23_fp8"],
    )
    # reference tests/test_indexing/test_orientation_similarity_map.py:27-51
    assert np.allclose(orc.orientation_similarity_map(g["idx_tiled"], (10, 10)), 5)
    assert np.allclose(
        orc.orientation_similarity_map(g["idx_tiled"], (10, 10), normalize=True), 1
    )
    with pytest.raises(ValueError, match="n_best 6 cannot be greater than"):
        orc.orientation_similarity_map(g["idx_tiled"], (10, 10), n_best=6)


def test_circular_mask_counts():
    # 3600 -> 2819, 14400 -> 11287, 6400 -> 5023 kept pixels
    for shape, kept in (((60, 60), 2819), ((120, 120), 11287), ((80, 80), 5023)):
        assert int((~orc.circular_signal_mask(shape)).sum()) == kept


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not mounted")
def test_live_reference_modules_agree():
    """Oracle == the reference's modules executed in place, on fresh random inputs."""
    _, NCC, NDP, _ = ref_loader.load_metrics()
    rng = np.random.default_rng(11)
    exp = rng.integers(0, 256, (4, 5, 16, 16), dtype=np.uint8)
    dic = rng.random((300, 16, 16), dtype=np.float32)
    smask = orc.circular_signal_mask((16, 16))
    nav = rng.random((4, 5)) < 0.3
    for name, cls in (("ncc", NCC), ("ndp", NDP)):
        for dt in (np.float32, np.float64):
            m = cls(20, 300, navigation_mask=nav, signal_mask=smask, dtype=dt)
            ref = np.asarray(m(exp, dic))
            mine = _sim(exp, dic, name, 20, smask, nav, dt)
            assert np.array_equal(ref, mine)
    ref_osm = ref_loader.load_osm()
    idx = rng.integers(0, 30, (6 * 7, 9))
    assert np.array_equal(
        ref_osm(ref_loader.FakeXmap(idx, (6, 7)), from_n_best=4, normalize=True),
        orc.orientation_similarity_map(idx, (6, 7), from_n_best=4, normalize=True),
    )


# ---- dictionary generation (master-pattern projection) ------------------------------------------

def test_projection_oracle_matches_reference_golden(golden):
    """oracle/projection_oracle.py against the outputs of the reference's own Numba kernels
    (tests/golden/make_golden_projection.py)."""
    from oracle import projection_oracle as po

    z = golden("projection.npz")
    nrows, ncols = int(z["nrows"]), int(z["ncols"])
    d = po.direction_cosines_fixed_pc(z["gnomonic_bounds"], float(z["pcz"]), nrows, ncols, z["om"])
    assert np.abs(d - z["dc"]).max() < 1e-14
    dm = po.direction_cosines_fixed_pc(z["gnomonic_bounds"], float(z["pcz"]), nrows, ncols, z["om"], z["dc_mask"])
    assert dm.shape == z["dc_masked"].shape and np.abs(dm - z["dc_masked"]).max() < 1e-14
    o1 = po.project_patterns(z["rotations"], z["dc_all"], z["mu32"], z["ml32"])
    o2 = po.project_patterns(z["rotations"], z["dc_all"], z["mu8"], z["ml8"], rescale=True, out_min=-1.0, out_max=1.0)
    for got, ref in ((o1, z["out_f32"]), (o2, z["out_u8"])):
        assert got.dtype == np.float32 and got.shape == ref.shape
        ulp = np.spacing(np.maximum(np.abs(got), np.abs(ref)))
        assert np.all(np.abs(got - ref) <= ulp) and np.mean(got == ref) > 0.999
    # the rescaled patterns span exactly [-1, 1]; both hemispheres are used
    assert o2.min() == -1.0 and o2.max() == 1.0
    assert 0.1 < np.mean([(po.rotate_vector(r, z["dc_all"])[:, 2] < 0).mean() for r in z["rotations"]]) < 0.9


def test_host_direction_cosines_match_reference_golden(golden):
    import kikuchipy_b200 as kb

    z = golden("projection.npz")
    d = kb.direction_cosines(z["gnomonic_bounds"], z["pcz"], int(z["nrows"]), int(z["ncols"]), z["om"])
    assert np.abs(d - z["dc"]).max() < 1e-14
    assert np.allclose(np.linalg.norm(d, axis=1), 1.0, atol=1e-14)
