"""Generate the committed golden vectors from the REFERENCE's own modules.

Run in the build container (``/root/reference`` mounted):

    python tests/golden/make_golden.py

Writes small ``.npz`` fixtures next to this file.  Everything stored here is an
output of reference code executed in place through ``oracle.ref_loader`` (the
metric classes and ``orientation_similarity_map``), plus the nine bundled nickel
patterns (input fixture of BASELINE.json config 1).  ``tests/test_oracle.py``
checks the NumPy restatement against these on every run; the ``-m gpu`` parity
tests check the CUDA path against them through the C ABI.
"""

from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

from oracle import di_oracle, ref_loader  # noqa: E402


def dummy_signal() -> np.ndarray:
    # values: /root/reference/conftest.py:179-187
    # fmt: off
    a = np.array([
        5, 6, 5, 7, 6, 5, 6, 1, 0, 9, 7, 8, 7, 0, 8, 8, 7, 6, 0, 3, 3, 5, 2,
        9, 3, 3, 9, 8, 1, 7, 6, 4, 8, 8, 2, 2, 4, 0, 9, 0, 1, 0, 2, 2, 5, 8,
        6, 0, 4, 7, 7, 7, 6, 0, 4, 1, 6, 3, 4, 0, 1, 1, 0, 5, 9, 8, 4, 6, 0,
        2, 9, 2, 9, 4, 3, 6, 5, 6, 2, 5, 9], dtype=np.uint8)
    # fmt: on
    return a.reshape((3, 3, 3, 3))


def main() -> None:
    if not ref_loader.available():
        raise SystemExit("/root/reference is not mounted; cannot regenerate goldens")
    _, NCC, NDP, ncc_single = ref_loader.load_metrics()
    ref_osm = ref_loader.load_osm()

    # ---------------- config 1: nickel 9 x 1000 ---------------------------
    nickel = ref_loader.nickel_ebsd_small()  # (3,3,60,60) uint8
    dict1k = di_oracle.synthetic_dictionary(1000, (60, 60), seed=2)
    out = {"nickel": nickel}
    for name, cls in (("ncc", NCC), ("ndp", NDP)):
        m = cls(n_experimental_patterns=9, n_dictionary_patterns=1000)
        sim = m(nickel, dict1k)  # reference __call__: prepare + reshape + match
        out[f"{name}_sim_f32"] = np.asarray(sim)
        m64 = cls(n_experimental_patterns=9, n_dictionary_patterns=1000, dtype=np.float64)
        out[f"{name}_sim_f64"] = np.asarray(m64(nickel, dict1k))
        # self-similarity of the nine patterns (known answers quoted in SURVEY 8c)
        ms = cls(n_experimental_patterns=9, n_dictionary_patterns=9)
        out[f"{name}_self_f32"] = np.asarray(ms(nickel, nickel.reshape(9, 60, 60)))
    # masked variant (circular signal mask), NCC + NDP
    mask = di_oracle.circular_signal_mask((60, 60))
    for name, cls in (("ncc", NCC), ("ndp", NDP)):
        m = cls(n_experimental_patterns=9, n_dictionary_patterns=1000, signal_mask=mask)
        out[f"{name}_sim_masked_f32"] = np.asarray(m(nickel, dict1k))
    # navigation-masked variant
    nav = np.zeros((3, 3), dtype=bool)
    nav[1, 1] = True
    nav[0, 2] = True
    m = NCC(n_experimental_patterns=9, n_dictionary_patterns=1000, navigation_mask=nav)
    out["ncc_sim_navmask_f32"] = np.asarray(m(nickel, dict1k))
    out["nav_mask"] = nav
    out["signal_mask"] = mask
    # prepared (normalised) experimental rows from the reference
    m = NCC(n_experimental_patterns=9, n_dictionary_patterns=1000)
    out["ncc_prepared_exp"] = np.asarray(m.prepare_experimental(nickel))
    m = NDP(n_experimental_patterns=9, n_dictionary_patterns=1000)
    out["ndp_prepared_exp"] = np.asarray(m.prepare_experimental(nickel))
    np.savez_compressed(os.path.join(HERE, "config1_nickel_x_1000.npz"), **out)

    # ---------------- dummy signal 3x3|3x3 ----------------------------------
    ds = dummy_signal()
    dsd = ds.reshape(-1, 3, 3)
    out = {"dummy": ds}
    smask = np.array([[0, 0, 0], [0, 1, 0], [0, 0, 0]], dtype=bool)
    for name, cls in (("ncc", NCC), ("ndp", NDP)):
        m = cls(n_experimental_patterns=9, n_dictionary_patterns=9)
        out[f"{name}_sim"] = np.asarray(m(ds, dsd))
        m = cls(n_experimental_patterns=9, n_dictionary_patterns=9, signal_mask=smask)
        out[f"{name}_sim_masked"] = np.asarray(m(ds, dsd))
        m = cls(n_experimental_patterns=9, n_dictionary_patterns=9, signal_mask=smask,
                dtype=np.float64)
        out[f"{name}_sim_masked_f64"] = np.asarray(m(ds, dsd))
    out["signal_mask"] = smask
    np.savez_compressed(os.path.join(HERE, "dummy_signal.npz"), **out)

    # ---------------- repr + numba KAT --------------------------------------
    exp = np.linspace(0, 0.5, 100, dtype=np.float32)
    sim = np.linspace(0.5, 1, 100, dtype=np.float32)
    exp -= np.mean(exp)
    kat = float(ncc_single(exp, sim.copy(), np.square(exp).sum()))
    reprs = {
        "ncc_repr": repr(NCC(1, 1)),
        "ndp_repr": repr(NDP(1, 1)),
        "ncc_repr_masks": repr(
            NCC(9, 9, navigation_mask=nav, signal_mask=smask, dtype=np.float64, rechunk=True)
        ),
    }
    np.savez_compressed(
        os.path.join(HERE, "metric_kat.npz"),
        ncc_single_kat=np.float64(kat),
        **{k: np.array(v) for k, v in reprs.items()},
    )

    # ---------------- OSM goldens --------------------------------------------
    rng = np.random.default_rng(7)
    out = {}
    # the 3x4, k=3 case quoted in SURVEY.md 8c
    idx34 = np.array(
        [[5, 3, 4], [5, 3, 4], [5, 1, 0], [1, 1, 5], [5, 0, 2], [4, 0, 4], [0, 2, 4],
         [1, 2, 1], [4, 1, 5], [2, 2, 3], [3, 3, 3], [5, 4, 4]]
    )
    out["idx34"] = idx34
    out["osm34"] = ref_osm(ref_loader.FakeXmap(idx34, (3, 4)))
    out["osm34_norm_n2"] = ref_osm(ref_loader.FakeXmap(idx34, (3, 4)), n_best=2, normalize=True)
    out["osm34_from1"] = ref_osm(ref_loader.FakeXmap(idx34, (3, 4)), from_n_best=1)
    # random 17x23 map, k=20, indices drawn from a small pool so intersections occur
    idx = rng.integers(0, 60, (17 * 23, 20))
    out["idx_17x23"] = idx
    out["osm_17x23"] = ref_osm(ref_loader.FakeXmap(idx, (17, 23)))
    out["osm_17x23_n7_norm"] = ref_osm(ref_loader.FakeXmap(idx, (17, 23)), n_best=7, normalize=True)
    out["osm_17x23_from15"] = ref_osm(ref_loader.FakeXmap(idx, (17, 23)), from_n_best=15)
    # custom footprint (8-neighbourhood, centre is truthy index 4)
    fp = np.ones((3, 3), dtype=int)
    out["footprint8"] = fp
    out["osm_17x23_fp8"] = ref_osm(
        ref_loader.FakeXmap(idx, (17, 23)), footprint=fp, center_index=4
    )
    # identical lists -> OSM == keep_n (reference test_orientation_similarity_map.py:27-41)
    tiled = np.tile(np.arange(5), (100, 1))
    out["idx_tiled"] = tiled
    out["osm_tiled"] = ref_osm(ref_loader.FakeXmap(tiled, (10, 10)))
    out["osm_tiled_norm"] = ref_osm(ref_loader.FakeXmap(tiled, (10, 10)), normalize=True)
    np.savez_compressed(os.path.join(HERE, "osm.npz"), **out)

    # ---------------- driver-level goldens (reference metric + restated driver) ---
    # similarity blocks come from the reference metric classes; selection/merge is
    # the restated driver (the reference driver itself cannot run here).
    exp = di_oracle.synthetic_experimental(64, (60, 60), seed=1)
    dic = di_oracle.synthetic_dictionary(4096, (60, 60), seed=2)
    m = NCC(n_experimental_patterns=64, n_dictionary_patterns=4096)
    sim = np.asarray(m(exp, dic))
    idx = di_oracle.argtopk(sim, 20)
    sc = di_oracle.topk(sim, 20)
    pexp, planted = di_oracle.planted_experimental(dic, 64, seed=3)
    simp = np.asarray(m(pexp, dic))
    np.savez_compressed(
        os.path.join(HERE, "driver_64x4096.npz"),
        idx=idx, scores=sc, planted_j=planted,
        planted_idx=di_oracle.argtopk(simp, 20), planted_scores=di_oracle.topk(simp, 20),
    )
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
