"""merge_crystal_maps: the oracle against the hard-coded expectations of the
reference's tests (/root/reference/tests/test_indexing/test_merge_crystal_maps.py), the host
mirror's argument handling on CPU (oracle standing in for the device call), and - marked ``gpu`` -
the CUDA path against the oracle through the C ABI."""

import warnings

import numpy as np
import pytest

from oracle import merge_oracle as mo

import kikuchipy_b200 as kb
from kikuchipy_b200 import merge_maps as mm


class XMap:
    """Duck-typed crystal map with the attributes merge_crystal_maps reads (what the reference's
    ``get_single_phase_xmap`` fixture builds, /root/reference/conftest.py:313-340)."""

    def __init__(self, nav_shape, rotations_per_point=5, prop_names=("scores", "simulation_indices"),
                 name="a", space_group=225, phase_id=0, rng=None, scores=None, is_in_data=None):
        rng = rng or np.random.default_rng(0)
        self.shape = tuple(nav_shape)
        size = int(np.prod(nav_shape))
        self.is_in_data = np.ones(size, dtype=bool) if is_in_data is None else is_in_data
        n = int(self.is_in_data.sum())
        ds = (n,) + ((rotations_per_point,) if rotations_per_point > 1 else ())
        q = rng.normal(size=ds + (4,))
        self.rotations = q / np.linalg.norm(q, axis=-1, keepdims=True)
        self.phase_id = np.ones(n, dtype=np.int64) * phase_id
        self.phases = {int(phase_id): mm.SimplePhase(name, space_group)}
        self.prop = {
            prop_names[0]: np.ones(ds, dtype=np.float32) if scores is None else scores,
            prop_names[1]: np.arange(int(np.prod(ds))).reshape(ds),
        }
        self.rotations_per_point = rotations_per_point
        self.dx = self.dy = 1.0
        self.scan_unit = "px"

    @property
    def scores(self):
        return self.prop["scores"]

    def sub(self, keep):
        """Map restricted to the points where ``keep`` (flat bool) is True: like ``xmap[keep]``."""
        out = XMap.__new__(XMap)
        out.__dict__.update(self.__dict__)
        out.is_in_data = np.ones(int(keep.sum()), dtype=bool)
        out.rotations = self.rotations[keep]
        out.phase_id = self.phase_id[keep]
        out.prop = {k: v[keep] for k, v in self.prop.items()}
        out.shape = (int(keep.sum()),)
        return out


class OracleContext:
    """Stands in for the device call on CPU: same arguments as Context.merge_crystal_maps."""

    def merge_crystal_maps(self, scores, rotations, simulation_indices, point_rows, not_indexed,
                           map_size, mean_n_best, sign, idx_as_double):
        n = scores[0].shape[1]
        maps, masks = [], []
        for k in range(len(scores)):
            sq = (lambda a: a[:, 0] if n == 1 else a)
            m = {"scores": sq(scores[k]), "rotations": rotations[k][:, 0] if n == 1 else rotations[k],
                 "phase_id": np.where(not_indexed[k], -1, 0) if not_indexed[k] is not None else np.zeros(scores[k].shape[0])}
            if simulation_indices is not None:
                m["simulation_indices"] = sq(simulation_indices[k])
            maps.append(m)
            rows = point_rows[k]
            masks.append(None if rows is None and not idx_as_double else
                         (np.ones(map_size, bool) if rows is None else rows >= 0))
        out = mo.merge_arrays(maps, masks, map_size, mean_n_best, sign > 0, simulation_indices is not None)
        out["scores"] = out["scores"].reshape(map_size, n)
        out["rotations"] = out["rotations"].reshape(map_size, n, 4)
        if simulation_indices is not None:
            out["simulation_indices"] = out["simulation_indices"].reshape(map_size, n)
        return out


def _contexts():
    return [pytest.param("oracle", id="host-logic"), pytest.param("gpu", marks=pytest.mark.gpu, id="gpu")]


def _ctx(kind):
    return OracleContext() if kind == "oracle" else kb.default_context()


# ---- reference expectations (test_merge_crystal_maps.py:289-370) --------------------------------
MEAN_N_BEST_CASES = [
    ((2,), 1, 1, [[1, 1], [2, 1]], [[0, 2], [3, 1]]),
    ((1, 2), 1, 1, [[1, 1], [2, 1]], [[0, 2], [3, 1]]),
    ((1, 3), 1, 1, [[1, 1, 1], [2, 1, 1], [3, 1, 1]], [[0, 3, 6], [4, 1, 7], [8, 2, 5]]),
    ((2, 1), 2, 2, [[1, 1, 1, 1], [2, 2, 1, 1]], [[0, 4, 1, 5], [6, 7, 2, 3]]),
    ((3, 2), 1, 1,
     [[1, 1, 1], [1, 1, 1], [2, 1, 1], [2, 1, 1], [3, 1, 1], [3, 1, 1]],
     [[0, 6, 12], [1, 7, 13], [8, 2, 14], [9, 3, 15], [16, 4, 10], [17, 5, 11]]),
]


def _mean_n_best_maps(nav_shape, rot_per_point, n_phases):
    maps = []
    for i in range(n_phases):
        x = XMap(nav_shape, rot_per_point, name=str(i), phase_id=0)
        if len(nav_shape) == 1 or min(nav_shape) == 1:
            x.prop["scores"][i] += i  # xmap[i].scores += i (orix drops axes of length one)
        else:  # xmap[i] selects map ROW i
            nx = nav_shape[1]
            x.prop["scores"][i * nx:(i + 1) * nx] += i
        maps.append(x)
    return maps


@pytest.mark.parametrize("nav_shape, rpp, mnb, want_scores, want_idx", MEAN_N_BEST_CASES)
def test_oracle_reference_mean_n_best_tables(nav_shape, rpp, mnb, want_scores, want_idx):
    n_phases = np.shape(want_scores)[-1] // rpp
    maps = _mean_n_best_maps(nav_shape, rpp, n_phases)
    arrs = [{"scores": x.prop["scores"], "rotations": x.rotations, "phase_id": x.phase_id,
             "simulation_indices": x.prop["simulation_indices"]} for x in maps]
    out = mo.merge_arrays(arrs, [None] * n_phases, int(np.prod(nav_shape)), mnb, None, True)
    assert np.allclose(out["merged_scores"], want_scores)
    assert np.array_equal(out["merged_simulation_indices"], want_idx)


@pytest.mark.parametrize("kind", _contexts())
@pytest.mark.parametrize("nav_shape, rpp, mnb, want_scores, want_idx", MEAN_N_BEST_CASES)
def test_mean_n_best_tables(kind, nav_shape, rpp, mnb, want_scores, want_idx):
    n_phases = np.shape(want_scores)[-1] // rpp
    maps = _mean_n_best_maps(nav_shape, rpp, n_phases)
    merged = kb.merge_crystal_maps(maps, mean_n_best=mnb, simulation_indices_prop="simulation_indices",
                                   context=_ctx(kind))
    assert len(merged.phases) == n_phases
    assert np.allclose(merged.merged_scores, want_scores)
    assert np.array_equal(merged.merged_simulation_indices, want_idx)
    assert merged.merged_simulation_indices.dtype == np.int64
    assert merged.shape == tuple(nav_shape)


@pytest.mark.parametrize("kind", _contexts())
@pytest.mark.parametrize(
    "map_shape, rpp, names, mnb",
    [((4, 3), 10, ["a", "b"], 5), ((5, 4), 1, ["a", "b", "c"], 1), ((3, 4), 5, ["austenite", "ferrite"], 4),
     ((4, 5), 1, ["al", "cu", "si"], 1), ((3,), 10, ["a", "b"], 1), ((4,), 1, ["al", "cu", "si"], 1)],
)
def test_merge_expected_winner(kind, map_shape, rpp, names, mnb):
    """test_merge_crystal_maps.py:27-186: every map is the best one in one point (the diagonal)."""
    rng = np.random.default_rng(3)
    size = int(np.prod(map_shape))
    nx = map_shape[1] if len(map_shape) == 2 else 0
    ds = (size,) + ((rpp,) if rpp > 1 else ())
    want_pid = np.zeros(size)
    want_scores = np.ones(ds)
    want_idx = np.arange(int(np.prod(ds))).reshape(ds)
    maps = []
    for i, name in enumerate(names):
        x = XMap(map_shape, rpp, ("scores", "sim_idx"), name, phase_id=i, rng=rng)
        j = i * (1 + nx)
        x.prop["scores"][j] += i + 1
        maps.append(x)
        want_pid[j] = i
        want_scores[j] = x.prop["scores"][j]
        if i == 0:
            want_rot = x.rotations.copy()
        else:
            want_rot[j] = x.rotations[j]
    merged = kb.merge_crystal_maps(maps, mean_n_best=mnb, scores_prop="scores", simulation_indices_prop="sim_idx",
                                   context=_ctx(kind))
    assert merged.shape == tuple(map_shape) and merged.size == size
    assert np.array_equal(merged.phase_id, want_pid)
    assert np.array_equal(merged.prop["scores"], want_scores)
    assert merged.prop["scores"].dtype == np.float32
    assert np.array_equal(merged.prop["sim_idx"], want_idx)
    assert merged.prop["sim_idx"].dtype == np.int32
    assert np.array_equal(merged.rotations, want_rot)
    assert merged.prop["merged_scores"].shape == (size, rpp * len(names))
    assert merged.prop["merged_sim_idx"].shape == (size, rpp * len(names))
    assert [p.name for p in merged.phases.values()] == names


@pytest.mark.parametrize("kind", _contexts())
def test_mean_n_best_varying_scores_and_lower_is_better(kind):
    """test_merge_crystal_maps.py:372-391 and :218-242."""
    a, b = XMap((2, 3), 3, name="a"), XMap((2, 3), 3, name="b")
    a.prop["scores"][0] = [1, 2, 2.1]
    b.prop["scores"][0] = [1, 1.9, 3]
    b.prop["scores"][1] = 2.0
    ctx = _ctx(kind)
    assert np.array_equal(kb.merge_crystal_maps([a, b], mean_n_best=2, context=ctx).phase_id, [0, 1, 0, 0, 0, 0])
    assert np.array_equal(kb.merge_crystal_maps([a, b], mean_n_best=3, context=ctx).phase_id, [1, 1, 0, 0, 0, 0])
    a, b = XMap((5, 6), 5, name="a"), XMap((5, 6), 5, name="b", phase_id=1)
    b.prop["scores"][3] = 0
    want = np.zeros(30)
    want[3] = 1
    m = kb.merge_crystal_maps([a, b], greater_is_better=False, simulation_indices_prop="simulation_indices", context=ctx)
    assert np.array_equal(m.phase_id, want)
    # negative mean_n_best means lower is better (:184-186)
    assert np.array_equal(kb.merge_crystal_maps([a, b], mean_n_best=-1, context=ctx).phase_id, want)


def _masked_maps():
    """test_merge_crystal_maps.py:451-472: integer scores 0..11, second map 1 at (0, 0)."""
    a = XMap((3, 4), 1, name="a", scores=np.arange(12))
    b = XMap((3, 4), 1, name="b", phase_id=1, scores=np.arange(12), rng=np.random.default_rng(9))
    b.prop["simulation_indices"] = b.prop["simulation_indices"] + 12
    b.prop["scores"][0] = 1
    return a, b


@pytest.mark.parametrize("kind", _contexts())
def test_navigation_masks_reference_tables(kind):
    """test_merge_crystal_maps.py:473-590."""
    ctx = _ctx(kind)
    a, b = _masked_maps()
    m3 = kb.merge_crystal_maps([a, b], simulation_indices_prop="simulation_indices", context=ctx)
    assert np.array_equal(m3.phase_id, [1] + [0] * 11)
    assert np.array_equal(m3.scores, [1, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11])
    assert np.array_equal(m3.simulation_indices, [12, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11])
    assert m3.scores.dtype == a.prop["scores"].dtype

    mask1 = np.ones((3, 4), dtype=bool)
    mask1[1:, 1:] = False
    mask2 = ~mask1
    m5 = kb.merge_crystal_maps([a.sub(~mask1.ravel()), b.sub(~mask2.ravel())], navigation_masks=[mask1, mask2],
                               simulation_indices_prop="simulation_indices", context=ctx)
    assert np.array_equal(m5.phase_id, [1, 1, 1, 1, 1, 0, 0, 0, 1, 0, 0, 0])
    assert np.array_equal(m5.simulation_indices, [12, 13, 14, 15, 16, 5, 6, 7, 20, 9, 10, 11])
    assert m5.merged_simulation_indices.dtype == np.float64  # NaN where a map has no point
    assert np.isnan(m5.merged_simulation_indices[:, 1]).all() and np.isnan(m5.merged_scores[:, 1]).all()

    m6 = kb.merge_crystal_maps([a.sub(mask1.ravel()), b], navigation_masks=[~mask1, None], context=ctx)
    assert np.array_equal(m6.phase_id, [1, 0, 0, 0, 0, 1, 1, 1, 0, 1, 1, 1])
    m7 = kb.merge_crystal_maps([a.sub(~mask1.ravel()), b], navigation_masks=[mask1, None], context=ctx)
    assert np.array_equal(m7.phase_id, [1, 1, 1, 1, 1, 0, 0, 0, 1, 0, 0, 0])

    # masks taken from is_in_data when the maps do not hold every point (:96-105)
    inner = np.zeros((3, 4), dtype=bool)
    inner[1:, 1:] = True
    a4 = XMap((3, 4), 1, name="a", scores=np.arange(12), is_in_data=inner.ravel().copy())
    a4.prop["scores"] = np.arange(12)[inner.ravel()]
    a4.prop["simulation_indices"] = np.arange(12)[inner.ravel()]
    b4 = XMap((3, 4), 1, name="b", phase_id=1, is_in_data=inner.ravel().copy(), rng=np.random.default_rng(2))
    b4.prop["scores"] = np.arange(12)[inner.ravel()]
    b4.prop["simulation_indices"] = np.arange(12)[inner.ravel()] + 12
    with pytest.raises(ValueError, match="All-NaN slice encountered"):
        kb.merge_crystal_maps([a4, b4], simulation_indices_prop="simulation_indices", context=ctx)


@pytest.mark.parametrize("kind", _contexts())
def test_phase_list_bookkeeping(kind):
    """Duplicate names with different space groups are renamed with a warning
    (test_merge_crystal_maps.py:244-287); equal phases collapse to one id (:592-604)."""
    ctx = _ctx(kind)
    for names, want in [(["a"] * 3, ["a", "a1", "a2"]), (["1"] * 5, ["1", "11", "12", "13", "14"])]:
        maps = []
        for i, name in enumerate(names):
            x = XMap((5, 6), 5, name=name, space_group=i + 1, phase_id=i)
            x.prop["scores"][i * 7] += i + 1
            maps.append(x)
        with pytest.warns(UserWarning, match=f"There are duplicates of phase '{names[0]}'"):
            merged = kb.merge_crystal_maps(maps, simulation_indices_prop="simulation_indices", context=ctx)
        assert [p.name for p in merged.phases.values()] == want
    a, b = _masked_maps()
    b.phases = {1: mm.SimplePhase("a", 225)}
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        assert np.array_equal(kb.merge_crystal_maps([a, b], context=ctx).phase_id, np.zeros(12))


def test_argument_errors():
    """test_merge_crystal_maps.py:393-449, :606-640 (no device call is reached)."""
    ctx = OracleContext()
    with pytest.raises(ValueError, match=r"Crystal maps \(and/or navigation masks"):
        kb.merge_crystal_maps([XMap((4, 3)), XMap((3, 4))], context=ctx)
    with pytest.raises(ValueError, match="Crystal maps must have the"):
        kb.merge_crystal_maps([XMap((2, 3), 3), XMap((2, 3), 4)], context=ctx)
    a, b = _masked_maps()
    mask1 = np.ones((3, 4), dtype=bool)
    mask1[1:, 1:] = False
    with pytest.raises(ValueError, match="Number of crystal maps and navigation "):
        kb.merge_crystal_maps([a, b], navigation_masks=mask1, context=ctx)
    with pytest.raises(ValueError, match=r"Crystal maps \(and/or navigation masks"):
        kb.merge_crystal_maps([a.sub(~mask1.ravel()), b.sub(np.arange(12) < 5)], context=ctx)
    with pytest.raises(ValueError, match="0. navigation mask does not have as "):
        kb.merge_crystal_maps([a, b], navigation_masks=[mask1, ~mask1], context=ctx)
    with pytest.raises(ValueError, match="0. navigation mask must be a NumPy array or 'None'"):
        kb.merge_crystal_maps([a, b], navigation_masks=[[1], None], context=ctx)
    # refined maps: one score but ten simulation indices per point (:395-449)
    r1, r2 = XMap((3, 3), 1, name="a"), XMap((3, 3), 1, name="b")
    r1.prop["simulation_indices"] = np.zeros((9, 10), dtype=int)
    r2.prop["simulation_indices"] = np.zeros((9, 10), dtype=int)
    r1.prop["scores"][0] = 3
    r2.prop["scores"] *= 2
    m = kb.merge_crystal_maps([r1, r2], context=ctx)
    assert "simulation_indices" not in m.prop and "merged_simulation_indices" not in m.prop
    with pytest.raises(ValueError, match="Cannot merge maps with more"):
        kb.merge_crystal_maps([r1, r2], simulation_indices_prop="simulation_indices", context=ctx)


def _random_case(seed, map_shape, n_scores, n_maps, dtype, masked, ties):
    rng = np.random.default_rng(seed)
    size = int(np.prod(map_shape))
    maps, masks = [], []
    cover = np.zeros(size, dtype=bool)
    for k in range(n_maps):
        keep = np.ones(size, dtype=bool)
        if masked and k > 0:
            keep = rng.random(size) < 0.6
        if masked and k == n_maps - 1:
            keep |= ~cover  # every point in at least one map
        cover |= keep
        n = int(keep.sum())
        sc = rng.random((n, n_scores))
        if ties:
            sc = np.round(sc * 4) / 4
        sc = -np.sort(-sc, axis=1).astype(dtype)
        if n_scores > 2:
            sc[rng.random(n) < 0.05, -1] = np.nan  # a NaN score inside a map sorts last too
        x = XMap((n,), n_scores, name=f"p{k}", phase_id=k, rng=rng, scores=sc if n_scores > 1 else sc[:, 0])
        x.prop["simulation_indices"] = rng.integers(0, 5000, (n, n_scores) if n_scores > 1 else (n,))
        if not masked:
            x.shape = tuple(map_shape)
            x.phase_id[rng.random(n) < 0.1] = -1
        maps.append(x)
        masks.append(~keep.reshape(map_shape) if masked else None)
    return maps, (masks if masked else None)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("masked", [False, True])
@pytest.mark.parametrize("map_shape, n_scores, n_maps, mnb, ties",
                         [((7, 9), 20, 2, 1, False), ((40, 50), 20, 3, 7, True), ((33,), 1, 5, 1, True),
                          ((6, 5), 50, 4, -50, False), ((3, 4), 64, 32, 3, True), ((200, 200), 20, 2, 20, False)])
def test_gpu_equals_oracle_random(dtype, masked, map_shape, n_scores, n_maps, mnb, ties):
    """Bit-exact against the oracle (phase choice incl. float32 nanmean ties, stable order, NaNs)."""
    maps, masks = _random_case(11, map_shape, n_scores, n_maps, dtype, masked, ties)
    kw = dict(mean_n_best=mnb, simulation_indices_prop="simulation_indices", navigation_masks=masks)
    got = kb.merge_crystal_maps(maps, context=kb.default_context(), **kw)
    want = kb.merge_crystal_maps(maps, context=OracleContext(), **kw)
    assert np.array_equal(got.phase_id, want.phase_id)
    for name in ("scores", "merged_scores", "simulation_indices", "merged_simulation_indices"):
        g, w = got.prop[name], want.prop[name]
        assert g.dtype == w.dtype and g.shape == w.shape, name
        assert np.array_equal(g, w, equal_nan=True), name
    assert np.array_equal(got.rotations, want.rotations)
