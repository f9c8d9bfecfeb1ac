"""CPU-only checks: the C-ABI library loads and exports every symbol include/kdi.h declares, the
host-side mirror of the reference interface behaves like the reference (repr, dtype rejection,
argument validation messages), and the multi-GPU plumbing (shard bounds, all-gather layout) is
right under gloo with world_size 2.  No compute call is made without a GPU."""

import os
import re
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))

import kikuchipy_b200 as kb  # noqa: E402
from kikuchipy_b200 import _lib  # noqa: E402
from oracle import di_oracle as orc  # noqa: E402


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "kdi.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(kdi_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _header_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"libkdi.so lacks {name}"
        assert name in _lib.SIGNATURES, f"ctypes binding lacks {name}"
    assert lib.kdi_version() == 100


def test_certificate_bound_formula():
    """kdi_certificate_bound (a host function, no device needed) is the formula include/kdi.h states for
    KDI_OPT_CERT_STRICT: operand rounding by Cauchy-Schwarz, 8 ulp per 16-deep accumulation step (5.25 measured,
    profiles/r2_mma_accumulate_probe.txt), the float32 summation of the exact score; row length padded to 64."""
    lib = _lib.load()
    ulp = 2.0 ** -23
    for s_eff, kp in ((3600, 3648), (11287, 11328), (6400, 6400), (9, 64), (64, 64), (65, 128)):
        for code, u in ((0, 2.0 ** -11), (1, 2.0 ** -8)):
            steps = kp // 16
            want = (u * (2 + u) + 8 * steps * ulp) * (1 + 4e-6) + (0.5 * steps + 8) * ulp + 1e-6
            got = lib.kdi_certificate_bound(code, s_eff)
            assert abs(got - want) <= 1e-6 * want, (s_eff, code, got, want)
    assert 1.1e-3 < lib.kdi_certificate_bound(0, 3600) < 1.3e-3
    assert lib.kdi_certificate_bound(0, 14400) > lib.kdi_certificate_bound(0, 3600)  # grows with the K loop
    assert lib.kdi_candidate_capacity(20) == 32 and lib.kdi_candidate_capacity(52) == 64
    assert lib.kdi_candidate_capacity(104) == 128 and lib.kdi_candidate_capacity(105) == 0


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.KdiError, match="no CUDA device available"):
        kb.Context(0)
    with pytest.raises(_lib.KdiError):
        kb.dictionary_indexing(np.zeros((2, 3, 3), np.uint8), np.ones((4, 3, 3), np.float32), verbose=False)


def test_metric_repr_and_dtype_validation():
    # reference tests/test_indexing/test_similarity_metrics.py:28-39
    m = kb.NormalizedCrossCorrelationMetric(1, 1)
    assert repr(m) == (
        "NormalizedCrossCorrelationMetric: float32, greater is better, rechunk: False, "
        "navigation mask: False, signal mask: False"
    )
    m = kb.NormalizedDotProductMetric(1, 1, dtype=np.float16)
    with pytest.raises(ValueError, match="Data type float16 not among supported data types"):
        m.raise_error_if_invalid()
    assert kb.NormalizedDotProductMetric().sign == 1
    m = kb.NormalizedCrossCorrelationMetric(signal_mask=np.zeros((3, 3), bool), rechunk=True)
    assert "rechunk: True" in repr(m) and "signal mask: True" in repr(m)
    m.n_experimental_patterns = 7
    m.n_dictionary_patterns = 9
    assert (m.n_experimental_patterns, m.n_dictionary_patterns) == (7, 9)
    assert issubclass(kb.NormalizedCrossCorrelationMetric, kb.SimilarityMetric)


def test_argument_validation_messages(dummy_array):
    # reference tests/test_indexing/test_dictionary_indexing.py:90-117,147-164; signals/ebsd.py:1931-1964
    dic = dummy_array.reshape(-1, 3, 3).astype(np.float32)
    with pytest.raises(ValueError, match=r"The navigation mask shape \(3, 2\) and the signal's navigation"):
        kb.dictionary_indexing(dummy_array, dic, navigation_mask=np.zeros((3, 2), bool), verbose=False)
    with pytest.raises(ValueError, match="The navigation mask must allow for indexing of at least one"):
        kb.dictionary_indexing(dummy_array, dic, navigation_mask=np.ones((3, 3), bool), verbose=False)
    with pytest.raises(ValueError, match="The signal mask must be a NumPy array"):
        kb.dictionary_indexing(dummy_array, dic, signal_mask=[[0, 0, 0]] * 3, verbose=False)
    with pytest.raises(ValueError, match=r"Experimental \(3, 3\) and dictionary \(3, 2\) signal shapes must"):
        kb.dictionary_indexing(dummy_array, dic[:, :, :2], verbose=False)
    with pytest.raises(ValueError, match="must be either of"):
        kb.dictionary_indexing(dummy_array, dic, metric="nonexistent", verbose=False)
    with pytest.raises(ValueError, match="Data type float16 not among supported"):
        kb.dictionary_indexing(dummy_array, dic, dtype=np.float16, verbose=False)
    with pytest.raises(ValueError, match="only one navigation dimension"):
        kb.dictionary_indexing(dummy_array, dic.reshape(3, 3, 3, 3), verbose=False)


def test_shard_bounds_cover_dictionary():
    for n in (1, 7, 100_000, 300_000):
        for world in (1, 2, 3, 8):
            b = [kb.shard_bounds(n, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [e - s for s, e in b]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        kb.shard_bounds(10, 2, 2)


class _OracleStages:
    """CPU stand-in for the GPU stages of ``run_sharded_pipeline`` (the oracle is the checker here):
    same interface as ``kikuchipy_b200.distributed._KdiStages`` on CPU torch tensors."""

    KC = 8

    def __init__(self, exp, dic_shard, start, flag_every=0):
        import torch

        self.t = torch
        e = orc.prepare_experimental(exp, "ncc", exp.shape[0])
        d = orc.prepare_dictionary(dic_shard.reshape(dic_shard.shape[0], -1), "ncc")
        self.sim = orc.match(e, d)  # (rows, shard rows) exact scores
        self.start, self.flag_every = start, flag_every

    @staticmethod
    def _rank(scores, idx, k):
        order = np.lexsort((idx, -scores), axis=1)[:, :k]  # score desc, index asc
        return np.take_along_axis(idx, order, 1), np.take_along_axis(scores, order, 1)

    def candidates(self, pad_rows):
        rows, n = self.sim.shape
        kc = self.KC
        # "tensor-core" scores: the exact ones plus a small deterministic perturbation
        approx = self.sim + 1e-6 * np.cos(np.arange(rows * n, dtype=np.float32)).reshape(rows, n)
        gidx = np.broadcast_to(np.arange(n, dtype=np.int64) + self.start, (rows, n))
        i, s = self._rank(approx, gidx, min(kc, n))
        a = np.full((max(rows, pad_rows), kc), -np.inf, np.float32)
        g = np.full((max(rows, pad_rows), kc), -1, np.int64)
        a[:rows, : s.shape[1]] = s
        g[:rows, : i.shape[1]] = i
        return self.t.from_numpy(a), self.t.from_numpy(g), kc

    def merge(self, s_all, i_all, k):
        s = np.concatenate(list(s_all.numpy()), axis=1)
        i = np.concatenate(list(i_all.numpy()), axis=1)
        mi, ms = self._rank(s, i, k)
        return self.t.from_numpy(mi.copy()), self.t.from_numpy(ms.copy())

    def rescore_owned(self, gidx, approx, keep_n):
        g = gidx.numpy()
        rows, n = self.sim.shape
        out = np.full((max(rows, g.shape[0]), g.shape[1]), -np.inf, np.float32)
        loc = g[:rows] - self.start
        own = (loc >= 0) & (loc < n)
        r = np.nonzero(own)
        out[r[0], r[1]] = self.sim[r[0], loc[own]]
        return self.t.from_numpy(out)

    def finalize(self, approx, gidx, exact, keep_n, dict_total, row0, rows):
        i, s = self._rank(exact.numpy(), gidx.numpy(), keep_n)
        i[rows:], s[rows:] = -1, -np.inf
        flags = [r for r in range(row0, row0 + rows) if self.flag_every and r % self.flag_every == 0]
        for r in flags:  # a flagged row's provisional result must not survive
            i[r - row0], s[r - row0] = -7, np.nan
        return self.t.from_numpy(i.copy()), self.t.from_numpy(s.copy()), self.t.tensor(flags, dtype=self.t.int32)

    def exact_rows(self, rows, k_local):
        r = rows.numpy().astype(np.int64)
        n = self.sim.shape[1]
        gidx = np.broadcast_to(np.arange(n, dtype=np.int64) + self.start, (r.size, n))
        i, s = self._rank(self.sim[r], gidx, k_local)
        return self.t.from_numpy(i.copy()), self.t.from_numpy(s.copy())


def _gloo_worker(rank, world, port, tmp):
    import torch
    import torch.distributed as dist

    from kikuchipy_b200.distributed import run_sharded_pipeline

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        exp = orc.synthetic_experimental(25, (12, 12), seed=1)
        dic = orc.synthetic_dictionary(301, (12, 12), seed=2)
        start, end = kb.shard_bounds(301, world, rank)
        out = {}
        for name, flag_every in (("plain", 0), ("flagged", 4)):
            st = _OracleStages(exp, dic[start:end], start, flag_every)
            idx, sc = run_sharded_pipeline(st, 25, 5, 301, end - start)
            out[f"idx_{name}"], out[f"sc_{name}"] = idx.numpy(), sc.numpy()
        # layout of the plain all-gather helper: list r is rank r's tensor
        s_all, i_all = kb.gather_topk(torch.full((3, 2), float(rank)), torch.full((3, 2), rank, dtype=torch.int64))
        assert tuple(s_all.shape) == (world, 3, 2) and all(float(s_all[r, 0, 0]) == r for r in range(world))
        np.savez(os.path.join(tmp, f"r{rank}.npz"), **out)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_pipeline_matches_unsharded_gloo(tmp_path, world):
    """all-to-all by row slice -> merge -> all-gather -> owner rescoring -> reduce-scatter ->
    finalize -> all-gather (+ flagged rows), with ragged row slices (25 rows over 2 / 3 ranks)."""
    import torch.multiprocessing as mp

    port = 29500 + (os.getpid() % 2000) + world
    mp.spawn(_gloo_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    exp = orc.synthetic_experimental(25, (12, 12), seed=1)
    dic = orc.synthetic_dictionary(301, (12, 12), seed=2)
    ridx, rsc = orc.dictionary_indexing(exp, dic, keep_n=5)
    z0 = np.load(os.path.join(str(tmp_path), "r0.npz"))
    for r in range(world):
        z = np.load(os.path.join(str(tmp_path), f"r{r}.npz"))
        for name in ("plain", "flagged"):
            assert z[f"idx_{name}"].shape == (25, 5)
            c = orc.compare_topk(ridx, rsc, z[f"idx_{name}"], z[f"sc_{name}"])
            assert c["tie_ok"] and c["max_dscore"] < 1e-6 and c["exact_rows"] == 1.0, (name, c)
            assert np.array_equal(z[f"idx_{name}"], z0[f"idx_{name}"])  # identical on every rank


def test_row_slices_cover_rows():
    from kikuchipy_b200.distributed import row_slice

    for n in (1, 7, 25, 80_000):
        for world in (1, 2, 3, 8):
            sl = [row_slice(n, world, r) for r in range(world)]
            per = sl[0][0]
            assert all(s[0] == per for s in sl) and per * world >= n
            assert sl[0][1] == 0 and sl[-1][2] == n
            assert all(sl[i][2] == sl[i + 1][1] for i in range(world - 1))


def test_get_patterns_argument_checks_without_gpu():
    """Argument handling of the get_patterns mirror that happens before any device call
    (signals/ebsd_master_pattern.py:97-233)."""
    mp = np.zeros((11, 11), np.float32)
    rot = np.tile([1.0, 0, 0, 0], (3, 1))
    with pytest.raises(NotImplementedError, match="float32"):
        kb.get_patterns(mp, mp, rot, direction_cosines=np.ones((4, 3)), dtype_out="uint8")
    with pytest.raises(ValueError, match="either a detector or direction cosines"):
        kb.get_patterns(mp, mp, rot)
    with pytest.raises(ValueError, match="quaternions"):
        kb.get_patterns(mp, mp, np.zeros((3, 3)), direction_cosines=np.ones((4, 3)))
    with pytest.raises(NotImplementedError, match="one navigation dimension"):
        kb.get_patterns(mp, mp, np.zeros((2, 3, 4)), direction_cosines=np.ones((4, 3)))
    with pytest.raises(ValueError, match="detector shape"):
        kb.get_patterns(mp, mp, rot, direction_cosines=np.ones((4, 3)), detector_shape=(3, 3))

    class _Det:  # several projection centres: their number must equal the number of rotations
        navigation_shape = (2, 2)
        shape = (2, 2)
        pc = np.full((2, 2, 3), 0.5)
        om_detector_to_sample = np.eye(3)

    with pytest.raises(ValueError, match="is not \\(1,\\) or equal to `rotations.shape`"):
        kb.get_patterns(mp, mp, rot, _Det())


def test_synthetic_inputs_match_oracle_copies():
    from kikuchipy_b200 import synthetic as syn
    from oracle import projection_oracle as po

    for a, b in zip(syn.synthetic_master_pattern(51, 5), po.synthetic_master_pattern(51, 5)):
        assert np.array_equal(a, b)
    assert np.array_equal(syn.random_rotations(7, 4), po.random_rotations(7, 4))
    assert np.array_equal(syn.tilted_detector_matrix(70.0), po.tilted_detector_matrix(70.0))


def test_bench_cpu_arm_runs_small(monkeypatch):
    sys.path.insert(0, ROOT)
    import bench

    monkeypatch.setattr(bench, "N_DICT", 3000)
    exp, dic = bench.host_inputs(8)
    v, detail = bench.cpu_sample(exp, dic, 100)
    assert v > 0 and detail["sample_patterns"] == 8 and detail["dictionary"] == 3000


def test_inputs_of_neighbouring_rows_without_gpu():
    """Host-side input handling that needs no device: a NORDIF scan container is unwrapped like a
    signal; refinement refuses a map without rotations; merge_crystal_maps substitutes identity
    rotations for indexing results that carry none."""
    from kikuchipy_b200 import indexing, merge_maps
    from kikuchipy_b200.io_nordif import NordifScan

    scan = NordifScan(np.zeros((3, 4, 6, 5), np.uint8), None, None, (1.5, 1.5), {}, {})
    data, nav, sig, steps, unit, xmap = indexing._unwrap(scan)
    assert data is scan.data and nav == (3, 4) and sig == (6, 5) and steps == (1.5, 1.5) and unit == "um" and xmap is None
    res = kb.DictionaryIndexingResult(np.ones((12, 2), np.float32), np.zeros((12, 2), np.int64), None, (3, 4), None,
                                      np.ones(12, bool), 2, "ni")
    with pytest.raises(ValueError, match="no rotations to refine"):
        kb.refine_orientation(scan.data, res, kb.Detector((6, 5)), (np.zeros((5, 5), np.float32),) * 2)
    r = merge_maps._rotation_data(res)
    assert r.shape == (12, 2, 4) and np.all(r[..., 0] == 1) and not r[..., 1:].any()


# ---- the reference's generic driver over custom SimilarityMetric subclasses -------------------------

class _LazyArray:
    """Minimal Dask-like array: slicing and reshape stay lazy, ``compute()`` materialises."""

    def __init__(self, a, log=None):
        self._a, self.log = a, log if log is not None else []
        self.shape, self.chunksize = a.shape, (7,) + a.shape[1:]

    def __getitem__(self, key):
        return _LazyArray(self._a[key], self.log)

    def reshape(self, shape):
        return _LazyArray(self._a.reshape(shape), self.log)

    def compute(self):
        self.log.append(self._a.shape[0])
        return self._a


class _NegEuclid(kb.SimilarityMetric):
    """Toy lower-is-better metric (mean squared difference) with NumPy hooks.  Scores stay below 1:
    the reference initialises the running list with -sign = +1 as "worse than anything"
    (_dictionary_indexing.py:96-99), which only holds for metrics bounded by 1."""

    _allowed_dtypes = [np.float32, np.float64]
    _sign = -1

    def prepare_experimental(self, p):
        p = np.asarray(p, self.dtype).reshape((self.n_experimental_patterns, -1))
        return p if self.navigation_mask is None else p[~self.navigation_mask.ravel()]

    def prepare_dictionary(self, p):
        return np.asarray(p, self.dtype)

    def match(self, e, d):
        class Block:
            def __init__(self, s):
                self.s = s

            def argtopk(self, k, axis=-1):  # Dask: negative k = the k smallest
                return np.argsort(self.s, axis=1, kind="stable")[:, : abs(k)]

            def topk(self, k, axis=-1):
                return np.sort(self.s, axis=1, kind="stable")[:, : abs(k)]

        return Block(((e[:, None, :] - d[None, :, :]) ** 2).mean(-1))


@pytest.mark.parametrize("lazy", [False, True])
@pytest.mark.parametrize("n_per_iteration", [None, 7, 50])
def test_generic_driver_chunks_sign_and_lazy_dictionaries(lazy, n_per_iteration):
    """the path for custom metrics must follow the reference loop
    (_dictionary_indexing.py:94-128): chunking by n_per_iteration, keep_n clamped per chunk, chunk
    offset added, merge by argsort(-sign * scores) - checked with a lower-is-better metric - and a
    lazy dictionary computed one chunk at a time."""
    rng = np.random.default_rng(0)
    exp = rng.random((2, 3, 4, 5)).astype(np.float32)
    dic = rng.random((23, 4, 5)).astype(np.float32)
    log = []
    d_in = _LazyArray(dic, log) if lazy else dic
    res = kb.dictionary_indexing(exp, d_in, metric=_NegEuclid(), keep_n=9, n_per_iteration=n_per_iteration,
                                 verbose=False)
    ssd = ((exp.reshape(6, 1, -1) - dic.reshape(1, 23, -1)) ** 2).mean(-1)
    want = np.argsort(ssd, axis=1, kind="stable")[:, :9]
    assert np.array_equal(res.simulation_indices, want)
    assert np.allclose(res.scores, np.take_along_axis(ssd, want, 1))
    if lazy:
        per = 7 if n_per_iteration in (None, 7) else 23  # default: the dictionary's own chunk size
        assert log and max(log) <= per and sum(log) == 23  # never materialised in one piece


# ---- orix / kikuchipy branches, driven through stand-in modules -------------------------------------

def _install_orix_stub(monkeypatch):
    import types

    orix = types.ModuleType("orix")
    cm = types.ModuleType("orix.crystal_map")
    quat = types.ModuleType("orix.quaternion")

    class Rotation:
        def __init__(self, data):
            self.data = np.array(data, dtype=float)

        @classmethod
        def identity(cls, shape):
            d = np.zeros(tuple(np.atleast_1d(shape)) + (4,))
            d[..., 0] = 1
            return cls(d)

        @property
        def shape(self):
            return self.data.shape[:-1]

        def __getitem__(self, key):
            return Rotation(self.data[key])

        def __setitem__(self, key, value):
            self.data[key] = value.data if isinstance(value, Rotation) else value

        def flatten(self):
            return Rotation(self.data.reshape(-1, 4))

    Rotation.__module__ = "orix.quaternion"

    class CrystalMap:
        def __init__(self, rotations=None, phase_list=None, x=None, y=None, prop=None, is_in_data=None, **kw):
            self.rotations, self.phases, self.x, self.y, self.prop = rotations, phase_list, x, y, prop
            self.is_in_data = np.ones(len(x), bool) if is_in_data is None else is_in_data
            self.scan_unit = "px"

    CrystalMap.__module__ = "orix.crystal_map"

    def create_coordinate_arrays(shape, step_sizes=None):
        shape = tuple(shape) if len(shape) else (1,)
        steps = (1,) * len(shape) if step_sizes is None else step_sizes
        if len(shape) == 1:
            return {"x": np.arange(shape[0]) * steps[0]}, shape[0]
        ny, nx = shape
        return {"x": np.tile(np.arange(nx) * steps[1], ny), "y": np.repeat(np.arange(ny) * steps[0], nx)}, ny * nx

    cm.CrystalMap, cm.create_coordinate_arrays, quat.Rotation = CrystalMap, create_coordinate_arrays, Rotation
    orix.crystal_map, orix.quaternion = cm, quat
    for name, mod in (("orix", orix), ("orix.crystal_map", cm), ("orix.quaternion", quat)):
        monkeypatch.setitem(sys.modules, name, mod)
    return Rotation, CrystalMap


class _Axis:
    def __init__(self, scale, units="um"):
        self.scale, self.units = scale, units


class _AxesManager:
    def __init__(self, nav_shape, sig_shape, scales):
        self.navigation_shape, self.signal_shape = tuple(nav_shape[::-1]), tuple(sig_shape[::-1])
        self.navigation_axes = [_Axis(s) for s in scales[::-1]]


class _Signal:
    """What dictionary_indexing touches of a kikuchipy EBSD signal."""

    def __init__(self, data, nav_dims, scales, xmap=None):
        self.data, self.xmap = data, xmap
        self.axes_manager = _AxesManager(data.shape[:nav_dims], data.shape[nav_dims:], scales)


@pytest.mark.parametrize("keep_n, masked", [(3, False), (3, True), (1, True), (1, False)])
def test_crystal_map_branch_with_orix_stand_in(monkeypatch, keep_n, masked):
    """the CrystalMap return path (_dictionary_indexing.py:141-167) had never run.
    With stand-in orix modules the whole branch executes: coordinate arrays, rotations gathered from
    the dictionary's crystal map, navigation-mask scatter with identity rotations elsewhere, the
    keep_n == 1 squeeze inside the mask branch only, phase list and scan unit."""
    Rotation, CrystalMap = _install_orix_stub(monkeypatch)
    rng = np.random.default_rng(1)
    exp = rng.random((2, 3, 4, 5)).astype(np.float32)
    dic = rng.random((11, 4, 5)).astype(np.float32)
    q = rng.normal(size=(11, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)

    class _Phases:
        names = ["ni"]

    class _DictMap:
        rotations, phases, phases_in_data, shape = Rotation(q), _Phases(), "phase list of the dictionary", (11,)

    nav = None
    if masked:
        nav = np.zeros((2, 3), bool); nav[0, 1] = nav[1, 2] = True
    res = kb.dictionary_indexing(_Signal(exp, 2, (1.5, 0.5)), _Signal(dic, 1, (1,), _DictMap()), metric=_NegEuclid(),
                                 keep_n=keep_n, navigation_mask=nav, verbose=False)
    assert isinstance(res, CrystalMap) and res.phases == "phase list of the dictionary" and res.scan_unit == "um"
    assert np.allclose(res.x, np.tile(np.arange(3) * 0.5, 2)) and np.allclose(res.y, np.repeat(np.arange(2) * 1.5, 3))
    ssd = ((exp.reshape(6, 1, -1) - dic.reshape(1, 11, -1)) ** 2).mean(-1)
    want = np.argsort(ssd, axis=1, kind="stable")[:, :keep_n]
    idx = res.prop["simulation_indices"]
    if masked:
        keep = ~nav.ravel()
        assert np.array_equal(res.is_in_data, keep)
        if keep_n == 1:  # squeezed to 1-D only in this branch (:155-158)
            assert idx.shape == (6,) and res.rotations.data.shape == (6, 4)
            assert np.array_equal(idx[keep], want[keep, 0])
            assert np.allclose(res.rotations.data[keep], q[want[keep, 0]])
            assert np.allclose(res.rotations.data[~keep], [1, 0, 0, 0])
        else:
            assert idx.shape == (6, keep_n) and np.array_equal(idx[keep], want[keep])
            assert np.allclose(res.rotations.data[keep], q[want[keep]])
            assert np.allclose(res.rotations.data[~keep], [1, 0, 0, 0])
    else:
        assert idx.shape == (6, keep_n) and np.array_equal(idx, want)  # (M, 1) stays 2-D without a mask (:164-166)
        assert np.allclose(res.rotations.data, q[want])


def test_gpu_metrics_subclass_the_reference_abc(monkeypatch):
    """Level-1 drop-in: an unmodified kikuchipy only accepts ``metric=`` instances of ITS
    SimilarityMetric (signals/ebsd.py:3067) and then assigns sizes, masks and dtype onto them
    (:3071-3086).  With ``kikuchipy.indexing.SimilarityMetric`` importable - the reference's own class
    when /root/reference is mounted, else a stand-in with the same surface - the GPU metric classes must
    be its subclasses and survive that treatment."""
    import importlib
    import types

    from oracle import ref_loader

    if ref_loader.available():
        ref_abc = ref_loader.load_metrics()[0]
    else:
        ref_abc = type("SimilarityMetric", (kb.similarity_metrics._SimilarityMetricReplica,), {})
    pkg = types.ModuleType("kikuchipy"); pkg.__path__ = []
    indexing = types.ModuleType("kikuchipy.indexing"); indexing.__path__ = []
    indexing.SimilarityMetric = ref_abc
    monkeypatch.setitem(sys.modules, "kikuchipy", pkg)
    monkeypatch.setitem(sys.modules, "kikuchipy.indexing", indexing)
    sm = importlib.reload(kb.similarity_metrics)
    try:
        assert sm.SimilarityMetric is ref_abc
        for cls in (sm.NormalizedCrossCorrelationMetric, sm.NormalizedDotProductMetric):
            metric = cls()
            assert isinstance(metric, ref_abc)
            # what EBSD._prepare_metric does with a metric instance it accepted
            metric.n_experimental_patterns, metric.n_dictionary_patterns = 6, 11
            metric.navigation_mask = np.zeros((2, 3), bool)
            metric.signal_mask = np.zeros((4, 5), bool)
            metric.dtype = np.float32
            metric.raise_error_if_invalid()
            assert repr(metric) == (f"{cls.__name__}: float32, greater is better, rechunk: False, "
                                    "navigation mask: True, signal mask: True")
            metric.dtype = np.float64  # allowed, like in the reference
            metric.raise_error_if_invalid()
            assert repr(metric).startswith(f"{cls.__name__}: float64, greater is better")
            metric.dtype = np.float16  # _similarity_metric.py:244-253
            with pytest.raises(ValueError, match="Data type float16 not among supported data types"):
                metric.raise_error_if_invalid()
    finally:
        monkeypatch.undo()
        importlib.reload(kb.similarity_metrics)


def test_division_by_the_row_norm_through_a_double_reciprocal_is_exact():
    """The prepare kernels compute (x - mean) / norm as float32(float64(x - mean) * (1 / float64(norm)))
    (csrc/kdi_internal.cuh, kdi_div_by_norm: one reciprocal per row instead of an IEEE division per
    pixel).  That must be the correctly rounded float32 quotient - NumPy's - for every operand pair,
    including divisors whose significand is all ones and quotients next to 1."""
    rng = np.random.default_rng(0)
    n = 2_000_000
    for case in range(3):
        c = (rng.standard_normal(n) * rng.choice([1e-3, 1, 30, 1e4], n)).astype(np.float32)
        norm = np.abs(rng.standard_normal(n) * rng.choice([1e-2, 1, 500, 7e4], n)).astype(np.float32) + np.float32(1e-20)
        if case == 1:
            bits = (norm.view(np.uint32) & np.uint32(0xFF800000)) | np.uint32(0x7FFFC0) | rng.integers(0, 64, n).astype(np.uint32)
            norm = bits.view(np.float32)
        if case == 2:
            c = np.nextafter(norm, (np.inf * rng.choice([-1, 1], n)).astype(np.float32))
        want = c / norm
        got = (c.astype(np.float64) * (1.0 / norm.astype(np.float64))).astype(np.float32)
        assert np.array_equal(want.view(np.uint32), got.view(np.uint32))
    with np.errstate(invalid="ignore", divide="ignore"):
        z = (np.float64(np.float32(0.0)) * (1.0 / np.float64(np.float32(0.0)))).astype(np.float32)
    assert np.isnan(z)  # 0 / 0 (constant pattern) stays NaN


_FMA_DIV_C = r"""
#include <math.h>
#include <stdint.h>
#include <string.h>
static uint64_t s = 88172645463325252ULL;
static inline uint64_t rnd(void) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
/* number of operand pairs (out of n) for which the FMA sequence differs from the IEEE quotient;
   norms in [2^e_lo, 2^e_hi], dividends down to 2^-span below the norm */
long kdi_fma_div_mismatches(long n, int e_lo, int e_hi, int span) {
  long bad = 0;
  for (long i = 0; i < n; ++i) {
    uint32_t mn = (uint32_t)(rnd() & 0x7fffff);
    if ((i & 3) == 1) mn = 0x7fffff - (uint32_t)(rnd() & 0xff);
    if ((i & 3) == 2) mn = (uint32_t)(rnd() & 0xff);
    const int en = 127 + e_lo + (int)(rnd() % (uint64_t)(e_hi - e_lo + 1));
    const float nrm = u2f(((uint32_t)en << 23) | mn);
    uint32_t mc = (uint32_t)(rnd() & 0x7fffff);
    if (((i >> 2) & 3) == 1) mc = 0x7fffff - (uint32_t)(rnd() & 0xff);
    if (((i >> 2) & 3) == 2) mc = (uint32_t)(rnd() & 0xff);
    int ec = en - (int)(rnd() % (uint64_t)(span + 1));
    if (ec < 37) ec = 37; /* |c| >= 2^-90: the range the kernels send down this route */
    float c = u2f(((uint32_t)ec << 23) | mc | ((uint32_t)(rnd() & 1) << 31));
    if (fabsf(c) > nrm) c *= 0.5f;
    volatile float want = c / nrm;
    const float y = (float)(1.0 / (double)nrm);
    float q = c * y;
    float r = fmaf(-nrm, q, c);
    q = fmaf(r, y, q);
    r = fmaf(-nrm, q, c);
    q = fmaf(r, y, q);
    if (memcmp(&q, (const void*)&want, 4) != 0) ++bad;
  }
  return bad;
}
"""


def test_division_by_the_row_norm_through_the_fma_sequence_is_exact(tmp_path):
    """The prepare kernels' default route (csrc/kdi_internal.cuh, kdi_div_fma): y = float(1 / double(n)),
    q0 = c y, two residual corrections with FMAs.  Inside the operand range the kernels send down this
    route (2^-30 <= n <= 2^30, non-zero |c| >= 2^-90) it must give the IEEE float32 quotient bit for
    bit.  Checked with the host's own fmaf (a C helper built on the spot: NumPy has no fused
    multiply-add), adversarial significands (all ones, all zeros) included."""
    import ctypes
    import shutil
    import subprocess

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no C compiler")
    src = tmp_path / "fma_div.c"
    src.write_text(_FMA_DIV_C)
    lib = tmp_path / "fma_div.so"
    subprocess.run([gcc, "-O2", "-mfma", "-ffp-contract=off", "-shared", "-fPIC", str(src), "-o", str(lib), "-lm"],
                   check=True)
    f = ctypes.CDLL(str(lib)).kdi_fma_div_mismatches
    f.restype = ctypes.c_long
    f.argtypes = [ctypes.c_long, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    assert f(20_000_000, -30, 30, 60) == 0   # the whole admitted range
    assert f(20_000_000, -8, 8, 24) == 0     # where real rows live


def test_atan_polynomial_of_the_projection_kernel():
    """The odd polynomial the projection kernel uses for atan (kdi_project_dev.cuh, `kAtanC`), evaluated
    here in float64 exactly as the kernel evaluates it - Horner in r^2 with fused multiply-adds being at
    least as accurate as NumPy's separate operations - together with the reduction
    atan(a / b) = pi / 4 + atan((a - b) / (a + b)) for a > tan(pi / 8) b: within 3e-16 of NumPy's arctan
    over [0, 1], so the Lambert coordinates stay at float64 rounding level."""
    import re

    src = open(os.path.join(ROOT, "kikuchipy_b200", "csrc", "kdi_project_dev.cuh")).read()
    body = src[src.index("kAtanC[11] = {"):]
    body = body[: body.index("};")]
    coef = [float.fromhex(h) for h in re.findall(r"-?0x1\.[0-9a-f]+p[+-]\d+", body)]
    assert len(coef) == 11 and coef[0] == 1.0
    rng = np.random.default_rng(0)
    b = rng.uniform(0.1, 10.0, 400_000)
    a = b * np.concatenate([rng.uniform(0.0, 1.0, 399_990), [0.0, 1.0, np.tan(np.pi / 8), 0.41421356237309503, 0.5, 1e-300 / 0.1, 1e-9, 0.9999999999, 0.4142135623730951, 0.25]])
    red = a > 0.41421356237309504880 * b
    rn = np.where(red, a - b, a)
    rd = np.where(red, a + b, b)
    r = rn / rd
    u = r * r
    p = np.full_like(u, coef[10])
    for c in coef[9::-1]:
        p = p * u + c
    got = r * p + np.where(red, 0.78539816339744830962, 0.0)
    want = np.arctan(a / b)
    err = np.abs(got - want)
    assert np.all(err <= 3e-16 + 3e-16 * want), float(err.max())
    assert np.abs(np.abs(r).max()) <= 0.41421356237309515
