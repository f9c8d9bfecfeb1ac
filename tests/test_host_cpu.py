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


def _gloo_worker(rank, world, port, tmp):
    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        exp = orc.synthetic_experimental(24, (12, 12), seed=1)
        dic = orc.synthetic_dictionary(301, (12, 12), seed=2)
        start, end = kb.shard_bounds(301, world, rank)
        # per-shard stage: the oracle stands in for the GPU call (checker), global indices via the offset
        idx, sc = orc.dictionary_indexing(exp, dic[start:end], keep_n=7)
        idx = idx + start
        s_all, i_all = kb.gather_topk(torch.from_numpy(sc.copy()), torch.from_numpy(idx.copy()))
        assert tuple(s_all.shape) == (world, 24, 7)
        # list-major layout: list r must be rank r's result
        assert np.array_equal(s_all[rank].numpy(), sc) and np.array_equal(i_all[rank].numpy(), idx)
        alls = np.concatenate(list(s_all.numpy()), axis=1)
        alli = np.concatenate(list(i_all.numpy()), axis=1)
        best = np.argsort(-alls, axis=1, kind="stable")[:, :7]
        np.savez(os.path.join(tmp, f"r{rank}.npz"), idx=np.take_along_axis(alli, best, 1),
                 sc=np.take_along_axis(alls, best, 1))
    finally:
        dist.destroy_process_group()


def test_sharded_gather_matches_unsharded_gloo(tmp_path):
    import torch.multiprocessing as mp

    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_gloo_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    exp = orc.synthetic_experimental(24, (12, 12), seed=1)
    dic = orc.synthetic_dictionary(301, (12, 12), seed=2)
    ridx, rsc = orc.dictionary_indexing(exp, dic, keep_n=7)
    for r in range(world):
        z = np.load(os.path.join(str(tmp_path), f"r{r}.npz"))
        c = orc.compare_topk(ridx, rsc, z["idx"], z["sc"])
        assert c["tie_ok"] and c["max_dscore"] < 1e-6
        assert c["exact_rows"] == 1.0


def test_bench_cpu_arm_runs_small(monkeypatch):
    sys.path.insert(0, ROOT)
    import bench

    monkeypatch.setattr(bench, "N_DICT", 3000)
    exp, dic = bench.host_inputs(8)
    v, detail = bench.cpu_sample(exp, dic, 100)
    assert v > 0 and detail["sample_patterns"] == 8 and detail["dictionary"] == 3000
