"""NumPy restatement of kikuchipy's dictionary-indexing hot path (the oracle).

TEST INFRASTRUCTURE ONLY - see ``oracle/__init__.py``.  Every function cites the
reference lines it follows (paths relative to ``/root/reference/src/kikuchipy``).
The arithmetic is deliberately the reference's own: cast -> mask -> normalise in
the working dtype, ``np.einsum("ik,mk->im", optimize=True)`` (BLAS), NumPy
partition/sort top-k, ``hstack``/``argsort``/``take_along_axis`` chunk merge.
"""

from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------------- #
# Metric preparation
# --------------------------------------------------------------------------- #


def mask_patterns(patterns: np.ndarray, signal_mask: np.ndarray | None) -> np.ndarray:
    """Column compaction, ``False`` = keep.

    indexing/similarity_metrics/_normalized_cross_correlation.py:185-188
    indexing/similarity_metrics/_normalized_dot_product.py:176-179
    """
    if signal_mask is None:
        return patterns
    return patterns[:, ~signal_mask.ravel()]


def zero_mean_normalize(patterns: np.ndarray) -> np.ndarray:
    """In-place centre and scale to unit L2 norm, per row.

    indexing/similarity_metrics/_normalized_cross_correlation.py:228-233
    """
    patterns_mean = np.mean(patterns, axis=1, keepdims=True)
    patterns -= patterns_mean
    patterns_norm = np.sqrt(np.sum(np.square(patterns), axis=1, keepdims=True))
    patterns /= patterns_norm
    return patterns


def normalize(patterns: np.ndarray) -> np.ndarray:
    """Scale to unit L2 norm per row, NO centring.

    indexing/similarity_metrics/_normalized_dot_product.py:181-194
    """
    patterns_norm = np.sqrt(np.sum(np.square(patterns), axis=1))[..., np.newaxis]
    return patterns / patterns_norm


def prepare_experimental(
    patterns: np.ndarray,
    metric: str,
    n_experimental_patterns: int,
    navigation_mask: np.ndarray | None = None,
    signal_mask: np.ndarray | None = None,
    dtype=np.float32,
) -> np.ndarray:
    """cast -> reshape -> drop navigation-masked rows -> compact columns ->
    normalise.

    NCC: indexing/similarity_metrics/_normalized_cross_correlation.py:113-128
    NDP: indexing/similarity_metrics/_normalized_dot_product.py:105-120
    """
    patterns = np.asarray(patterns).astype(dtype)
    patterns = patterns.reshape((n_experimental_patterns, -1))
    if navigation_mask is not None:
        patterns = patterns[~navigation_mask.ravel()]
    patterns = mask_patterns(patterns, signal_mask)
    if metric == "ncc":
        return zero_mean_normalize(patterns)
    elif metric == "ndp":
        return normalize(patterns)
    raise ValueError(metric)


def prepare_dictionary(
    patterns: np.ndarray,
    metric: str,
    signal_mask: np.ndarray | None = None,
    dtype=np.float32,
) -> np.ndarray:
    """cast -> compact columns -> normalise (input must already be 2-D).

    NCC: indexing/similarity_metrics/_normalized_cross_correlation.py:151-159
    NDP: indexing/similarity_metrics/_normalized_dot_product.py:141-150
    """
    patterns = patterns.astype(dtype)
    patterns = mask_patterns(patterns, signal_mask)
    if metric == "ncc":
        return zero_mean_normalize(patterns)
    elif metric == "ndp":
        return normalize(patterns)
    raise ValueError(metric)


def match(experimental: np.ndarray, dictionary: np.ndarray, dtype=np.float32) -> np.ndarray:
    """``(M, S) x (n, S) -> (M, n)`` similarity block.

    indexing/similarity_metrics/_normalized_cross_correlation.py:181-183
    indexing/similarity_metrics/_normalized_dot_product.py:172-174
    (``da.einsum`` on NumPy blocks is ``np.einsum`` per block.)
    """
    return np.einsum("ik,mk->im", experimental, dictionary, optimize=True, dtype=dtype)


# --------------------------------------------------------------------------- #
# dask ``Array.topk`` / ``Array.argtopk`` semantics on one NumPy block
# --------------------------------------------------------------------------- #


def topk(a: np.ndarray, k: int) -> np.ndarray:
    """k largest along the last axis, sorted descending.

    Called at indexing/_dictionary_indexing.py:198.  dask (dependency, not
    vendored; ``dask[array] >= 2021.8.1`` in pyproject.toml:44) implements it as
    ``np.partition(a, -k)[..., -k:]`` followed by an ascending ``np.sort`` that is
    then reversed.
    """
    a = np.partition(a, -k, axis=-1)[..., -k:]
    return np.sort(a, axis=-1)[..., ::-1]


def argtopk(a: np.ndarray, k: int) -> np.ndarray:
    """Positions of the k largest along the last axis, best first.

    Called at indexing/_dictionary_indexing.py:197.  dask: ``np.argpartition``
    of the block, gather, ascending ``np.argsort`` of the gathered values,
    reversed.
    """
    idx = np.argpartition(a, -k, axis=-1)[..., -k:]
    vals = np.take_along_axis(a, idx, axis=-1)
    order = np.argsort(vals, axis=-1)[..., ::-1]
    return np.take_along_axis(idx, order, axis=-1)


def match_chunk(experimental, simulated, keep_n, metric, signal_mask=None, dtype=np.float32):
    """indexing/_dictionary_indexing.py:172-203."""
    simulated = prepare_dictionary(simulated, metric, signal_mask, dtype)
    similarities = match(experimental, simulated, dtype)
    simulation_indices = argtopk(similarities, keep_n).reshape((-1, keep_n))
    scores = topk(similarities, keep_n).reshape((-1, keep_n))
    return simulation_indices, scores


# --------------------------------------------------------------------------- #
# Driver
# --------------------------------------------------------------------------- #


def dictionary_indexing(
    experimental: np.ndarray,
    dictionary: np.ndarray,
    metric: str = "ncc",
    keep_n: int = 20,
    n_per_iteration: int | None = None,
    navigation_mask: np.ndarray | None = None,
    signal_mask: np.ndarray | None = None,
    dtype=np.float32,
    n_experimental_patterns: int | None = None,
):
    """Restatement of ``_dictionary_indexing`` up to (indices, scores).

    indexing/_dictionary_indexing.py:66-71 (keep_n clip, prepare, reshape),
    :88-93 (single shot), :94-128 (chunk loop + merge).  ``metric.sign`` is +1
    for both built-in metrics (``_normalized_cross_correlation.py:62``,
    ``_normalized_dot_product.py:54``).

    Returns ``(simulation_indices, scores)`` of the matched (navigation-mask
    compacted) rows: shapes ``(M, keep_n)``; indices int64, scores ``dtype``.
    """
    sign = 1
    dictionary_size = dictionary.shape[0]
    if n_per_iteration is None:
        n_per_iteration = dictionary_size
    keep_n = min(keep_n, dictionary_size)
    n_iterations = int(np.ceil(dictionary_size / n_per_iteration))
    if n_experimental_patterns is None:
        # signals/ebsd.py:3076 - max(navigation_size, 1); the signal axes are the last two
        n_experimental_patterns = max(int(np.prod(experimental.shape[:-2])), 1)

    exp = prepare_experimental(
        experimental, metric, n_experimental_patterns, navigation_mask, signal_mask, dtype
    )
    dictionary = dictionary.reshape((dictionary_size, -1))
    n_experimental = exp.shape[0]

    if dictionary_size == n_per_iteration:
        simulation_indices, scores = match_chunk(
            exp, dictionary, keep_n, metric, signal_mask, dtype
        )
    else:
        negative_sign = -sign
        simulation_indices = np.zeros((n_experimental, keep_n), dtype=np.int32)
        scores = np.full((n_experimental, keep_n), negative_sign, dtype=dtype)
        chunk_starts = np.cumsum([0] + [n_per_iteration] * (n_iterations - 1))
        chunk_ends = np.cumsum([n_per_iteration] * n_iterations)
        chunk_ends[-1] = max(chunk_ends[-1], dictionary_size)
        for start, end in zip(chunk_starts, chunk_ends):
            chunk = dictionary[start:end]
            idx_i, sc_i = match_chunk(
                exp, chunk, min(keep_n, chunk.shape[0]), metric, signal_mask, dtype
            )
            idx_i = idx_i + start
            all_scores = np.hstack((scores, sc_i))
            all_idx = np.hstack((simulation_indices, idx_i))
            best = np.argsort(negative_sign * all_scores, axis=1)[:, :keep_n]
            scores = np.take_along_axis(all_scores, best, axis=1)
            simulation_indices = np.take_along_axis(all_idx, best, axis=1)
    return simulation_indices.astype(np.int64), scores


def assemble_result(
    simulation_indices: np.ndarray,
    scores: np.ndarray,
    nav_shape: tuple,
    navigation_mask: np.ndarray | None,
    keep_n: int,
):
    """Scatter matched rows back into full-map arrays.

    indexing/_dictionary_indexing.py:141-167.  With a navigation mask the
    arrays are scattered into ``np.empty`` buffers of the full map size
    (masked points hold unspecified values) and squeezed when ``keep_n == 1``;
    without one they are returned as is.  Returns
    ``(scores, simulation_indices, is_in_data)``.
    """
    n_all = int(np.prod(nav_shape)) if len(nav_shape) else 1
    if navigation_mask is not None:
        nav = ~navigation_mask.ravel()
        s_all = np.zeros((n_all, keep_n), dtype=scores.dtype)
        s_all[nav] = scores
        i_all = np.zeros((n_all, keep_n), dtype=simulation_indices.dtype)
        i_all[nav] = simulation_indices
        if keep_n == 1:
            s_all = s_all.squeeze()
            i_all = i_all.squeeze()
        return s_all, i_all, nav
    return scores, simulation_indices, np.ones(n_all, dtype=bool)


# --------------------------------------------------------------------------- #
# Orientation similarity map
# --------------------------------------------------------------------------- #


def orientation_similarity_map(
    simulation_indices: np.ndarray,
    map_shape: tuple,
    n_best: int | None = None,
    normalize: bool = False,
    from_n_best: int | None = None,
    footprint: np.ndarray | None = None,
    center_index: int = 2,
) -> np.ndarray:
    """Restatement of ``orientation_similarity_map`` without scipy.

    indexing/_orientation_similarity_map.py:96-128 (driver) and :131-152 (per
    pixel).  ``generic_filter(mode="constant", cval=-1)`` hands the callback the
    footprint's truthy positions in row-major order; out-of-map positions read
    -1.  The callback drops -1 and every position equal to the centre value,
    counts ``len(np.intersect1d(centre_list, neighbour_list))`` (set semantics)
    and takes ``np.nanmean`` (``nan`` when no neighbour is left).
    """
    sim = np.asarray(simulation_indices)
    nav_size, keep_n = sim.shape
    if n_best is None:
        n_best = keep_n
    elif n_best > keep_n:
        raise ValueError(f"n_best {n_best} cannot be greater than keep_n {keep_n}")
    if from_n_best is None:
        from_n_best = n_best
    if footprint is None:
        footprint = np.array([[0, 1, 0], [1, 1, 1], [0, 1, 0]])
    footprint = np.asarray(footprint).astype(bool)
    ny, nx = map_shape
    fy, fx = footprint.shape
    # scipy.ndimage centres the footprint at shape // 2
    oy, ox = fy // 2, fx // 2
    offsets = [(r - oy, c - ox) for r in range(fy) for c in range(fx) if footprint[r, c]]
    flat = np.arange(nav_size).reshape(map_shape)
    osm = np.zeros(tuple(map_shape) + (n_best - from_n_best + 1,), dtype=np.float32)
    for i, n in enumerate(range(n_best, from_n_best - 1, -1)):
        mi = sim[:, :n]
        for r in range(ny):
            for c in range(nx):
                v = []
                for dr, dc in offsets:
                    rr, cc = r + dr, c + dc
                    v.append(flat[rr, cc] if 0 <= rr < ny and 0 <= cc < nx else -1)
                v = np.array(v, dtype=int)
                centre = v[center_index]
                neighbours = v[(v != -1) & (v != centre)]
                counts = [len(np.intersect1d(mi[centre], m)) for m in mi[neighbours]]
                with np.errstate(invalid="ignore"):
                    import warnings

                    with warnings.catch_warnings():
                        warnings.simplefilter("ignore", category=RuntimeWarning)
                        val = np.nanmean(counts) if len(counts) else np.nan
                if normalize:
                    val /= n
                osm[r, c, i] = val
    return osm.squeeze()


# --------------------------------------------------------------------------- #
# Synthetic workloads - shared by tests and bench
# --------------------------------------------------------------------------- #


def circular_signal_mask(sig_shape: tuple) -> np.ndarray:
    """``~Window("circular", shape).astype(bool)``: True = excluded.

    filters/window.py:249-269 (make_circular), :388-415 (distance_to_origin);
    use in benchmarks/indexing/test_dictionary_indexing.py:52.
    """
    sy, sx = sig_shape
    oy, ox = sy // 2, sx // 2
    r, c = np.ogrid[:sy, :sx]
    dist = np.sqrt((r - oy) ** 2 + (c - ox) ** 2)
    return dist > max(oy, ox)


def synthetic_experimental(m: int, sig_shape: tuple, seed: int = 1) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, (m,) + tuple(sig_shape), dtype=np.uint8)


def synthetic_dictionary(n: int, sig_shape: tuple, seed: int = 2) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return rng.random((n,) + tuple(sig_shape), dtype=np.float32)


def planted_experimental(dictionary: np.ndarray, m: int, seed: int = 3):
    """Experimental patterns that are noisy copies of known dictionary rows, so
    the best match is known and score gaps are large (strict index parity)."""
    rng = np.random.default_rng(seed)
    n = dictionary.shape[0]
    j = rng.integers(0, n, m)
    noise = rng.random((m,) + dictionary.shape[1:], dtype=np.float32)
    exp = np.clip(np.rint(255.0 * (0.7 * dictionary[j] + 0.3 * noise)), 0, 255).astype(np.uint8)
    return exp, j


# --------------------------------------------------------------------------- #
# Parity helpers
# --------------------------------------------------------------------------- #


def compare_topk(idx_ref, sc_ref, idx_new, sc_new, tie_tol: float = 1e-6, score_tol: float = 1e-4):
    """Tie-tolerant comparison of two ranked lists (SURVEY.md section 8d parity gates).

    Returns a dict with: ``exact_rows`` (fraction of rows whose index lists are
    identical), ``tie_ok`` (True when every differing position involves
    reference scores closer than ``tie_tol`` to the score the other side holds
    at that position, and the index SETS differ only by such near-ties),
    ``max_dscore``.
    """
    idx_ref = np.asarray(idx_ref)
    idx_new = np.asarray(idx_new)
    sc_ref = np.asarray(sc_ref, dtype=np.float64)
    sc_new = np.asarray(sc_new, dtype=np.float64)
    same = idx_ref == idx_new
    exact_rows = float(np.mean(np.all(same, axis=1))) if idx_ref.size else 1.0
    max_dscore = float(np.max(np.abs(sc_ref - sc_new))) if sc_ref.size else 0.0
    diff = ~same
    # at a differing position both sides must hold (nearly) the same score
    tie_ok = bool(np.all(np.abs(sc_ref[diff] - sc_new[diff]) <= tie_tol)) if diff.any() else True
    return {
        "exact_rows": exact_rows,
        "tie_ok": tie_ok,
        "max_dscore": max_dscore,
        "n_diff_positions": int(diff.sum()),
        "scores_ok": max_dscore < score_tol,
    }
