"""CPU oracle for ``merge_crystal_maps``.  TEST INFRASTRUCTURE ONLY: imported by
``tests/`` (and nothing in ``kikuchipy_b200/``).

NumPy restatement of the array arithmetic of
/root/reference/src/kikuchipy/indexing/_merge_crystal_maps.py:28-354.  The reference function itself
needs orix ``CrystalMap`` / ``PhaseList`` objects (orix is not installed and not installable here), so
it cannot be executed in this container; the restatement works on the plain arrays those objects hold
and is **pinned** by the hard-coded expectations of the reference's own tests
(tests/test_indexing/test_merge_crystal_maps.py:289-370, :451-590), which ``tests/test_merge_maps.py``
replays through it.

Each map is given as a dict with

* ``scores``  ``(n_i,)`` or ``(n_i, N)`` - ``xmap.prop[scores_prop]``
* ``rotations`` ``(n_i, 4)`` or ``(n_i, N, 4)`` float64 - ``xmap.rotations.data``
* ``simulation_indices`` (optional) same leading shape as ``scores``
* ``phase_id`` ``(n_i,)`` - ``xmap.phase_id`` (only ``== -1`` is looked at)

and ``masks1d`` is the reference's ``navigation_masks1d`` (:154-165): per map ``None`` or a boolean
``(map_size,)`` array, ``True`` = the map holds that point.
"""

from __future__ import annotations

from math import copysign

import numpy as np


def sign_and_n_best(mean_n_best, greater_is_better):
    """_merge_crystal_maps.py:184-191."""
    if greater_is_better is None:
        return copysign(1, mean_n_best), abs(mean_n_best)
    return (1 if greater_is_better else -1), mean_n_best


def merge_arrays(maps, masks1d, map_size, mean_n_best=1, greater_is_better=None,
                 with_simulation_indices=False):
    n_maps = len(maps)
    sign, mean_n_best = sign_and_n_best(mean_n_best, greater_is_better)
    first = np.asarray(maps[0]["scores"])
    n_scores = first.shape[1] if first.ndim > 1 else 1

    # :199-214 combined (unsorted) scores, NaN where a map has no point
    comb_shape = (map_size,) + ((n_scores,) if n_scores > 1 else ()) + (n_maps,)
    scores_dtype = first.dtype
    combined = np.full(comb_shape, np.nan, dtype=np.dtype(f"f{scores_dtype.itemsize}"))
    for i, (mask, m) in enumerate(zip(masks1d, maps)):
        if mask is not None:
            combined[mask, ..., i] = m["scores"]
        else:
            combined[..., i] = m["scores"]

    # :216-225 best score per point and map; phase of the best score
    if n_scores > 1:
        best = combined[:, :mean_n_best].squeeze()
        if best.ndim > 2:
            best = np.nanmean(best, axis=1)
    else:
        best = combined
    phase_id = np.nanargmax(sign * best, axis=1)

    # :227-237 points not indexed in every map -> -1 (the fancy-indexed assignment of the masked
    # branch writes into a temporary copy, so maps with a mask never contribute a True)
    not_indexed = np.zeros((n_maps, map_size), dtype=bool)
    for i in range(n_maps):
        if masks1d[i] is None:
            not_indexed[i, np.asarray(maps[i]["phase_id"]) == -1] = True
    not_indexed = np.logical_and.reduce(not_indexed)
    phase_id[not_indexed] = -1

    # :239-296 per-point values of the winning map
    new_rot = np.zeros(comb_shape[:-1] + (4,), dtype="float")
    new_scores = np.zeros(comb_shape[:-1], dtype=scores_dtype)
    new_idx = np.zeros(comb_shape[:-1], dtype="int32") if with_simulation_indices else None
    for i, (mask, m) in enumerate(zip(masks1d, maps)):
        pm = phase_id == i
        if not pm.any():
            continue
        pm2 = pm[mask] if mask is not None else pm
        new_rot[pm] = np.asarray(m["rotations"])[pm2]
        new_scores[pm] = np.asarray(m["scores"])[pm2]
        if with_simulation_indices:
            new_idx[pm] = np.asarray(m["simulation_indices"])[pm2]

    # :298-308 stable merge sort of all scores of a point
    ms_shape = (comb_shape[0], int(np.prod(comb_shape[1:])))
    flat = combined.reshape(ms_shape)
    order = np.argsort(sign * -flat, kind="mergesort", axis=1)
    merged_scores = np.take_along_axis(flat, order, axis=-1)

    out = {"phase_id": phase_id, "scores": new_scores, "rotations": new_rot,
           "merged_scores": merged_scores}
    if with_simulation_indices:
        # :313-347 indices made unique across maps, then sorted like the scores
        lst = []
        for mask, m in zip(masks1d, maps):
            if mask is not None:
                s = np.full(comb_shape[:-1], np.nan)
                s[mask] = m["simulation_indices"]
            else:
                s = np.asarray(m["simulation_indices"])
            lst.append(s)
        comb_idx = np.dstack(lst)
        for i in range(1, comb_idx.shape[-1]):
            inc = abs(np.nanmax(comb_idx[..., i - 1]) - np.nanmin(comb_idx[..., i])) + 1
            comb_idx[..., i] += inc
        comb_idx = comb_idx.reshape(ms_shape)
        out["simulation_indices"] = new_idx
        out["merged_simulation_indices"] = np.take_along_axis(comb_idx, order, axis=-1)
    return out
