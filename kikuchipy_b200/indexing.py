"""``EBSD.dictionary_indexing()`` / ``orientation_similarity_map()`` on the GPU.

Mirrors, for the dictionary-indexing path only,
/root/reference/src/kikuchipy/signals/ebsd.py:1827-1984 (argument handling and error messages),
``indexing/_dictionary_indexing.py:36-237`` (driver, printed messages, result assembly) and
``indexing/_orientation_similarity_map.py:30-128``.  The matching itself is one call into
``libkdi`` (``kdi_dictionary_indexing``): prepare once, stream the dictionary, fused
tensor-core match + top-k, exact rescoring.
"""

from __future__ import annotations

import time

import numpy as np

from . import _lib
from .master_pattern import GeneratedDictionary
from .similarity_metrics import (
    NormalizedCrossCorrelationMetric,
    NormalizedDotProductMetric,
    SimilarityMetric,
    _GpuMetric,
)

_METRICS = {"ncc": NormalizedCrossCorrelationMetric, "ndp": NormalizedDotProductMetric}


class DictionaryIndexingResult:
    """What callers read from the ``CrystalMap`` the reference returns
    (``_dictionary_indexing.py:141-167``), for use without orix.

    Attributes: ``scores`` and ``simulation_indices`` (``(n points, keep_n)``; 1-D when
    ``keep_n == 1`` and a navigation mask was given), ``rotations`` (quaternions gathered from
    the dictionary, or ``None``), ``shape`` (navigation shape), ``is_in_data``, ``size`` (indexed
    points), ``rotations_per_point``, ``scan_unit``, ``x`` / ``y`` coordinates and ``prop``.
    """

    def __init__(self, scores, simulation_indices, rotations, nav_shape, step_sizes, is_in_data,
                 keep_n, phase_name, scan_unit="px"):
        self.prop = {"scores": scores, "simulation_indices": simulation_indices}
        self.rotations = rotations
        self.shape = tuple(int(s) for s in nav_shape)
        self.is_in_data = is_in_data
        self.rotations_per_point = keep_n
        self.phase_name = phase_name
        self.scan_unit = scan_unit
        # coordinates: row-major flattening, x varies fastest (orix create_coordinate_arrays)
        shape2 = self.shape if len(self.shape) else (1,)
        steps = tuple(step_sizes) if step_sizes is not None else (1.0,) * len(shape2)
        steps = steps if len(steps) == len(shape2) else (1.0,) * len(shape2)
        if len(shape2) == 1:
            self.x = np.arange(shape2[0]) * steps[0]
            self.y = None
        else:
            ny, nx = shape2[-2], shape2[-1]
            self.x = np.tile(np.arange(nx) * steps[-1], ny)
            self.y = np.repeat(np.arange(ny) * steps[-2], nx)

    @property
    def scores(self):
        s = self.prop["scores"]
        return s[self.is_in_data] if s.shape[0] == self.is_in_data.size else s

    @property
    def simulation_indices(self):
        s = self.prop["simulation_indices"]
        return s[self.is_in_data] if s.shape[0] == self.is_in_data.size else s

    @property
    def size(self):
        return int(self.is_in_data.sum())

    def __repr__(self):
        return (f"DictionaryIndexingResult(shape={self.shape}, size={self.size}, "
                f"rotations_per_point={self.rotations_per_point}, phase={self.phase_name!r})")


def _info_message(metric, n_experimental_all, dictionary_size, phase_name, n_experimental=None):
    """``_dictionary_indexing_info_message`` (``_dictionary_indexing.py:206-237``)."""
    info = f"Dictionary indexing information:\n  Phase name: {phase_name}\n"
    if n_experimental is not None and n_experimental != n_experimental_all:
        info += f"  Matching {n_experimental}/{n_experimental_all} experimental pattern(s)"
    else:
        info += f"  Matching {n_experimental_all} experimental pattern(s)"
    info += f" to {dictionary_size} dictionary pattern(s)\n  {metric}"
    return info


def _unwrap(signal):
    """(data, navigation shape, signal shape, step sizes, scan unit, xmap) of a kikuchipy-like
    signal (``.data`` / ``.axes_manager``) or a plain array whose last two axes are the detector."""
    if hasattr(signal, "axes_manager") and hasattr(signal, "data"):
        am = signal.axes_manager
        nav_shape = tuple(am.navigation_shape[::-1])
        sig_shape = tuple(am.signal_shape[::-1])
        steps = tuple(a.scale for a in am.navigation_axes[::-1])
        unit = "px"
        try:
            units = [str(a.units) for a in am.navigation_axes]
            if units and units[0] not in ("<undefined>", "None", ""):
                unit = units[0]
        except Exception:  # noqa: BLE001
            pass
        return signal.data, nav_shape, sig_shape, steps, unit, getattr(signal, "xmap", None)
    data = signal
    if hasattr(signal, "step_sizes") and hasattr(signal, "data") and not hasattr(signal, "ndim"):
        # a scan container of this package's readers (io_nordif.NordifScan): patterns + step sizes in um
        data = signal.data
        if len(data.shape) < 2:
            raise ValueError("pattern arrays need at least the two detector axes")
        return data, tuple(data.shape[:-2]), tuple(data.shape[-2:]), tuple(signal.step_sizes), "um", None
    if isinstance(data, GeneratedDictionary):
        return data, (data.shape[0],), data.sig_shape if len(data.sig_shape) == 2 else (1,) + data.sig_shape, None, "px", None
    if len(data.shape) < 2:
        raise ValueError("pattern arrays need at least the two detector axes")
    return data, tuple(data.shape[:-2]), tuple(data.shape[-2:]), None, "px", None


def _prepare_metric(metric, navigation_mask, signal_mask, dtype, rechunk, n_exp, n_dict, context):
    """``EBSD._prepare_metric`` (``signals/ebsd.py:3049-3088``)."""
    if isinstance(metric, str) and metric in _METRICS:
        metric = _METRICS[metric](context=context)
        metric.rechunk = rechunk
    if not isinstance(metric, SimilarityMetric):
        raise ValueError(
            f"'{metric}' must be either of {_METRICS.keys()} or a custom metric class inheriting "
            "from SimilarityMetric. See kikuchipy.indexing.SimilarityMetric"
        )
    metric.n_experimental_patterns = max(n_exp, 1)
    metric.n_dictionary_patterns = max(n_dict, 1)
    if navigation_mask is not None:
        metric.navigation_mask = navigation_mask
    if signal_mask is not None:
        metric.signal_mask = signal_mask
    if dtype is not None:
        metric.dtype = dtype
    metric.raise_error_if_invalid()
    return metric


def dictionary_indexing(
    experimental,
    dictionary,
    metric="ncc",
    keep_n: int = 20,
    n_per_iteration: int | None = None,
    navigation_mask: np.ndarray | None = None,
    signal_mask: np.ndarray | None = None,
    rechunk: bool = False,
    dtype=None,
    *,
    dictionary_rotations=None,
    phase_name: str = "",
    context=None,
    index_offset: int = 0,
    verbose: bool = True,
):
    """Match each experimental pattern to a dictionary of simulated patterns and keep the
    ``keep_n`` best matches - ``EBSD.dictionary_indexing`` (``signals/ebsd.py:1827-1984``).

    ``experimental``: kikuchipy ``EBSD`` signal, or array ``(..navigation.., sy, sx)`` (NumPy,
    or a CUDA torch tensor for device-resident input).  ``dictionary``: ``EBSD`` signal with a
    1-D ``xmap``, or array ``(N, sy, sx)`` (+ ``dictionary_rotations``, an ``(N, 4)`` quaternion
    array, when rotations are wanted in the result).  The remaining parameters are the
    reference's.  Returns an orix ``CrystalMap`` when orix and a dictionary ``xmap`` are
    available, else a :class:`DictionaryIndexingResult`.
    """
    exp_data, nav_shape, sig_shape_exp, steps, scan_unit, _ = _unwrap(experimental)
    dict_data, dict_nav, sig_shape_dict, _, _, dict_xmap = _unwrap(dictionary)
    is_signal = hasattr(dictionary, "axes_manager")
    dict_size = int(np.prod(dict_nav)) if len(dict_nav) else 1

    if n_per_iteration is None:
        chunks = getattr(dict_data, "chunksize", None)  # lazy (Dask) dictionary
        n_per_iteration = chunks[0] if chunks else dict_size

    if navigation_mask is not None:
        if navigation_mask.shape != nav_shape:
            raise ValueError(
                f"The navigation mask shape {navigation_mask.shape} and the signal's navigation "
                f"shape {nav_shape} must be identical"
            )
        elif navigation_mask.all():
            raise ValueError(
                "The navigation mask must allow for indexing of at least one pattern (at least "
                "one value equal to `False`)"
            )
        elif not isinstance(navigation_mask, np.ndarray):
            raise ValueError("The navigation mask must be a NumPy array")
    if signal_mask is not None:
        if not isinstance(signal_mask, np.ndarray):
            raise ValueError("The signal mask must be a NumPy array")
    if sig_shape_exp != sig_shape_dict:
        raise ValueError(
            f"Experimental {sig_shape_exp} and dictionary {sig_shape_dict} signal shapes must be "
            "identical"
        )
    if len(dict_nav) != 1 or (is_signal and (dict_xmap is None or dict_xmap.shape != (dict_size,))):
        raise ValueError(
            "Dictionary signal must have a non-empty `EBSD.xmap` attribute of equal size as the "
            "number of dictionary patterns, and both the signal and crystal map must have only "
            "one navigation dimension"
        )

    n_exp_all = int(np.prod(nav_shape)) if len(nav_shape) else 1
    metric = _prepare_metric(
        metric, navigation_mask, signal_mask, dtype, rechunk, n_exp_all, dict_size, context
    )
    generated = isinstance(dict_data, GeneratedDictionary)
    # a lazy (Dask-like) dictionary is NOT materialised: it is computed chunk by chunk inside the
    # loop, like the reference does (_dictionary_indexing.py:105-108)
    lazy_dictionary = hasattr(dict_data, "compute") and not generated
    if hasattr(exp_data, "compute"):
        exp_data = exp_data.compute()

    keep_n = min(int(keep_n), dict_size)  # _dictionary_indexing.py:67
    n_exp = n_exp_all if navigation_mask is None else int((~navigation_mask).sum())
    if not phase_name and dict_xmap is not None:
        try:
            phase_name = dict_xmap.phases.names[0]
        except Exception:  # noqa: BLE001
            phase_name = ""
    if verbose:
        print(_info_message(metric, n_exp_all, dict_size, phase_name, n_exp))

    t0 = time.time()
    if generated and dictionary_rotations is None:
        dictionary_rotations = dict_data.rotations
    f64_mode = isinstance(metric, _GpuMetric) and np.dtype(metric.dtype) == np.float64
    if f64_mode:
        # dtype=float64 (reference: the metric casts to float64 and everything downstream is float64):
        # the reference's chunk loop over the metric's hooks, whose match() ranks float32-nominated
        # candidates by float64 scores computed on the device (similarity_metrics.SimilarityBlock64)
        if generated:
            dict_data = dict_data.compute()
        simulation_indices, scores = _generic_driver(exp_data, dict_data, metric, keep_n, n_per_iteration)
        simulation_indices = simulation_indices.astype(np.int64)
        simulation_indices = simulation_indices + index_offset if index_offset else simulation_indices
    elif isinstance(metric, _GpuMetric) and generated:
        # dictionary generated on the device from rotations of a master pattern: the reference's
        # `dictionary_chunk.compute()` (_dictionary_indexing.py:106-108) fused with the prepare step
        ctx = metric.context
        if dict_data.context is not ctx:
            raise ValueError("the generated dictionary lives on a different device context than the metric")
        ctx.set_signal_mask(metric.signal_mask)
        simulation_indices, scores = ctx.dictionary_indexing_projected(
            exp_data, n_exp_all, dict_data.master_pattern, dict_data.rotations, metric._kdi_metric, keep_n,
            nav_mask=metric.navigation_mask, index_offset=index_offset,
        )
    elif isinstance(metric, _GpuMetric) and lazy_dictionary:
        # the reference's chunk loop with the device working behind it: chunk i is uploaded, prepared
        # and matched while the host computes chunk i + 1 (kdi_job_begin / _append / _finish)
        ctx = metric.context
        ctx.set_signal_mask(metric.signal_mask)
        n_per_iteration = max(1, int(n_per_iteration))
        job = ctx.indexing_job(exp_data, n_exp_all, dict_size, metric._kdi_metric, keep_n,
                               nav_mask=metric.navigation_mask, index_offset=index_offset)
        try:
            for start in range(0, dict_size, n_per_iteration):
                chunk = dict_data[start:start + n_per_iteration]
                if hasattr(chunk, "compute"):
                    chunk = chunk.compute()
                chunk = np.asarray(chunk)
                job.append(chunk.reshape((chunk.shape[0], -1)))
            simulation_indices, scores = job.finish()
        finally:
            job.abort()  # no-op after finish()
    elif isinstance(metric, _GpuMetric):
        ctx = metric.context
        ctx.set_signal_mask(metric.signal_mask)
        simulation_indices, scores = ctx.dictionary_indexing(
            exp_data, n_exp_all, dict_data, dict_size, metric._kdi_metric, keep_n,
            n_per_iteration=n_per_iteration, nav_mask=metric.navigation_mask, index_offset=index_offset,
        )
    else:
        # custom SimilarityMetric subclass: the reference's generic driver over its three hooks
        # (_dictionary_indexing.py:70, 193-201); selection is whatever match() returns
        if generated:
            dict_data = dict_data.compute()
        simulation_indices, scores = _generic_driver(exp_data, dict_data, metric, keep_n, n_per_iteration)
        simulation_indices = simulation_indices + index_offset if index_offset else simulation_indices
    total_time = max(time.time() - t0, 1e-12)
    if verbose:
        print(
            f"  Indexing speed: {n_exp / total_time:.5f} patterns/s, "
            f"{n_exp * dict_size / total_time:.5f} comparisons/s"
        )

    rotations = None
    rot_src = dictionary_rotations
    if rot_src is None and dict_xmap is not None and hasattr(dict_xmap, "rotations"):
        rot_src = dict_xmap.rotations
    # result assembly (_dictionary_indexing.py:141-167)
    if navigation_mask is not None:
        nav = ~navigation_mask.ravel()
        scores_all = np.zeros((n_exp_all, keep_n), dtype=scores.dtype)
        scores_all[nav] = scores
        idx_all = np.zeros((n_exp_all, keep_n), dtype=simulation_indices.dtype)
        idx_all[nav] = simulation_indices
        if rot_src is not None and not _is_orix(rot_src):
            rot = np.zeros((n_exp_all, keep_n, 4))
            rot[..., 0] = 1.0  # identity elsewhere
            rot[nav] = np.take(np.asarray(rot_src), simulation_indices - index_offset, axis=0)
            rotations = rot
        if keep_n == 1:
            scores_all = scores_all.squeeze()
            idx_all = idx_all.squeeze()
            if rotations is not None:
                rotations = rotations.reshape(n_exp_all, 4)
        out_scores, out_idx, is_in_data = scores_all, idx_all, nav
    else:
        out_scores, out_idx, is_in_data = scores, simulation_indices, np.ones(n_exp_all, dtype=bool)
        if rot_src is not None and not _is_orix(rot_src):
            rotations = np.take(np.asarray(rot_src), simulation_indices - index_offset, axis=0)

    if _is_orix(rot_src):
        return _to_crystal_map(out_scores, out_idx, simulation_indices, rot_src, dict_xmap, nav_shape,
                               steps, navigation_mask, keep_n, scan_unit, index_offset)
    return DictionaryIndexingResult(out_scores, out_idx, rotations, nav_shape, steps, is_in_data,
                                    keep_n, phase_name, scan_unit)


def _generic_driver(exp_data, dict_data, metric, keep_n, n_per_iteration):
    """The reference's driver over the three hooks of a custom ``SimilarityMetric``
    (``_dictionary_indexing.py:66-71, 88-128, 193-201``): experimental rows prepared once, the
    dictionary taken in chunks of ``n_per_iteration`` rows (computed on the spot when lazy), per
    chunk ``keep_n`` clamped to the chunk length and the chunk start added to the indices, running
    merge by ``argsort(-sign * scores)`` - so lower-is-better metrics work like in the reference."""
    dict_size = metric.n_dictionary_patterns
    keep_n = min(int(keep_n), dict_size)
    experimental = metric.prepare_experimental(exp_data)
    lazy = hasattr(dict_data, "compute")
    on_device = hasattr(dict_data, "is_cuda")  # a CUDA tensor stays where it is; chunks are views of it
    dictionary = dict_data if (lazy or on_device) else np.asarray(dict_data)
    dictionary = dictionary.reshape((dict_size, -1))

    def match_chunk(simulated, k):
        sim = metric.match(experimental, metric.prepare_dictionary(simulated))
        idx, sc = sim.argtopk(k, axis=-1), sim.topk(k, axis=-1)
        idx = idx.compute() if hasattr(idx, "compute") else idx
        sc = sc.compute() if hasattr(sc, "compute") else sc
        return np.asarray(idx).reshape((-1, k)), np.asarray(sc).reshape((-1, k))

    n_per_iteration = max(1, int(n_per_iteration))
    if dict_size == n_per_iteration and not lazy:
        return match_chunk(dictionary, keep_n)
    negative_sign = -metric.sign
    n_exp = int(experimental.shape[0])
    indices = np.zeros((n_exp, keep_n), dtype=np.int32)
    scores = np.full((n_exp, keep_n), negative_sign, dtype=metric.dtype)
    for start in range(0, dict_size, n_per_iteration):
        end = min(start + n_per_iteration, dict_size)
        chunk = dictionary[start:end]
        if hasattr(chunk, "compute"):
            chunk = chunk.compute()
        idx_i, sc_i = match_chunk(chunk, min(keep_n, end - start))
        all_scores = np.hstack((scores, sc_i))
        all_indices = np.hstack((indices, idx_i + start))
        best = np.argsort(negative_sign * all_scores, axis=1)[:, :keep_n]
        scores = np.take_along_axis(all_scores, best, axis=1)
        indices = np.take_along_axis(all_indices, best, axis=1)
    return indices, scores


def _is_orix(rot) -> bool:
    return rot is not None and type(rot).__module__.startswith("orix")


def _to_crystal_map(scores, idx, idx_matched, rotations, dict_xmap, nav_shape, steps, navigation_mask,
                    keep_n, scan_unit, index_offset):
    """Result assembly as an orix ``CrystalMap`` (``_dictionary_indexing.py:141-167``)."""
    from orix.crystal_map import CrystalMap, create_coordinate_arrays
    from orix.quaternion import Rotation

    xmap_kw, _ = create_coordinate_arrays(nav_shape, steps)
    if navigation_mask is not None:
        nav = ~navigation_mask.ravel()
        xmap_kw["is_in_data"] = nav
        rot = Rotation.identity((nav.size, keep_n))
        rot[nav] = rotations[idx_matched - index_offset].data
        if keep_n == 1:
            rot = rot.flatten()
        xmap_kw["rotations"] = rot
    else:
        xmap_kw["rotations"] = rotations[idx_matched - index_offset]
    xmap_kw["prop"] = {"scores": scores, "simulation_indices": idx}
    xmap = CrystalMap(phase_list=dict_xmap.phases_in_data, **xmap_kw)
    xmap.scan_unit = scan_unit
    return xmap


def orientation_similarity_map(
    xmap,
    n_best: int | None = None,
    simulation_indices_prop: str = "simulation_indices",
    normalize: bool = False,
    from_n_best: int | None = None,
    footprint: np.ndarray | None = None,
    center_index: int = 2,
    *,
    context=None,
) -> np.ndarray:
    """Orientation similarity map (OSM) - ``_orientation_similarity_map.py:30-128``.

    ``xmap``: anything with ``.prop[simulation_indices_prop]`` of shape ``(n points, keep_n)`` and
    a 2-D ``.shape`` (an orix ``CrystalMap`` or a :class:`DictionaryIndexingResult`).
    """
    simulation_indices = np.asarray(xmap.prop[simulation_indices_prop])
    nav_size, keep_n = simulation_indices.shape
    if n_best is None:
        n_best = keep_n
    elif n_best > keep_n:
        raise ValueError(f"n_best {n_best} cannot be greater than keep_n {keep_n}")
    if from_n_best is None:
        from_n_best = n_best
    if footprint is None:
        footprint = np.array([[0, 1, 0], [1, 1, 1], [0, 1, 0]])
    footprint = np.asarray(footprint)
    if len(xmap.shape) != 2 or footprint.ndim != 2:
        raise RuntimeError("filter footprint array has incorrect shape.")  # scipy's error for non-2-D maps
    ny, nx = (int(s) for s in xmap.shape)
    if ny * nx != nav_size:
        raise ValueError("map shape and number of indexed points differ")
    ctx = context if context is not None else _lib.default_context()
    osm = ctx.orientation_similarity_map(
        simulation_indices, ny, nx, int(n_best), int(from_n_best), normalize, footprint, int(center_index)
    )
    return osm.squeeze()
