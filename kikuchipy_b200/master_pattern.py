"""Dictionary generation on the GPU: the ``EBSDMasterPattern.get_patterns`` step of the
dictionary-indexing workflow.

Mirrors /root/reference/src/kikuchipy/signals/ebsd_master_pattern.py:97-329 (``get_patterns``
with a fixed projection centre: argument handling, the rescale rule, ``scale = (npx - 1) / 2``,
lazy vs. computed result) and signals/util/_master_pattern.py:83-204 (direction cosines of the
detector pixels).  The projection itself is ``kdi_project_patterns`` /
``kdi_dictionary_indexing_projected`` in ``libkdi`` (csrc/kdi_project.cu).

In the reference a lazy dictionary is a Dask graph that ``_dictionary_indexing`` evaluates chunk
by chunk on the CPU inside the matching loop (indexing/_dictionary_indexing.py:106-108).  The
equivalent here is :class:`GeneratedDictionary`: it holds the master pattern on the device and the
rotations, and ``dictionary_indexing`` projects + normalises + matches it without the float32
patterns ever existing in host memory or crossing PCIe.
"""

from __future__ import annotations

import numpy as np

from . import _lib

# value ranges get_patterns rescales to when dtype_out differs from the master pattern's dtype
# (skimage.util.dtype.dtype_range, used at signals/ebsd_master_pattern.py:226)
_DTYPE_RANGE = {np.dtype(np.float32): (-1.0, 1.0), np.dtype(np.float64): (-1.0, 1.0)}


def direction_cosines(gnomonic_bounds, pcz, nrows, ncols, om_detector_to_sample, signal_mask=None):
    """Unit vectors from the source point to the detector pixels in the sample frame -
    ``_get_direction_cosines_for_fixed_pc`` (signals/util/_master_pattern.py:133-204).

    ``gnomonic_bounds`` = ``(x_min, x_max, y_min, y_max)`` and ``pcz`` of an ``EBSDDetector`` with
    one projection centre, ``om_detector_to_sample`` = ``(~detector.sample_to_detector).to_matrix()``.
    ``signal_mask``: 1-D, True = pixel kept (the polarity of THIS reference function).
    Host-side geometry (a few thousand pixels, once per detector); float64 ``(n, 3)``."""
    gb = np.asarray(gnomonic_bounds, dtype=np.float64).ravel()
    pcz = float(np.asarray(pcz).squeeze())
    x_scale = (gb[1] - gb[0]) / ncols
    y_scale = (gb[3] - gb[2]) / nrows
    det_gn_x = np.arange(gb[0], gb[1], x_scale)
    det_gn_y = np.arange(gb[3], gb[2], -y_scale)
    idx = np.arange(nrows * ncols)
    if signal_mask is not None:
        idx = idx[np.asarray(signal_mask, dtype=bool).ravel()]
    r = np.empty((idx.size, 3), dtype=np.float64)
    r[:, 0] = (det_gn_x[np.mod(idx, ncols)] + x_scale / 2) * pcz
    r[:, 1] = (det_gn_y[idx // ncols] - y_scale / 2) * pcz
    r[:, 2] = pcz
    r = r @ np.asarray(om_detector_to_sample, dtype=np.float64).squeeze().T
    return r / np.sqrt(np.sum(np.square(r), axis=-1))[:, None]


def _detector_direction_cosines(detector):
    """Direction cosines of a kikuchipy ``EBSDDetector`` (duck-typed) with one PC."""
    if int(np.prod(tuple(detector.navigation_shape))) != 1:
        raise ValueError("`detector.navigation_shape` is not (1,) or equal to `rotations.shape`")
    om = (detector.om_detector_to_sample if hasattr(detector, "om_detector_to_sample")
          else (~detector.sample_to_detector).to_matrix().squeeze())
    return direction_cosines(np.asarray(detector.gnomonic_bounds).squeeze(), detector.pcz, detector.nrows,
                             detector.ncols, om), (int(detector.nrows), int(detector.ncols))


class GeneratedDictionary:
    """A dictionary defined by rotations of a master pattern (the lazy result of
    ``get_patterns(compute=False)``).  ``shape`` / ``dtype`` / ``compute()`` behave like the
    lazy signal's data; ``rotations`` are the ``(N, 4)`` unit quaternions."""

    def __init__(self, context, handle, rotations, sig_shape):
        self.context = context
        self.master_pattern = handle
        self.rotations = np.ascontiguousarray(np.asarray(rotations, dtype=np.float64).reshape(-1, 4))
        self.sig_shape = tuple(int(s) for s in sig_shape)
        self.shape = (self.rotations.shape[0],) + self.sig_shape
        self.dtype = np.dtype(np.float32)
        self.ndim = len(self.shape)

    def __len__(self):
        return self.shape[0]

    def compute(self, **kwargs) -> np.ndarray:
        """Materialise the patterns on the host (``(N, sy, sx)`` float32)."""
        return self.context.project_patterns(self.master_pattern, self.rotations).reshape(self.shape)

    def __array__(self, dtype=None, copy=None):
        out = self.compute()
        return out if dtype is None else out.astype(dtype)


def _varying_pc_geometry(detector, pcs, om_detector_to_sample, detector_shape, n_rot):
    """``(pcs (n, 3), om (3, 3), (nrows, ncols))`` when one projection centre per rotation is
    asked for (``detector.navigation_shape == rotations.shape``, signals/ebsd_master_pattern.py
    :236-254), else ``None``."""
    if pcs is None and detector is not None and hasattr(detector, "pc"):
        pc = np.asarray(getattr(detector, "pc_flattened", detector.pc), dtype=np.float64).reshape(-1, 3)
        if pc.shape[0] > 1:
            pcs = pc
            detector_shape = tuple(int(s) for s in detector.shape)
            if om_detector_to_sample is None:
                om_detector_to_sample = (detector.om_detector_to_sample if hasattr(detector, "om_detector_to_sample")
                                         else (~detector.sample_to_detector).to_matrix().squeeze())
    if pcs is None:
        return None
    pcs = np.asarray(pcs, dtype=np.float64).reshape(-1, 3)
    if pcs.shape[0] != n_rot:
        raise ValueError("`detector.navigation_shape` is not (1,) or equal to `rotations.shape`")
    if om_detector_to_sample is None or detector_shape is None or len(detector_shape) != 2:
        raise ValueError("projection centres per pattern need the detector shape and the detector-to-sample matrix")
    return pcs, np.asarray(om_detector_to_sample, dtype=np.float64).reshape(3, 3), tuple(detector_shape)


def get_patterns(master_upper, master_lower, rotations, detector=None, *, direction_cosines=None,
                 detector_shape=None, dtype_out="float32", compute=False, context=None, pcs=None,
                 om_detector_to_sample=None):
    """Patterns projected onto a detector from a square-Lambert master pattern for the given
    rotations - ``EBSDMasterPattern.get_patterns`` (signals/ebsd_master_pattern.py:97-329).

    ``master_upper`` / ``master_lower``: the two hemispheres ``(npy, npx)`` (pass the upper one
    twice for a centrosymmetric phase, as the reference does).  ``rotations``: ``(N, 4)`` unit
    quaternions or an orix ``Rotation`` (``.data``).  The detector is given either as a kikuchipy
    ``EBSDDetector`` with one projection centre or as ``direction_cosines`` ``(S, 3)`` +
    ``detector_shape``.  ``dtype_out``: only float32 (the reference's default).  As in the
    reference, intensities are rescaled to the ``dtype_out`` range [-1, 1] per pattern when the
    master pattern's dtype differs from ``dtype_out``.

    Returns a :class:`GeneratedDictionary` (``compute=False``) or the ``(N, sy, sx)`` array.
    """
    if np.dtype(dtype_out) != np.float32:
        raise NotImplementedError("patterns are generated as float32 (get_patterns' default dtype_out)")
    rot = np.asarray(getattr(rotations, "data", rotations), dtype=np.float64)
    if rot.ndim > 3 or rot.shape[-1] != 4:
        raise ValueError("`rotations` must be an array of quaternions with at most two navigation dimensions")
    if rot.ndim == 3 and pcs is None and not (detector is not None and np.asarray(getattr(detector, "pc", [[0]])).reshape(-1, 3).shape[0] > 1):
        raise NotImplementedError("a dictionary has one navigation dimension; flatten the rotations first")
    varying = _varying_pc_geometry(detector, pcs, om_detector_to_sample, detector_shape, rot.reshape(-1, 4).shape[0])
    if varying is not None:
        # one projection centre per rotation (_project_patterns_from_master_pattern_with_varying_pc):
        # the patterns belong to map points, not to a dictionary, and are returned computed
        pc, om, shape = varying
        up = np.asarray(master_upper)
        ctx = context if context is not None else _lib.default_context()
        out_min, out_max = _DTYPE_RANGE[np.dtype(dtype_out)]
        handle = ctx.master_pattern(up, np.asarray(master_lower), np.zeros((shape[0] * shape[1], 3)),
                                    rescale=up.dtype != np.dtype(dtype_out), out_min=out_min, out_max=out_max)
        out = ctx.project_patterns_varying_pc(handle, rot.reshape(-1, 4), pc, shape[0], shape[1], om)
        return out.reshape((rot.reshape(-1, 4).shape[0],) + shape)
    if direction_cosines is None:
        if detector is None:
            raise ValueError("either a detector or direction cosines are needed")
        direction_cosines, detector_shape = _detector_direction_cosines(detector)
    direction_cosines = np.asarray(direction_cosines, dtype=np.float64).reshape(-1, 3)
    if detector_shape is None:
        detector_shape = (direction_cosines.shape[0],)
    if int(np.prod(detector_shape)) != direction_cosines.shape[0]:
        raise ValueError("detector shape and number of direction cosines differ")
    up = np.asarray(master_upper)
    rescale = up.dtype != np.dtype(dtype_out)  # ebsd_master_pattern.py:222-233
    out_min, out_max = _DTYPE_RANGE[np.dtype(dtype_out)]
    ctx = context if context is not None else _lib.default_context()
    handle = ctx.master_pattern(up, np.asarray(master_lower), direction_cosines, rescale=rescale,
                                out_min=out_min, out_max=out_max)
    gen = GeneratedDictionary(ctx, handle, rot.reshape(-1, 4), detector_shape)
    return gen.compute() if compute else gen
