"""CPU restatement (NumPy, float64) of the reference's dictionary GENERATION step: projecting a
square-Lambert master pattern onto the detector for a set of crystal rotations.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``): imported by ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU arm, never by the product path.

Reference (paths relative to /root/reference/src/kikuchipy):
  signals/util/_master_pattern.py
    :133-204  _get_direction_cosines_for_fixed_pc
    :299-370  _project_patterns_from_master_pattern_with_fixed_pc
    :449-527  _project_single_pattern_from_master_pattern
    :531-568  _vector2lambert
    :580-678  _get_lambert_interpolation_parameters
    :682-708  _get_pixel_from_master_pattern
  _utils/numba.py:62-81       rotate_vector
  pattern/_pattern.py:97-111  _rescale_with_min_max
  signals/ebsd_master_pattern.py:222-233,255-275  (rescale rule, scale = (npx - 1) / 2)

Pinned: ``tests/golden/make_golden_projection.py`` ran the reference's own Numba functions in
place (``oracle/ref_loader.load_master_pattern``) and stored inputs + outputs in
``tests/golden/projection.npz``; ``tests/test_oracle.py`` checks this restatement against them.
The reference compiles with ``fastmath=True``, so agreement is to float64 rounding (~1e-15
relative before the final cast), not bit-for-bit in float64; after the cast to float32 the values
agree to one float32 ulp.
"""

from __future__ import annotations

import numpy as np

SQRT_PI = np.sqrt(np.pi)
SQRT_PI_HALF = np.sqrt(np.pi / 2)
SQRT_PI_OVER_2 = SQRT_PI / 2
TWO_OVER_SQRT_PI = 2 / SQRT_PI


def direction_cosines_fixed_pc(gnomonic_bounds, pcz, nrows, ncols, om_detector_to_sample, signal_mask=None):
    """``_get_direction_cosines_for_fixed_pc`` (_master_pattern.py:133-204): unit vectors from
    the source point to the detector pixels (kept by ``signal_mask``, True = keep - note: the
    polarity of THIS mask in the reference is the opposite of the similarity metrics' masks),
    in the sample reference frame.  Shape ``(n pixels, 3)``, float64."""
    gb = np.asarray(gnomonic_bounds, dtype=np.float64)
    pcz = float(pcz)
    x_scale = (gb[1] - gb[0]) / ncols
    y_scale = (gb[3] - gb[2]) / nrows
    det_gn_x = np.arange(gb[0], gb[1], x_scale)
    det_gn_y = np.arange(gb[3], gb[2], -y_scale)
    idx_1d = np.arange(nrows * ncols)
    if signal_mask is not None:
        idx_1d = idx_1d[np.asarray(signal_mask, dtype=bool).ravel()]
    rows = idx_1d // ncols
    cols = np.mod(idx_1d, ncols)
    r_g = np.zeros((idx_1d.size, 3), dtype=np.float64)
    r_g[:, 0] = (det_gn_x[cols] + x_scale / 2) * pcz
    r_g[:, 1] = (det_gn_y[rows] - y_scale / 2) * pcz
    r_g[:, 2] = pcz
    r_g = np.dot(r_g, np.asarray(om_detector_to_sample, dtype=np.float64).T)
    return r_g / np.sqrt(np.sum(np.square(r_g), axis=-1))[:, None]


def rotate_vector(rotation, vector):
    """``rotate_vector`` (_utils/numba.py:62-81): rotate ``(n, 3)`` vectors by one unit
    quaternion ``(a, b, c, d)``."""
    a, b, c, d = (float(x) for x in rotation)
    x, y, z = vector[:, 0], vector[:, 1], vector[:, 2]
    aa, bb, cc, dd = a * a, b * b, c * c, d * d
    ac, ab, ad, bc, bd, cd = a * c, a * b, a * d, b * c, b * d, c * d
    out = np.zeros(vector.shape, dtype=np.float64)
    out[:, 0] = (aa + bb - cc - dd) * x + 2 * ((ac + bd) * z + (bc - ad) * y)
    out[:, 1] = (aa - bb + cc - dd) * y + 2 * ((ad + bc) * x + (cd - ab) * z)
    out[:, 2] = (aa - bb - cc + dd) * z + 2 * ((ab + cd) * y + (bd - ac) * x)
    return out


def vector2lambert(v):
    """``_vector2lambert`` (_master_pattern.py:531-568): square Lambert (X, Y) of ``(n, 3)``
    vectors."""
    w = v / np.sqrt(np.sum(np.square(v), axis=1))[:, None]
    x, y, z = w[:, 0], w[:, 1], w[:, 2]
    abs_z = np.abs(z)
    sqrt_z = np.sqrt(2 * (1 - abs_z))
    xy = np.zeros((v.shape[0], 2))
    pole = abs_z == 1
    first = (~pole) & (np.abs(y) <= np.abs(x))
    second = (~pole) & ~first
    with np.errstate(divide="ignore", invalid="ignore"):
        sx, sy = np.sign(x), np.sign(y)
        xy[first, 0] = (sx * sqrt_z * SQRT_PI_OVER_2)[first]
        xy[first, 1] = (sx * sqrt_z * TWO_OVER_SQRT_PI * np.arctan(y / x))[first]
        xy[second, 0] = (sy * sqrt_z * TWO_OVER_SQRT_PI * np.arctan(x / y))[second]
        xy[second, 1] = (sy * sqrt_z * SQRT_PI_OVER_2)[second]
    return xy


def lambert_interpolation_parameters(v, npx, npy, scale):
    """``_get_lambert_interpolation_parameters`` (_master_pattern.py:580-678)."""
    xy = scale * vector2lambert(v) / SQRT_PI_HALF
    i, j = xy[:, 1], xy[:, 0]
    nii = (i + scale).astype(np.int32)  # truncation towards zero, as ``np.int32(float)``
    nij = (j + scale).astype(np.int32)
    niip, nijp = nii + 1, nij + 1
    niip = np.where(niip >= npx, nii, niip)
    nijp = np.where(nijp >= npy, nij, nijp)
    nii = np.where(nii < 0, niip, nii)
    nij = np.where(nij < 0, nijp, nij)
    di = i - nii + scale
    dj = j - nij + scale
    return nii, nij, niip, nijp, di, dj, 1 - di, 1 - dj


def project_single_pattern(rotation, direction_cosines, master_upper, master_lower, npx, npy, scale, rescale,
                           out_min, out_max, dtype_out=np.float32):
    """``_project_single_pattern_from_master_pattern`` (_master_pattern.py:449-527)."""
    dc = rotate_vector(rotation, direction_cosines)
    nii, nij, niip, nijp, di, dj, dim, djm = lambert_interpolation_parameters(dc, npx, npy, scale)
    up = dc[:, 2] >= 0
    mu = np.asarray(master_upper)
    ml = np.asarray(master_lower)

    def pixel(mp):  # _get_pixel_from_master_pattern (:682-708)
        return (mp[nii, nij] * dim * djm + mp[niip, nij] * di * djm + mp[nii, nijp] * dim * dj
                + mp[niip, nijp] * di * dj)

    pattern = np.where(up, pixel(mu), pixel(ml)).astype(np.float64)
    if rescale:  # _rescale_with_min_max (pattern/_pattern.py:97-111)
        imin, imax = pattern.min(), pattern.max()
        pattern = (pattern - imin) / float(imax - imin) * (out_max - out_min) + out_min
    return pattern.astype(dtype_out)


def project_patterns(rotations, direction_cosines, master_upper, master_lower, npx=None, npy=None, scale=None,
                     rescale=False, out_min=1, out_max=2, dtype_out=np.float32):
    """``_project_patterns_from_master_pattern_with_fixed_pc`` (_master_pattern.py:299-370):
    ``(n rotations, n pixels)`` patterns of ``dtype_out``.  ``npx, npy`` default to the master
    pattern's signal shape and ``scale`` to ``(npx - 1) / 2`` as in
    ``EBSDMasterPattern.get_patterns`` (signals/ebsd_master_pattern.py:255-257)."""
    rotations = np.asarray(rotations, dtype=np.float64).reshape(-1, 4)
    if npx is None:
        npy, npx = np.asarray(master_upper).shape[0], np.asarray(master_upper).shape[1]
        # (get_patterns takes npx, npy = axes_manager.signal_shape, i.e. (columns, rows))
    if scale is None:
        scale = (npx - 1) / 2
    out = np.zeros((rotations.shape[0], direction_cosines.shape[0]), dtype=dtype_out)
    for r in range(rotations.shape[0]):
        out[r] = project_single_pattern(rotations[r], direction_cosines, master_upper, master_lower, int(npx),
                                        int(npy), float(scale), rescale, out_min, out_max, dtype_out)
    return out


def synthetic_master_pattern(n: int = 401, seed: int = 5, dtype=np.float32):
    """A smooth synthetic (upper, lower) master pattern pair of shape ``(n, n)``: a sum of a few
    random low-frequency cosines, rescaled to [0, 1] (float) or [0, 255] (uint8)."""
    rng = np.random.default_rng(seed)
    y, x = np.meshgrid(np.linspace(-1, 1, n), np.linspace(-1, 1, n), indexing="ij")
    out = []
    for _ in range(2):
        m = np.zeros((n, n))
        for _ in range(24):
            fx, fy = rng.uniform(-14, 14, 2)
            m += rng.uniform(0.3, 1.0) * np.cos(fx * x + fy * y + rng.uniform(0, 2 * np.pi))
        m = (m - m.min()) / (m.max() - m.min())
        out.append((m * 255).astype(np.uint8) if np.dtype(dtype) == np.uint8 else m.astype(dtype))
    return out[0], out[1]


def random_rotations(n: int, seed: int = 4) -> np.ndarray:
    """``n`` random unit quaternions, float64 (seed 4)."""
    rng = np.random.default_rng(seed)
    q = rng.normal(size=(n, 4))
    return q / np.linalg.norm(q, axis=1, keepdims=True)


def tilted_detector_matrix(tilt_deg: float = 70.0) -> np.ndarray:
    """A detector-to-sample orientation matrix for tests: rotation by ``tilt_deg`` about x
    (the real one comes from ``EBSDDetector.sample_to_detector``, out of scope here)."""
    t = np.deg2rad(tilt_deg)
    return np.array([[1, 0, 0], [0, np.cos(t), -np.sin(t)], [0, np.sin(t), np.cos(t)]], dtype=np.float64)


def gnomonic_bounds(nrows, ncols, pcx, pcy, pcz):
    """``EBSDDetector.gnomonic_bounds`` / ``get_gnomonic_bounds`` (detectors/_ebsd_detector.py:731-818,
    _utils/_gnonomic_bounds.py:23-62): ``(x_min, x_max, y_min, y_max)``."""
    aspect = ncols / nrows
    return np.array([-aspect * (pcx / pcz), aspect * (1 - pcx) / pcz, -(1 - pcy) / pcz, pcy / pcz])


def project_patterns_varying_pc(rotations, pcs, nrows, ncols, om_detector_to_sample, master_upper, master_lower,
                                rescale=False, out_min=1, out_max=2, dtype_out=np.float32):
    """``_get_direction_cosines_for_varying_pc`` + ``_project_patterns_from_master_pattern_with_
    varying_pc`` (_master_pattern.py:207-296, :374-445): rotation ``i`` seen from ``pcs[i]``."""
    rotations = np.asarray(rotations, dtype=np.float64).reshape(-1, 4)
    pcs = np.asarray(pcs, dtype=np.float64).reshape(-1, 3)
    npy, npx = np.asarray(master_upper).shape
    out = np.zeros((rotations.shape[0], nrows * ncols), dtype=dtype_out)
    for i, (r, pc) in enumerate(zip(rotations, pcs)):
        dc = direction_cosines_fixed_pc(gnomonic_bounds(nrows, ncols, *pc), pc[2], nrows, ncols, om_detector_to_sample)
        out[i] = project_single_pattern(r, dc, master_upper, master_lower, int(npx), int(npy), (npx - 1) / 2, rescale,
                                        out_min, out_max, dtype_out)
    return out
