"""CPU restatement (NumPy + SciPy) of the reference's experimental-side PREPROCESSING
(SURVEY.md section 8f.4): static and dynamic background removal and neighbour pattern averaging.
TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Reference (paths relative to /root/reference/src/kikuchipy):
  pattern/_pattern.py
    :96-111   _rescale_with_min_max
    :393-437  _remove_static_background_subtract / _divide
    :440-487  _remove_dynamic_background, :490-517 _remove_background_subtract / _divide
    :604-631  _dynamic_background_frequency_space_setup
  filters/fft_barnes.py:29-195  _fft_filter_setup, _pad_window, _pad_image, _fft_filter
  filters/window.py:117-180     Window (outer product of scipy.signal.windows.get_window)
  pattern/chunk.py:130-164      _average_neighbour_patterns, _rescale_neighbour_averaged_patterns
  signals/ebsd.py:442-557 (remove_static_background), :559-697 (remove_dynamic_background),
                 :943-1112 (average_neighbour_patterns)
Third party: scipy.ndimage.gaussian_filter / correlate and scipy.fft (scipy >= 1.7; 1.18.1 here) - the
oracle calls the same SciPy functions the reference calls.

Pinned: ``tests/golden/make_golden_preprocess.py`` ran the reference's own functions in place
(``oracle/ref_loader.load_preprocessing``) and stored inputs and outputs in
``tests/golden/preprocess.npz``.  The arithmetic of the Numba-compiled rescale was established
against them: with ``fastmath=True`` LLVM turns the division by the (loop-invariant) intensity range
into a multiplication by its reciprocal, and for float32 data everything stays in float32:
``out = (p - min) * (1 / (max - min)) * (omax - omin) + omin`` - which decides e.g. whether the
brightest pixel of a uint8 pattern becomes 255 or 254 (the result is truncated, not rounded).
"""

from __future__ import annotations

import numpy as np

DTYPE_RANGE = {np.dtype(np.uint8): (0, 255), np.dtype(np.uint16): (0, 65535), np.dtype(np.float32): (-1.0, 1.0),
               np.dtype(np.float64): (-1.0, 1.0)}


def rescale_f32(p, imin, imax, omin, omax):
    """``_rescale_with_min_max`` as compiled for float32 data (see the module docstring)."""
    f = np.float32
    rc = f(1) / (f(imax) - f(imin))
    t = (p.astype(f) - f(imin)) * rc
    # t * (omax - omin) + omin is contracted to one fused multiply-add (exact product, one rounding)
    return (t.astype(np.float64) * np.float64(f(omax - omin)) + np.float64(f(omin))).astype(f)


def _cast(x, dtype):
    dtype = np.dtype(dtype)
    if dtype.kind in "ui":
        with np.errstate(invalid="ignore"):
            return x.astype(dtype)  # truncation towards zero
    return x.astype(dtype)


def remove_static_background(patterns, static_bg, operation="subtract", scale_bg=False):
    """``EBSD.remove_static_background`` on an array ``(..., sy, sx)``; same dtype out."""
    pats = np.asarray(patterns)
    omin, omax = DTYPE_RANGE[pats.dtype]
    bg32 = np.asarray(static_bg).astype(np.float32)
    out = np.empty_like(pats)
    for idx in np.ndindex(pats.shape[:-2]):
        p = pats[idx].astype(np.float32)
        bg = bg32
        if scale_bg:
            bg = rescale_f32(bg32, bg32.min(), bg32.max(), p.min(), p.max())
        p = p - bg if operation == "subtract" else p / bg
        out[idx] = _cast(rescale_f32(p, p.min(), p.max(), omin, omax), pats.dtype)
    return out


def gaussian_window(std, truncate):
    """The window ``_dynamic_background_frequency_space_setup`` builds (:612-615)."""
    from scipy.signal.windows import get_window

    n = int(truncate * std)
    g = get_window(("gaussian", std), Nx=n, fftbins=False)
    w = np.outer(g, g)
    w = w / (2 * np.pi * std**2)
    return w / np.sum(w)


def fft_filter(image, window):
    """``_fft_filter`` with its set-up (filters/fft_barnes.py:29-195): linear convolution with the
    window, the image continued by its edge values."""
    from scipy.fft import irfft2, next_fast_len, rfft2

    iy, ix = image.shape
    wy, wx = window.shape
    fy, fx = next_fast_len(iy + wy - 1, real=True), next_fast_len(ix + wx - 1, real=True)
    wp = np.zeros((fy, fx), dtype=np.float32)
    wp[:wy, :wx] = np.flipud(np.fliplr(window))
    tf = rfft2(wp)
    oy, ox = wy - ((wy - 1) // 2) - 1, wx - ((wx - 1) // 2) - 1
    ay, ax = (wy - 1) // 2, (wx - 1) // 2
    pad = np.zeros((fy, fx), dtype=np.float32)
    pad[:iy, :ix] = image
    pad[iy:iy + ay, :ix] = image[-1, :]
    pad[:iy, ix:ix + ax] = image[:, -1:]
    pad[fy - oy:, :ix] = image[0, :]
    pad[:iy, fx - ox:] = image[:, :1]
    pad[iy:iy + ay, ix:ix + ax] = image[-1, -1]
    pad[fy - oy:, ix:ix + ax] = image[0, -1]
    pad[iy:iy + ay, fx - ox:] = image[-1, 0]
    pad[fy - oy:, fx - ox:] = image[0, 0]
    res = irfft2(rfft2(pad) * tf, (fy, fx))
    return np.real(res[ay:ay + iy, ax:ax + ix])


def dynamic_background(pattern_f32, filter_domain="frequency", std=None, truncate=4.0):
    """The blurred pattern both domains subtract or divide by (float32)."""
    from scipy.ndimage import gaussian_filter

    if std is None:
        std = pattern_f32.shape[1] / 8  # signals/ebsd.py:644-645: signal_shape[0] = columns
    if filter_domain == "frequency":
        return fft_filter(pattern_f32, gaussian_window(std, truncate)).astype(np.float32)
    if filter_domain == "spatial":
        return gaussian_filter(pattern_f32, sigma=std, truncate=truncate)
    raise ValueError(f"{filter_domain} must be either of ['frequency', 'spatial']")


def remove_dynamic_background(patterns, operation="subtract", filter_domain="frequency", std=None, truncate=4.0):
    """``EBSD.remove_dynamic_background`` on an array ``(..., sy, sx)``; same dtype out."""
    pats = np.asarray(patterns)
    omin, omax = DTYPE_RANGE[pats.dtype]
    out = np.empty_like(pats)
    for idx in np.ndindex(pats.shape[:-2]):
        p = pats[idx].astype(np.float32)
        bg = dynamic_background(p, filter_domain, std, truncate)
        p = p - bg if operation == "subtract" else p / bg
        out[idx] = _cast(rescale_f32(p, p.min(), p.max(), omin, omax), pats.dtype)
    return out


def circular_window(shape):
    """``Window("circular", shape)``: ones with the corners outside the inscribed ellipse removed
    (filters/window.py ``make_circular``: distance to the centre > half the shape)."""
    shape = tuple(shape)
    w = np.ones(shape)
    if len(shape) == 1:
        return w
    ny, nx = shape
    y, x = np.indices(shape)
    origin = ny // 2, nx // 2
    dist = np.sqrt((y - origin[0]) ** 2 + (x - origin[1]) ** 2)
    w[dist > max(origin)] = 0  # see tests: pinned against the reference's Window in the goldens
    return w


def average_neighbour_patterns(patterns, window):
    """``EBSD.average_neighbour_patterns`` on ``(ny, nx, sy, sx)`` (or ``(n, sy, sx)``) with a
    window array over the navigation axes; same dtype out."""
    from scipy.ndimage import correlate

    pats = np.asarray(patterns)
    nav_shape = pats.shape[:-2]
    w = np.asarray(window)
    if len(nav_shape) > w.ndim:
        w = w.reshape(w.shape + (1,))
    sums = correlate(np.ones(nav_shape, dtype=int), weights=w, mode="constant")
    w4 = w.reshape(w.shape + (1, 1))
    corr = correlate(pats.astype(np.float32), weights=w4, mode="constant")
    omin, omax = DTYPE_RANGE[pats.dtype]
    out = np.zeros(pats.shape, dtype=pats.dtype)
    for idx in np.ndindex(nav_shape):
        p = corr[idx] / np.float32(sums[idx])
        out[idx] = _cast(rescale_f32(p, p.min(), p.max(), omin, omax), pats.dtype)
    return out
