"""Experimental-side preprocessing on the GPU: what a kikuchipy workflow
runs on the measured patterns before dictionary indexing.

Mirrors ``EBSD.remove_static_background`` (/root/reference/src/kikuchipy/signals/ebsd.py:442-557),
``EBSD.remove_dynamic_background`` (:559-697) and ``EBSD.average_neighbour_patterns`` (:943-1112):
same keywords and ``ValueError`` texts; the per-pattern work is ``kdi_preprocess_patterns`` /
``kdi_average_neighbour_patterns`` in ``libkdi`` (csrc/kdi_preprocess.cu).  ``preprocess`` runs the
static and the dynamic correction in one launch.  With ``device_output=True`` the result is a CUDA
tensor that ``dictionary_indexing`` accepts as it is: raw detector bytes cross PCIe once.

Patterns are arrays ``(..., sy, sx)`` of uint8, uint16 or float32 (NumPy, or CUDA torch tensors), or
kikuchipy-like signals (``.data``; with ``inplace=True``, the default of the reference, the signal's
``data`` is replaced and ``None`` returned).
"""

from __future__ import annotations

import warnings

import numpy as np

from . import _lib

_OPS = {"subtract": 1, "divide": 2}


def _data_of(signal):
    return signal.data if hasattr(signal, "axes_manager") and hasattr(signal, "data") else signal


def _finish(signal, out, shape, inplace):
    out = out.reshape(shape)
    if hasattr(signal, "axes_manager") and hasattr(signal, "data"):
        if inplace:
            signal.data = out
            return None
        import copy

        new = copy.copy(signal)
        new.data = out
        return new
    return out


def gaussian_kernel1d(sigma, truncate=4.0):
    """The normalised kernel ``scipy.ndimage.gaussian_filter`` correlates with (SciPy's
    ``_gaussian_kernel1d``, order 0): radius ``int(truncate * sigma + 0.5)``."""
    sigma = float(sigma)
    radius = int(truncate * sigma + 0.5)
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (sigma * sigma) * x**2)
    return phi / phi.sum()


def gaussian_window1d(n, std):
    """``scipy.signal.windows.gaussian(n, std, sym=True)``: one factor of the reference's 2-D
    window (``filters/window.py:166-173``), normalised to unit sum."""
    k = np.arange(0, n) - (n - 1.0) / 2.0
    w = np.exp(-(k**2) / (2 * float(std) * float(std)))
    return w / w.sum()


def _dynamic_weights(filter_domain, std, truncate, sig_shape):
    if std is None:
        std = sig_shape[1] / 8  # signals/ebsd.py:644-645: axes_manager.signal_shape[0] = columns
    if filter_domain == "frequency":
        n = int(truncate * std)  # pattern/_pattern.py:612
        if n < 1:
            raise ValueError(f"All window axes {(n, n)} must be > 0.")
        w = gaussian_window1d(n, std)
        return 0, w, w
    if filter_domain == "spatial":
        w = gaussian_kernel1d(std, truncate)
        return 1, w, w
    raise ValueError(f"{filter_domain} must be either of ['frequency', 'spatial']")


def _check_static(data, static_bg):
    if not isinstance(static_bg, np.ndarray):
        raise ValueError("`EBSD.static_background` is not a valid array")
    if np.dtype(str(data.dtype).replace("torch.", "")) != static_bg.dtype:
        raise ValueError(f"Static background dtype_out {static_bg.dtype} is not the same as pattern dtype_out "
                         f"{data.dtype}")
    if tuple(static_bg.shape) != tuple(data.shape[-2:]):
        raise ValueError(f"Signal {tuple(data.shape[-2:])} and static background {static_bg.shape} shapes are not "
                         "the same")
    return static_bg.astype(np.float32)


def preprocess(signal, static_bg=None, static_operation="subtract", scale_bg=False, dynamic=True,
               dynamic_operation="subtract", filter_domain="frequency", std=None, truncate=4.0, inplace=False,
               device_output=False, context=None):
    """Static (when ``static_bg`` is given) then dynamic (when ``dynamic``) background removal in
    one launch: equal to ``remove_static_background`` followed by ``remove_dynamic_background``."""
    data = _data_of(signal)
    shape = tuple(data.shape)
    if len(shape) < 2:
        raise ValueError("patterns need the two detector axes")
    sy, sx = shape[-2:]
    kw = {}
    if static_bg is not None:
        if static_operation not in _OPS:
            raise ValueError(f"{static_operation} must be either of {list(_OPS)}")
        kw.update(static_op=_OPS[static_operation], static_bg=_check_static(data, static_bg), scale_bg=scale_bg)
    if dynamic:
        if dynamic_operation not in _OPS:
            raise ValueError(f"{dynamic_operation} must be either of {list(_OPS)}")
        dom, wy, wx = _dynamic_weights(filter_domain, std, truncate, (sy, sx))
        kw.update(dynamic_op=_OPS[dynamic_operation], dynamic_domain=dom, weights_y=wy, weights_x=wx)
    ctx = context if context is not None else _lib.default_context()
    out = ctx.preprocess_patterns(data, sy, sx, device_output=device_output, **kw)
    return _finish(signal, out, shape, inplace)


def remove_static_background(signal, operation="subtract", static_bg=None, scale_bg=False, show_progressbar=None,
                             inplace=True, lazy_output=None, *, device_output=False, context=None):
    """``EBSD.remove_static_background`` (``signals/ebsd.py:442-557``)."""
    if lazy_output and inplace:
        raise ValueError("'lazy_output=True' requires 'inplace=False'")
    if static_bg is None:
        static_bg = getattr(signal, "static_background", None)
        if not isinstance(static_bg, np.ndarray):
            raise ValueError("`EBSD.static_background` is not a valid array")
    return preprocess(signal, static_bg=static_bg, static_operation=operation,
                      scale_bg=scale_bg, dynamic=False, inplace=inplace and hasattr(signal, "axes_manager"),
                      device_output=device_output, context=context)


def remove_dynamic_background(signal, operation="subtract", filter_domain="frequency", std=None, truncate=4.0,
                              show_progressbar=None, inplace=True, lazy_output=None, *, device_output=False,
                              context=None, **kwargs):
    """``EBSD.remove_dynamic_background`` (``signals/ebsd.py:559-697``)."""
    if lazy_output and inplace:
        raise ValueError("'lazy_output=True' requires 'inplace=False'")
    return preprocess(signal, dynamic=True, dynamic_operation=operation, filter_domain=filter_domain, std=std,
                      truncate=truncate, inplace=inplace and hasattr(signal, "axes_manager"),
                      device_output=device_output, context=context)


def averaging_window(window="circular", shape=(3, 3), **kwargs):
    """The window ``Window(window, shape, **kwargs)`` (``filters/window.py:117-180``) for the names
    this path uses: ``"circular"`` (ones without the corners farther from the centre than half the
    larger axis), ``"rectangular"`` and ``"gaussian"`` (``std=``), or an array passed through."""
    if isinstance(window, np.ndarray):
        return np.asarray(window, dtype=np.float64)
    shape = tuple(int(s) for s in shape)
    if any(s < 1 for s in shape):
        raise ValueError(f"All window axes {shape} must be > 0.")
    if window in ("circular", "rectangular", "boxcar"):
        w = np.ones(shape)
        if window == "circular" and len(shape) == 2:
            origin = tuple(s // 2 for s in shape)
            y, x = np.indices(shape)
            w[np.sqrt((y - origin[0]) ** 2 + (x - origin[1]) ** 2) > max(origin)] = 0
        return w
    if window == "gaussian":
        std = kwargs["std"]

        def g(n):
            k = np.arange(0, n) - (n - 1.0) / 2.0
            return np.exp(-(k**2) / (2 * float(std) * float(std)))

        return g(shape[0]) if len(shape) == 1 else np.outer(g(shape[0]), g(shape[1]))
    raise NotImplementedError(f"window {window!r}: pass the window as an array")


def window_sums(nav_shape, window):
    """``correlate(np.ones(nav_shape, dtype=int), weights=window, mode="constant")``
    (``signals/ebsd.py:1029-1033``): the window summed over the neighbours inside the map, in C
    order, truncated to integers like SciPy's integer output."""
    ny, nx = nav_shape
    wy, wx = window.shape
    # loop over the window taps only (C order, like SciPy's accumulation): tap (a, b) contributes to
    # the map points whose neighbour (y + a - wy // 2, x + b - wx // 2) lies inside the map
    acc = np.zeros((ny, nx), dtype=np.float64)
    for a in range(wy):
        dy = a - wy // 2
        y0, y1 = max(0, -dy), min(ny, ny - dy)
        if y0 >= y1:
            continue
        for b in range(wx):
            dx = b - wx // 2
            x0, x1 = max(0, -dx), min(nx, nx - dx)
            if x0 < x1:
                acc[y0:y1, x0:x1] += window[a, b]
    return acc.astype(np.int32)  # truncation, like SciPy's integer output


def average_neighbour_patterns(signal, window="circular", window_shape=(3, 3), show_progressbar=None, inplace=True,
                               lazy_output=None, *, device_output=False, context=None, **kwargs):
    """``EBSD.average_neighbour_patterns`` (``signals/ebsd.py:943-1112``)."""
    if lazy_output and inplace:
        raise ValueError("'lazy_output=True' requires 'inplace=False'")
    data = _data_of(signal)
    shape = tuple(data.shape)
    nav_shape = shape[:-2]
    if len(nav_shape) not in (1, 2):
        raise ValueError("patterns must have one or two navigation axes")
    w = averaging_window(window, window_shape, **kwargs)
    if w.shape in [(1,), (1, 1)]:
        warnings.warn(f"A window of shape {w.shape} was passed, no averaging is therefore performed")
        return None
    if w.ndim == 1:  # a 1-D window acts along the first navigation axis (window.reshape(shape + (1,)), :1024-1025)
        w = w.reshape(-1, 1)
    nav2 = nav_shape if len(nav_shape) == 2 else (nav_shape[0], 1)
    sums = window_sums(nav2, w)
    ctx = context if context is not None else _lib.default_context()
    out = ctx.average_neighbour_patterns(data, nav2[0], nav2[1], shape[-2] * shape[-1], w, sums,
                                         device_output=device_output)
    return _finish(signal, out, shape, inplace and hasattr(signal, "axes_manager"))
