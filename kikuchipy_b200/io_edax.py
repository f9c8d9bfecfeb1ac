"""EDAX binary (.up1 / .up2) reader: raw detector patterns straight to the GPU (SURVEY.md section
8f.4, data formats on the experimental side of the path).

Mirrors /root/reference/src/kikuchipy/io/plugins/edax_binary/_api.py:25-175 (``EDAXBinaryFileReader``):
a uint32 version (1 or >= 3), pattern width, height and the byte offset of the first pattern; from
version 3 on also the map width and height, a hexagonal-grid flag and the two step sizes.  ``.up1``
holds uint8, ``.up2`` uint16 patterns.  A hexagonal grid is returned as a line of patterns with a
warning, as in the reference.  The result is a :class:`~kikuchipy_b200.io_nordif.NordifScan`-like
container (``data``, ``step_sizes``, ``metadata``); with ``device=True`` ``data`` is a CUDA tensor.
"""

from __future__ import annotations

import os
import warnings

import numpy as np

from .io_nordif import NordifScan


def read_header(filename):
    """``EDAXBinaryFileReader.__init__`` + ``read_header``."""
    ext = os.path.splitext(filename)[1][1:].lower()
    if ext not in ("up1", "up2"):
        raise ValueError(f"{filename!r} is not an EDAX .up1 / .up2 file")
    dtype = np.dtype(np.uint8 if ext == "up1" else np.uint16)
    size = os.path.getsize(filename)
    with open(filename, "rb") as f:
        version = int(np.fromfile(f, "uint32", 1)[0])
        if version == 2:
            raise ValueError("Only files with version 1 or >= 3, not 2, can be read")
        sx, sy, offset = (int(v) for v in np.fromfile(f, "uint32", 3))
        in_file = int((size - offset) / (sx * sy * dtype.itemsize))
        if version == 1:
            nx, ny, n, dx, dy, is_hex = in_file, 1, in_file, 1, 1, False
        else:
            nx, ny = (int(v) for v in np.fromfile(f, "uint32", 2, offset=1))
            is_hex = bool(np.fromfile(f, "uint8", 1)[0])
            if is_hex:
                warnings.warn("Returned signal has one navigation dimension since an hexagonal grid is not supported")
                nx, ny, n = in_file, 1, in_file
            else:
                n = nx * ny
            dx, dy = (float(v) for v in np.fromfile(f, "float64", 2))
    return {"sx": sx, "sy": sy, "pattern_offset": offset, "nx": nx, "ny": ny, "n_patterns": n, "dx": dx, "dy": dy,
            "is_hex": is_hex, "dtype": dtype, "version": version}


def load_edax_binary(filename, nav_shape=None, device=False, context=None):
    """``read_scan``: ``data`` ``(ny, nx, sy, sx)`` (``(nx, sy, sx)`` for one row) of uint8 / uint16."""
    h = read_header(filename)
    if nav_shape is not None and not h["is_hex"]:
        ny, nx = nav_shape
        if int(ny * nx) != h["n_patterns"]:
            raise ValueError(f"Given `nav_shape` {nav_shape} does not match the number of patterns in the file, "
                             f"{h['n_patterns']}.")
    else:
        ny, nx = h["ny"], h["nx"]
    shape = (ny, nx, h["sy"], h["sx"]) if ny != 1 else (nx, h["sy"], h["sx"])
    count = int(np.prod(shape))
    with open(filename, "rb") as f:
        f.seek(h["pattern_offset"])
        if device:
            from . import _lib

            ctx = context if context is not None else _lib.default_context()
            data = ctx.to_device(np.fromfile(f, dtype=h["dtype"], count=count)).reshape(shape)
        else:
            data = np.fromfile(f, dtype=h["dtype"], count=count).reshape(shape)
    md = {"General": {"original_filename": filename, "title": os.path.splitext(os.path.basename(filename))[0]},
          "Signal": {"signal_type": "EBSD", "record_by": "image"}}
    return NordifScan(data, None, None, (h["dy"], h["dx"]), md, {"edax_header": {k: v for k, v in h.items() if k != "dtype"}})
