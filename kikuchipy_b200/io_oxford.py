"""Oxford Instruments binary (.ebsp) reader (data formats on the
experimental side of the path): uncompressed patterns, any file version the reference reads.

Mirrors /root/reference/src/kikuchipy/io/plugins/oxford_binary/_api.py:45-585
(``OxfordBinaryFileReader``).  Layout: an int64 that is minus the version (a non-negative value means
version 0 and is already the first pattern position), one extra byte from version 4 on, then one int64
start position per map point (0 = pattern not stored), then per pattern a header (``map_x, map_y``
from version 5 on; ``is_compressed, nrows, ncols, n_bytes`` as int32), the pixels (uint8, or uint16
when ``n_bytes`` is twice the pixel count) and a footer (version 1: ``beam_x, beam_y`` float64;
version > 1: a flag byte before each of the two, present only when the flag is set; version 0:
none).  The number of map points is not stored and is guessed like the reference does, from the
first jump in the list of start positions.  Patterns come back in map order; if some are missing,
or no beam positions are stored, as a line of the present ones.
"""

from __future__ import annotations

import os
import struct

import numpy as np

from .io_nordif import NordifScan


def _header_dtype(version):
    fields = [("map_x", "<i4"), ("map_y", "<i4"), ("is_compressed", "<i4"), ("nrows", "<i4"), ("ncols", "<i4"),
              ("n_bytes", "<i4")]
    return np.dtype(fields[2:] if version < 5 else fields)


def _guess_number_of_patterns(f, size, starts_pos, header_size, version, min_pixels=1600):
    """``guess_number_of_patterns`` (:548-572): the start positions grow by one pattern record at a
    time; the first entry that is not a position (pattern bytes read as int64) breaks that."""
    f.seek(starts_pos)
    assumed = np.fromfile(f, dtype="<i8", count=size // (min_pixels + header_size))
    jump = np.abs(np.diff(assumed)) > 20 * (1024 * 1344 * 2 + header_size)
    n = int(np.nonzero(jump)[0][0])
    return n + 1 if version < 5 else n


def load_oxford_binary(filename, device=False, context=None):
    """Patterns ``(ny, nx, sy, sx)`` (or ``(n, sy, sx)``) of uint8 / uint16 with ``step_sizes`` and,
    in ``original_metadata``, ``map1d_id``, ``file_order`` and the stored beam / map positions."""
    size = os.path.getsize(filename)
    with open(filename, "rb") as f:
        first = struct.unpack("<q", f.read(8))[0]
        version = -first if first < 0 else 0
        starts_pos = 0 if version == 0 else (9 if version > 3 else 8)
        hdt = _header_dtype(version)
        n_points = _guess_number_of_patterns(f, size, starts_pos, hdt.itemsize, version)
        f.seek(starts_pos)
        starts = np.fromfile(f, dtype="<i8", count=n_points)
        first_pos = starts_pos + n_points * 8
        f.seek(first_pos)
        h = np.fromfile(f, dtype=hdt, count=1)[0]
        if h["is_compressed"]:
            raise NotImplementedError(f"Cannot read compressed EBSD patterns from {filename!r}")
        sy, sx, n_bytes = int(h["nrows"]), int(h["ncols"]), int(h["n_bytes"])
        pix = np.dtype("u1") if n_bytes == sy * sx else np.dtype("<u2")
        # footer layout, read off the first pattern (:226-248)
        f.seek(first_pos + hdt.itemsize + n_bytes)
        footer = []
        if version == 1:
            footer = [("beam_x", "<f8"), ("beam_y", "<f8")]
        elif version > 1:
            if struct.unpack("?", f.read(1))[0]:
                footer += [("has_beam_x", "?"), ("beam_x", "<f8")]
                f.seek(8, 1)
            else:
                footer += [("pad_x", "?")]
            if struct.unpack("?", f.read(1))[0]:
                footer += [("has_beam_y", "?"), ("beam_y", "<f8")]
            else:
                footer += [("pad_y", "?")]
        rec = np.dtype(hdt.descr + [("pattern", pix, (sy, sx))] + footer)
        present = starts != 0
        n_present = int(present.sum())
        f.seek(first_pos)
        records = np.fromfile(f, dtype=rec, count=n_present)
    order = ((starts - first_pos) / rec.itemsize).astype(int)
    names = rec.names
    has_beam = "beam_x" in names and "beam_y" in names
    data = records["pattern"]
    if present.all() and not np.allclose(np.diff(order), 1):
        data = data[order]
    if not present.all() or not has_beam:
        nav_shape, steps = (n_present,), (1.0,)
    else:  # :199-224 from the beam positions of the first, second and last pattern of the map
        ps = starts[present]

        def beam(pos):
            r = records[int((pos - first_pos) // rec.itemsize)]
            return float(r["beam_x"]), float(r["beam_y"])

        (x0, y0), (x1, _), (xl, yl) = beam(ps[0]), beam(ps[1]), beam(ps[-1])
        step = x1 - x0
        ny = int(abs(np.around(yl / step) - np.around(y0 / step))) + 1
        nx = int(abs(np.around(xl / step) - np.around(x0 / step))) + 1
        nav_shape, steps = (ny, nx), (step, step)
    data = np.ascontiguousarray(data).reshape(nav_shape + (sy, sx))
    sel = order[present]
    om = {"map1d_id": np.arange(n_points)[present], "file_order": sel}
    for key in ("beam_y", "beam_x", "map_x", "map_y"):
        if key in names:
            om[key] = records[key][sel]
    if device:
        from . import _lib

        ctx = context if context is not None else _lib.default_context()
        data = ctx.to_device(data)
    md = {"General": {"original_filename": filename, "title": os.path.splitext(os.path.basename(filename))[0]},
          "Signal": {"signal_type": "EBSD", "record_by": "image"}}
    scan = NordifScan(data, None, None, steps if len(steps) == 2 else (1.0, steps[0]), md, om)
    scan.version = version
    return scan
