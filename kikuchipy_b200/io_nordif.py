"""NORDIF reader: raw detector bytes straight to the GPU (data format on
the experimental side of the path).

Mirrors the non-lazy branch of /root/reference/src/kikuchipy/io/plugins/nordif/_api.py:36-157
(``file_reader``) and ``_get_settings_from_file`` (:160-233): ``Pattern.dat`` is ``ny * nx`` frames of
``sy * sx`` uint8 values, ``Setting.txt`` (latin-1, tab separated) carries the scan and pattern sizes,
the step size and the detector angles, ``Background acquisition pattern.bmp`` the static background.
No HyperSpy signal is built (out of scope): the result is a small container whose ``data`` every
function of this package accepts; with ``device=True`` it is a CUDA tensor uploaded once through
pinned memory, so preprocessing and indexing run without the patterns returning to the host.
"""

from __future__ import annotations

import os
import re
import struct
import warnings

import numpy as np

from .refinement import Detector


class NordifScan:
    """``data`` ``(ny, nx, sy, sx)`` uint8 (squeezed like the reference's), ``static_background``
    ``(sy, sx)`` uint8 or ``None``, ``detector`` (:class:`~kikuchipy_b200.refinement.Detector`
    with the file's tilts, or ``None`` without a setting file), ``step_sizes`` ``(dy, dx)`` in um,
    ``metadata`` (beam energy, magnification, microscope, working distance)."""

    def __init__(self, data, static_background, detector, step_sizes, metadata, original_metadata):
        self.data = data
        self.static_background = static_background
        self.detector = detector
        self.step_sizes = step_sizes
        self.metadata = metadata
        self.original_metadata = original_metadata

    @property
    def shape(self):
        return tuple(self.data.shape)


def read_bmp_gray8(path):
    """An uncompressed 8-bit BMP with a grey palette (what the NORDIF software writes) as a uint8
    array; rows are stored bottom-up and padded to four bytes."""
    with open(path, "rb") as f:
        raw = f.read()
    if raw[:2] != b"BM":
        raise ValueError(f"{path!r} is not a BMP file")
    offset = struct.unpack_from("<I", raw, 10)[0]
    width, height = struct.unpack_from("<ii", raw, 18)
    planes, bpp, compression = struct.unpack_from("<HHI", raw, 26)
    if bpp != 8 or compression != 0:
        raise NotImplementedError("only uncompressed 8-bit BMP files are read")
    n_colors = struct.unpack_from("<I", raw, 46)[0] or 256
    header = struct.unpack_from("<I", raw, 14)[0]
    palette = np.frombuffer(raw, dtype=np.uint8, count=4 * n_colors, offset=14 + header).reshape(n_colors, 4)
    stride = (width + 3) & ~3
    rows = abs(height)
    px = np.frombuffer(raw, dtype=np.uint8, count=stride * rows, offset=offset).reshape(rows, stride)[:, :width]
    if height > 0:
        px = px[::-1]
    return np.ascontiguousarray(palette[:, 0][px])  # grey palette: blue = green = red


def _get(content, block, key, pattern):
    """Value of ``key`` inside ``[block]`` matched with ``pattern`` (one capture group)."""
    start = next((i for i, line in enumerate(content) if block in line), None)
    if start is None:
        return None
    for line in content[start + 1:]:
        if line.startswith("["):
            break
        if line.startswith(key):
            m = re.search(pattern, line)
            return m.group(1) if m else None
    return None


def read_settings(filename, pattern_type="acquisition"):
    """``_get_settings_from_file``: ``(metadata, header lines, scan sizes, detector keywords)``."""
    with open(filename, encoding="latin-1") as f:
        content = f.read().splitlines()
    num = _get(content, "[Area]", "Number of samples", r"Number of samples\t(.*)\t#")
    res = _get(content, f"[{pattern_type.capitalize()} settings]", "Resolution", r"Resolution\t(.*)\tpx")
    step = _get(content, "[Area]", "Step size", r"Step size\t(.*)\t")
    if num is None or res is None or step is None:
        raise ValueError(f"Could not read the scan size, pattern size or step size from {filename!r}")
    ny, nx = (int(i) for i in num.split("x"))
    sx, sy = (int(i) for i in res.split("x"))
    sizes = {"ny": ny, "nx": nx, "sy": sy, "sx": sx, "step_y": float(step), "step_x": float(step)}

    def num_or(block, key, pattern, default=0.0):
        v = _get(content, block, key, pattern)
        try:
            return float(v)
        except (TypeError, ValueError):
            warnings.warn(f"Failed to read {key!r} in settings file {filename!r}")
            return default

    tilt = -num_or("[Detector angles]", "Elevation", r"Elevation\t(.*)\t")
    detector = {"shape": (sy, sx), "sample_tilt": num_or("[Microscope]", "Tilt angle", r"Tilt angle\t(.*)\t"),
                "tilt": 0.0 if np.isclose(tilt, 0) else tilt,
                "azimuthal": num_or("[Detector angles]", "Azimuthal", r"Azimuthal\t(.*)\t")}
    md = {"Acquisition_instrument": {"SEM": {
        "beam_energy": num_or("[Microscope]", "Accelerating voltage", r"Accelerating voltage\t(.*)\tkV"),
        "magnification": int(num_or("[Microscope]", "Magnification", r"Magnification\t(.*)\t#")),
        "microscope": f"{_get(content, '[Microscope]', 'Manufacturer', 'Manufacturer' + chr(9) + '(.*)' + chr(9))} "
                      f"{_get(content, '[Microscope]', 'Model', 'Model' + chr(9) + '(.*)' + chr(9))}",
        "working_distance": num_or("[Microscope]", "Working distance", r"Working distance\t(.*)\tmm"),
    }}}
    return md, content, sizes, detector


def load_nordif(filename, scan_size=None, pattern_size=None, setting_file=None, device=False, context=None):
    """Read a NORDIF ``Pattern.dat`` (``file_reader``, non-lazy).  ``scan_size`` ``(nx, ny)`` (or an
    int for a line scan) and ``pattern_size`` ``(sx, sy)`` override / replace the setting file."""
    folder = os.path.dirname(os.path.abspath(filename))
    if setting_file is None:
        setting_file = os.path.join(folder, "Setting.txt")
    md, omd, detector, sizes = {}, {}, None, None
    if os.path.isfile(setting_file):
        md, header, sizes, det_kw = read_settings(setting_file)
        omd = {"nordif_header": header}
        detector = Detector(**det_kw)
        if not scan_size:
            scan_size = (sizes["nx"], sizes["ny"])
        if not pattern_size:
            pattern_size = (sizes["sx"], sizes["sy"])
    elif scan_size is None or pattern_size is None:
        raise ValueError("No setting file found and no scan_size or pattern_size detected in input arguments. "
                         "These must be set if no setting file is provided")
    bg_file = os.path.join(folder, "Background acquisition pattern.bmp")
    try:
        static_bg = read_bmp_gray8(bg_file)
    except FileNotFoundError:
        static_bg = None
        warnings.warn(f"Could not read static background pattern {bg_file!r}, however it can be set as "
                      "'EBSD.static_background'")
    nx, ny = (scan_size, 1) if isinstance(scan_size, int) else scan_size
    sx, sy = pattern_size
    count = ny * nx * sy * sx
    step = (sizes["step_y"], sizes["step_x"]) if sizes else (1.0, 1.0)
    if sizes is None:
        warnings.warn("Could not calibrate scan dimensions, this can be done using set_scan_calibration()")
    if device:
        from . import _lib

        ctx = context if context is not None else _lib.default_context()
        got = np.fromfile(filename, dtype=np.uint8, count=count)
        if got.size < count:
            warnings.warn("Pattern size and scan size larger than file size! Will attempt to load by zero padding "
                          "incomplete frames.")
            got = np.pad(got, [(0, count - got.size)])
        # (same squeeze as the host branch and the reference, io/plugins/nordif/_api.py)
        data = ctx.to_device(got).reshape(ny, nx, sy, sx).squeeze()
    else:
        data = np.fromfile(filename, dtype=np.uint8, count=count)
        if data.size < count:
            warnings.warn("Pattern size and scan size larger than file size! Will attempt to load by zero padding "
                          "incomplete frames.")
            data = np.pad(data, [(0, count - data.size)]).reshape(ny, nx, sy, sx)
        else:
            data = data.reshape(ny, nx, sy, sx).squeeze()
    return NordifScan(data, static_bg, detector, step, md, omd)


def load(filename, device=False, context=None, **kwargs):
    """Read a pattern file by its extension, like ``kikuchipy.load`` does for the binary formats
    this package reads: ``.dat`` (NORDIF), ``.up1`` / ``.up2`` (EDAX), ``.ebsp`` (Oxford Instruments),
    ``.h5`` / ``.hdf5`` / ``.h5ebsd`` (kikuchipy h5ebsd)."""
    ext = os.path.splitext(filename)[1].lower()
    if not os.path.isfile(filename):
        raise IOError(f"No filename matches {filename!r}")
    if ext == ".dat":
        return load_nordif(filename, device=device, context=context, **kwargs)
    if ext in (".up1", ".up2"):
        from .io_edax import load_edax_binary

        return load_edax_binary(filename, device=device, context=context, **kwargs)
    if ext == ".ebsp":
        from .io_oxford import load_oxford_binary

        return load_oxford_binary(filename, device=device, context=context, **kwargs)
    if ext in (".h5", ".hdf5", ".h5ebsd"):
        from .io_h5ebsd import load_h5ebsd

        return load_h5ebsd(filename, device=device, ctx=context, **kwargs)
    raise IOError(f"Could not read {filename!r}: only .dat, .up1, .up2, .ebsp and kikuchipy .h5 / .hdf5 / .h5ebsd files are read")
