"""kikuchipy h5ebsd reader and writer: detector patterns from the file to the GPU (SURVEY.md section
8f.4, the data format on the experimental side of the path).

Mirrors /root/reference/src/kikuchipy/io/plugins/kikuchipy_h5ebsd/_api.py (``KikuchipyH5EBSDReader.
scan2dict`` :63-176, ``file_reader`` :179-217, ``KikuchipyH5EBSDWriter.write`` :365-462) and the shared
base class io/plugins/_h5ebsd.py (``check_file`` :177-212, ``get_manufacturer_version`` :214-241,
``get_desired_scan_groups`` :260-305, ``get_data`` :307-382, ``_hdf5group2dict`` :35-87,
``_dict2hdf5group`` :90-125): top-level ``manufacturer`` / ``version`` datasets, one group per scan
with ``EBSD/Data/patterns`` ``(ny * nx, sy, sx)``, ``EBSD/Header`` (scan and pattern sizes, steps,
detector angles, projection centres, static background), optional ``EBSD/CrystalMap`` and
``SEM/Header``.  Same selection rules, warnings and ``IOError`` texts.

HDF5 itself is read by the small parser in ``_hdf5.py`` (there is no h5py in this image); files written
with ``libver="latest"`` are refused with a message that says so.  No HyperSpy signal and no orix
``CrystalMap`` are built (out of scope): the result is a small container whose ``data`` every function
of this package accepts, and with ``device=True`` it is a CUDA tensor uploaded once through the
context's pinned ring, so preprocessing and indexing run without the patterns returning to the host.
"""

from __future__ import annotations

import os
import warnings

import numpy as np

from . import _hdf5
from .refinement import Detector

SUPPORTED_MANUFACTURERS = ["bruker nano", "edax", "kikuchipy", "oxford instruments"]


class H5EBSDScan:
    """``data`` ``(ny, nx, sy, sx)`` squeezed like the reference's, ``static_background`` ``(sy, sx)``
    or ``None``, ``detector`` (:class:`~kikuchipy_b200.refinement.Detector`; ``pc`` per map point when
    the file holds one per point), ``step_sizes`` ``(dy, dx)``, ``metadata``, ``original_metadata``
    (manufacturer, version and the whole ``EBSD/Header``), ``xmap`` (the ``crystal_map`` group as a
    nested dictionary, or ``None``), ``axes`` (the reference's axis descriptions)."""

    def __init__(self, data, static_background, detector, step_sizes, metadata, original_metadata, xmap, axes):
        self.data = data
        self.static_background = static_background
        self.detector = detector
        self.step_sizes = step_sizes
        self.metadata = metadata
        self.original_metadata = original_metadata
        self.xmap = xmap
        self.axes = axes

    @property
    def shape(self):
        return tuple(self.data.shape)


def _group_to_dict(group, recursive=False, skip=()):
    """``_hdf5group2dict``: one-element datasets become scalars, byte strings are decoded as latin-1,
    subgroups are followed when ``recursive``."""
    out = {}
    for key in group.keys():
        try:
            val = group[key]
        except NotImplementedError as e:  # a datatype this parser does not know: not needed for the scan
            warnings.warn(f"Could not read {key!r} from {group.name!r}: {e}")
            continue
        if isinstance(val, _hdf5.Dataset):
            if key in skip:
                continue
            val = val[()]
            if isinstance(val, np.ndarray) and val.ndim > 0 and len(val) == 1:
                val = val[0]
                key = key.lstrip()
            if isinstance(val, (bytes, np.bytes_)):
                val = bytes(val).decode("latin-1")
            elif isinstance(val, np.generic):
                val = val.item()
            out[key] = val
        elif recursive:
            out[key] = _group_to_dict(val, True, skip)
    return out


def _axes_list(data_shape, data_scale):
    """``H5EBSDReader.get_axes_list`` (:384-447)."""
    ny, nx = data_shape[:2]
    dy, dx, px_size = data_scale
    ndim = int(ny != 1) + int(nx != 1) + 2
    names = ["y", "x", "dy", "dx"]
    scales = [1.0 * dy, 1.0 * dx, 1.0 * px_size, 1.0 * px_size]
    shape = list(data_shape)
    if ndim == 3:
        drop = 1 if ny > nx else 0
        for seq in (names, scales, shape):
            del seq[drop]
    elif ndim == 2:
        names, scales, shape = names[2:], scales[2:], shape[2:]
    return [{"size": shape[i], "index_in_array": i, "name": names[i], "scale": scales[i], "offset": 0.0,
             "units": "um"} for i in range(ndim)]


def _manufacturer_version(f):
    manufacturer = version = None
    for key, val in _group_to_dict(f.root).items():
        if key.lower() == "manufacturer":
            manufacturer = str(val).lower()
        elif key.lower() in ("version", "format version"):
            version = str(val).lower()
    if manufacturer is None:
        raise IOError(f"Could not find 'manufacturer' key in {f.filename!r}")
    if version is None:
        raise IOError(f"Could not find 'version' key in {f.filename!r}")
    return manufacturer, version


def _scan2dict(f, name, group, manufacturer, version, ctx, device):
    header = _group_to_dict(group["EBSD/Header"], recursive=True)
    ny, nx = int(header["n_rows"]), int(header["n_columns"])
    sy, sx = int(header["pattern_height"]), int(header["pattern_width"])
    dy, dx = header.get("step_y", 1), header.get("step_x", 1)
    px_size = header.get("detector_pixel_size", 1)
    fname = os.path.basename(f.filename).split(".")[0]
    title = fname + " " + name
    if len(title) > 20:
        title = f"{title:.20}..."
    sem = _group_to_dict(group["SEM/Header"]) if "SEM/Header" in group else {}
    metadata = {
        "Acquisition_instrument": {"SEM": sem},
        "General": {"original_filename": fname, "title": title},
        "Signal": {"signal_type": "EBSD", "record_by": "image"},
    }
    try:
        dset = group["EBSD/Data/patterns"]
    except KeyError:
        raise KeyError("Could not find patterns in the expected dataset 'EBSD/Data/patterns'")
    data = dset.read()
    try:
        data = data.reshape((ny, nx, sy, sx)).squeeze()
    except ValueError:
        warnings.warn(
            f"Signal shape ({sy}, {sy}) and navigation shape ({ny}, {nx}) larger than file size. Will attempt to "
            "load by zero padding incomplete patterns."
        )
        flat = data.ravel()
        data = np.pad(flat, [(0, ny * nx * sy * sx - flat.size)]).reshape((ny, nx, sy, sx))
    xmap = None
    if "EBSD/CrystalMap" in group and "crystal_map" in group["EBSD/CrystalMap"]:
        xmap = _group_to_dict(group["EBSD/CrystalMap/crystal_map"], recursive=True)
    static_bg = header.get("static_background")
    if not isinstance(static_bg, np.ndarray):  # the writer stores -1 for "none"
        static_bg = None
    pc = np.dstack((header.get("pcx", 0.5), header.get("pcy", 0.5), header.get("pcz", 0.5)))
    if pc.size > 3:
        try:
            pc = pc.reshape((ny, nx, 3))
        except ValueError:
            warnings.warn(
                f"Data navigation shape ({(ny, nx)}) greater than the navigation shape of the projection center "
                f"(PC) array {pc.shape[:2]}. Will attempt to pad PC array with (PCx, PCy, PCz) = (0.5, 0.5, 0.5)"
            )
            try:
                pad = [(0, d) for d in (np.array((ny, nx)) - np.array(pc.shape[:2]))] + [(0, 0)]
                pc = np.pad(pc, pad, constant_values=0.5)
            except ValueError:
                warnings.warn("Could not pad PC array, detector will have a single PC with (PCx, PCy, PCz) = "
                              "(0.5, 0.5, 0.5)")
                pc = np.array((0.5, 0.5, 0.5))
    pc = pc.squeeze()
    detector = Detector(
        shape=(sy, sx), pc=pc, tilt=float(header.get("elevation_angle", 0.0)),
        azimuthal=float(header.get("azimuth_angle", 0.0)), twist=float(header.get("twist_angle", 0.0)),
        sample_tilt=float(header.get("sample_tilt", 70.0)), px_size=float(px_size),
        binning=header.get("binning", 1),
    )
    original = {"manufacturer": manufacturer, "version": version}
    original.update(header)
    if device:
        from ._lib import default_context

        data = (ctx or default_context()).to_device(np.ascontiguousarray(data))
    return H5EBSDScan(data, static_bg, detector, (float(dy), float(dx)), metadata, original, xmap,
                      _axes_list((ny, nx, sy, sx), (dy, dx, px_size)))


def load_h5ebsd(filename, scan_group_names=None, device=False, ctx=None):
    """Read one or more scans from a kikuchipy h5ebsd file.  Without ``scan_group_names`` the first scan
    is returned (an :class:`H5EBSDScan`); with a name, that scan; with a list of names, a list."""
    with _hdf5.File(filename) as f:
        manufacturer, version = _manufacturer_version(f)
        groups = [(k, f.root[k]) for k in f.root.keys()]
        groups = [(k, g) for k, g in groups if isinstance(g, _hdf5.Group)]
        error = None
        if not any("EBSD/Data" in g and "EBSD/Header" in g for _, g in groups):
            error = "no top groups with subgroup name 'EBSD' with subgroups 'Data' and 'Header' were found"
        elif manufacturer not in SUPPORTED_MANUFACTURERS:
            error = f"{manufacturer!r} is not among supported manufacturers {SUPPORTED_MANUFACTURERS}"
        if error is not None:
            raise IOError(f"{f.filename} is not a supported h5ebsd file, as {error}")
        if manufacturer != "kikuchipy":
            raise NotImplementedError(f"only kikuchipy h5ebsd files are read here, not {manufacturer!r} ones")
        names = [k for k, _ in groups]
        if scan_group_names is None:
            wanted = [groups[0]]
        else:
            asked = [scan_group_names] if isinstance(scan_group_names, str) else list(scan_group_names)
            wanted = []
            for want in asked:
                hit = [(k, g) for k, g in groups if k == want]
                if hit:
                    wanted.append(hit[0])
                else:
                    msg = f"Scan {want!r} is not among the available scans {names} in {f.filename!r}"
                    if len(asked) == 1:
                        raise IOError(msg)
                    warnings.warn(msg)
        scans = [_scan2dict(f, k, g, manufacturer, version, ctx, device) for k, g in wanted]
    if scan_group_names is None or isinstance(scan_group_names, str):
        return scans[0]
    return scans


def save_h5ebsd(filename, data, detector=None, static_background=None, step_sizes=(1.0, 1.0), metadata=None,
                xmap=None, scan_number=1, add_scan=False, version="0.1.0"):
    """Write patterns ``(ny, nx, sy, sx)`` (or ``(n, sy, sx)`` / ``(sy, sx)``) as scan ``scan_number`` of a
    kikuchipy h5ebsd file with the reference writer's layout.  ``add_scan``: keep the scans an existing
    kikuchipy file already holds (the file is rewritten: this writer has no in-place update)."""
    if hasattr(data, "is_cuda"):
        data = data.cpu().numpy()
    data = np.asarray(data)
    if data.ndim == 2:
        data = data[None, None]
    elif data.ndim == 3:
        data = data[None]
    if data.ndim != 4:
        raise ValueError("patterns must have two signal and at most two navigation dimensions")
    ny, nx, sy, sx = data.shape
    tree = {}
    if add_scan and os.path.isfile(filename):
        with _hdf5.File(filename) as f:
            top = _group_to_dict(f.root)
            if top.get("manufacturer") != "kikuchipy":
                raise IOError(f"{filename} is not a supported kikuchipy h5ebsd file, as it was not created with kikuchipy")
            tree = _tree_of(f.root)
        if f"Scan {scan_number}" in tree:
            raise IOError("Invalid scan number")
    else:
        tree = {"manufacturer": "kikuchipy", "version": version}
    det = detector if detector is not None else Detector(shape=(sy, sx))
    pc = np.asarray(det.pc, dtype=np.float64)
    if pc.ndim == 1:
        pc = pc.reshape(1, 3)
    dy, dx = step_sizes
    sem = (metadata or {}).get("Acquisition_instrument", {}).get("SEM", {})
    scan = {
        "EBSD": {
            "Data": {"patterns": data.reshape(ny * nx, sy, sx)},
            "Header": {
                "azimuth_angle": float(det.azimuthal), "twist_angle": float(getattr(det, "twist", 0.0)),
                "binning": getattr(det, "binning", 1), "elevation_angle": float(det.tilt),
                "n_columns": int(nx), "n_rows": int(ny), "pattern_width": int(sx), "pattern_height": int(sy),
                "pcx": pc[..., 0], "pcy": pc[..., 1], "pcz": pc[..., 2],
                "detector_pixel_size": float(getattr(det, "px_size", 1.0)), "sample_tilt": float(det.sample_tilt),
                "static_background": -1 if static_background is None else np.asarray(static_background),
                "step_x": float(dx), "step_y": float(dy),
            },
        },
        "SEM": {"Header": {
            "beam_energy": sem.get("beam_energy", 0), "magnification": sem.get("magnification", 0),
            "microscope": sem.get("microscope", ""), "working_distance": sem.get("working_distance", 0)}},
    }
    if xmap is not None:
        scan["EBSD"]["CrystalMap"] = {"manufacturer": "orix", "version": "0.0", "crystal_map": xmap}
    tree[f"Scan {scan_number}"] = scan
    _hdf5.write(filename, tree)


def _tree_of(group):
    """A group as the nested dictionary ``_hdf5.write`` takes (arrays as stored)."""
    out = {}
    for key in group.keys():
        val = group[key]
        out[key] = _tree_of(val) if isinstance(val, _hdf5.Group) else val.read()
    return out
