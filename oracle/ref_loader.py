"""Run the reference's own metric / OSM modules in place (build container only).

TEST INFRASTRUCTURE ONLY - see ``oracle/__init__.py``.  ``/root/reference`` is
mounted read-only in the build container and does not exist on the GPU box, so
this loader is used (a) by ``tests/golden/make_golden.py`` to generate the
committed golden vectors and (b) by CPU tests that skip when the mount is absent.

kikuchipy cannot be imported as a package here (dask, hyperspy, orix, h5py ...
are not installed, no network).  Its similarity-metric modules only touch dask
through ``da.asarray / da.einsum / da.Array`` and ``dask.config.set`` and take
their NumPy branches when handed ndarrays, so a ~15-line ``dask`` stub that
routes those names to NumPy is enough to execute the reference's real cast /
mask / normalise / einsum lines unmodified.  ``_orientation_similarity_map.py``
needs only a ``CrystalMap`` name from orix for its type annotation.
"""

from __future__ import annotations

import contextlib
import importlib.util
import os
import sys
import types

import numpy as np

# Numba's on-disk cache (the reference's kernels are declared cache=True) defaults to a __pycache__
# directory NEXT TO THE SOURCE FILE when that directory is writable: keep it out of /root/reference
os.environ.setdefault("NUMBA_CACHE_DIR", os.path.join(os.environ.get("TMPDIR", "/tmp"), "kdi_numba_cache"))

REFERENCE_ROOT = "/root/reference"
_SRC = os.path.join(REFERENCE_ROOT, "src", "kikuchipy")


def available() -> bool:
    return os.path.isfile(
        os.path.join(_SRC, "indexing", "similarity_metrics", "_similarity_metric.py")
    )


def _install_stubs() -> None:
    if "dask" not in sys.modules:
        dask = types.ModuleType("dask")
        da = types.ModuleType("dask.array")
        cfg = types.ModuleType("dask.config")

        class Array:  # placeholder so isinstance(x, da.Array) is False for ndarrays
            pass

        da.Array = Array
        da.asarray = np.asarray
        da.einsum = lambda *a, **k: np.einsum(*a, **k)
        da.mean, da.sum, da.sqrt, da.square = np.mean, np.sum, np.sqrt, np.square
        cfg.set = lambda *a, **k: contextlib.nullcontext()
        dask.array, dask.config = da, cfg
        sys.modules.update({"dask": dask, "dask.array": da, "dask.config": cfg})
    if "orix" not in sys.modules:
        orix = types.ModuleType("orix")
        cm = types.ModuleType("orix.crystal_map")

        class CrystalMap:  # annotation only
            pass

        cm.CrystalMap = CrystalMap
        orix.crystal_map = cm
        sys.modules.update({"orix": orix, "orix.crystal_map": cm})
    for name in (
        "kikuchipy",
        "kikuchipy.indexing",
        "kikuchipy.indexing.similarity_metrics",
    ):
        if name not in sys.modules:
            mod = types.ModuleType(name)
            mod.__path__ = []  # mark as package
            sys.modules[name] = mod


def _load(modname: str, relpath: str):
    if modname in sys.modules and getattr(sys.modules[modname], "__file__", None):
        return sys.modules[modname]
    spec = importlib.util.spec_from_file_location(modname, os.path.join(_SRC, relpath))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    return mod


def load_metrics():
    """Return (SimilarityMetric, NCC class, NDP class, ncc_single_numba)."""
    if not available():
        raise RuntimeError("reference tree not mounted")
    _install_stubs()
    base = "kikuchipy.indexing.similarity_metrics"
    sm = _load(f"{base}._similarity_metric", "indexing/similarity_metrics/_similarity_metric.py")
    ncc = _load(
        f"{base}._normalized_cross_correlation",
        "indexing/similarity_metrics/_normalized_cross_correlation.py",
    )
    ndp = _load(
        f"{base}._normalized_dot_product",
        "indexing/similarity_metrics/_normalized_dot_product.py",
    )
    return (
        sm.SimilarityMetric,
        ncc.NormalizedCrossCorrelationMetric,
        ndp.NormalizedDotProductMetric,
        ncc._ncc_single_patterns_1d_float32_exp_centered,
    )


def load_osm():
    """Return the reference ``orientation_similarity_map`` function."""
    if not available():
        raise RuntimeError("reference tree not mounted")
    _install_stubs()
    mod = _load(
        "kikuchipy.indexing._orientation_similarity_map",
        "indexing/_orientation_similarity_map.py",
    )
    return mod.orientation_similarity_map


class FakeXmap:
    """Minimal object with what ``orientation_similarity_map`` reads."""

    def __init__(self, simulation_indices, shape, prop_name="simulation_indices"):
        self.prop = {prop_name: np.asarray(simulation_indices)}
        self.shape = tuple(shape)


def nickel_ebsd_small() -> np.ndarray:
    """The nine 60x60 uint8 patterns of ``kp.data.nickel_ebsd_small()`` read from
    the raw NORDIF file (data/_data.py:97-126)."""
    raw = np.fromfile(os.path.join(_SRC, "data", "nordif", "Pattern.dat"), dtype=np.uint8)
    return raw.reshape((3, 3, 60, 60))


def load_master_pattern():
    """Return the reference's ``signals/util/_master_pattern.py`` module, executed in place.

    It imports two helpers from modules that cannot be imported here
    (``kikuchipy/pattern/_pattern.py`` needs scikit-image): ``_rescale_with_min_max`` is taken
    from that file by executing only that function's own source lines (AST extraction, nothing
    is copied into this repository); ``kikuchipy/_utils/numba.py`` loads as is.  Needs numba.
    """
    if not available():
        raise RuntimeError("reference tree not mounted")
    import ast

    _install_stubs()
    for name in ("kikuchipy._utils", "kikuchipy.pattern", "kikuchipy.signals", "kikuchipy.signals.util",
                 "kikuchipy.detectors"):
        if name not in sys.modules:
            mod = types.ModuleType(name)
            mod.__path__ = []
            sys.modules[name] = mod
    _load("kikuchipy._utils.numba", "_utils/numba.py")
    if "kikuchipy.pattern._pattern" not in sys.modules:
        path = os.path.join(_SRC, "pattern", "_pattern.py")
        tree = ast.parse(open(path).read())
        fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "_rescale_with_min_max"]
        stub = types.ModuleType("kikuchipy.pattern._pattern")
        stub.__file__ = path
        from numba import njit

        # executed inside the stub module itself (it has a __name__): Numba's on-disk cache
        # (cache=True) re-imports the defining module by name when it loads an entry
        stub.njit, stub.np = njit, np
        sys.modules["kikuchipy.pattern._pattern"] = stub
        exec(compile(ast.Module(body=fn, type_ignores=[]), path, "exec"), stub.__dict__)
    return _load("kikuchipy.signals.util._master_pattern", "signals/util/_master_pattern.py")


def load_refinement():
    """Return the reference's ``indexing/_refinement/_solvers.py`` module, executed in place
    (with ``_objective_functions.py`` and everything they import).  ``kikuchipy/pattern/_pattern.py``
    cannot be imported here (scikit-image), so the three small Numba helpers the solvers take from
    it are compiled from their own source lines (AST extraction, nothing copied into this
    repository).  Needs numba and scipy."""
    import ast

    mp = load_master_pattern()
    load_metrics()
    from numba import njit

    path = os.path.join(_SRC, "pattern", "_pattern.py")
    stub = sys.modules["kikuchipy.pattern._pattern"]
    if not hasattr(stub, "_zero_mean_sum_square_1d_float32"):
        tree = ast.parse(open(path).read())
        want = ("_rescale_without_min_max_1d_float32", "_zero_mean_sum_square_1d_float32")
        fns = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in want]
        # executed inside the stub module itself: Numba's on-disk cache (cache=True) re-imports
        # the defining module by name
        stub.njit, stub.np = njit, np
        exec(compile(ast.Module(body=fns, type_ignores=[]), path, "exec"), stub.__dict__)
    if "kikuchipy.indexing._refinement" not in sys.modules or not getattr(
            sys.modules["kikuchipy.indexing._refinement"], "__file__", None):
        pkg = _load("kikuchipy.indexing._refinement", "indexing/_refinement/__init__.py")
        pkg.__path__ = []
    _load("kikuchipy._utils._gnonomic_bounds", "_utils/_gnonomic_bounds.py")
    _load("kikuchipy.indexing._refinement._objective_functions", "indexing/_refinement/_objective_functions.py")
    solvers = _load("kikuchipy.indexing._refinement._solvers", "indexing/_refinement/_solvers.py")
    return solvers, mp


def load_preprocessing():
    """Return a namespace with the reference's background-removal / neighbour-averaging functions,
    executed in place: ``filters/window.py`` and ``filters/fft_barnes.py`` load as they are (with a
    stub for matplotlib, which ``window.py`` imports for its plot method only); from
    ``pattern/_pattern.py`` (scikit-image) and ``pattern/chunk.py`` (dask) the needed functions are
    compiled from their own source lines (AST extraction, nothing copied into this repository)."""
    import ast

    load_master_pattern()  # stubs + kikuchipy.pattern._pattern with _rescale_with_min_max
    for name in ("matplotlib", "matplotlib.figure", "matplotlib.pyplot"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.Figure = object
            m.subplots = None
            sys.modules[name] = m
    sys.modules["matplotlib"].figure = sys.modules["matplotlib.figure"]
    if "kikuchipy.filters" not in sys.modules:
        pkg = types.ModuleType("kikuchipy.filters")
        pkg.__path__ = []
        sys.modules["kikuchipy.filters"] = pkg
    window = _load("kikuchipy.filters.window", "filters/window.py")
    barnes = _load("kikuchipy.filters.fft_barnes", "filters/fft_barnes.py")
    from numba import njit
    from scipy.ndimage import correlate, gaussian_filter

    stub = sys.modules["kikuchipy.pattern._pattern"]
    if not hasattr(stub, "_remove_dynamic_background"):
        path = os.path.join(_SRC, "pattern", "_pattern.py")
        want = ("_remove_static_background_subtract", "_remove_static_background_divide", "_remove_dynamic_background",
                "_remove_background_subtract", "_remove_background_divide", "_dynamic_background_frequency_space_setup")
        fns = [n for n in ast.parse(open(path).read()).body if isinstance(n, ast.FunctionDef) and n.name in want]
        stub.njit, stub.np, stub.Callable = njit, np, __import__("typing").Callable
        stub.Window, stub._fft_filter, stub._fft_filter_setup = window.Window, barnes._fft_filter, barnes._fft_filter_setup
        stub.gaussian_filter = gaussian_filter
        exec(compile(ast.Module(body=fns, type_ignores=[]), path, "exec"), stub.__dict__)
    if "kikuchipy.pattern.chunk" not in sys.modules:
        path = os.path.join(_SRC, "pattern", "chunk.py")
        chunk = types.ModuleType("kikuchipy.pattern.chunk")
        chunk.__file__ = path
        sys.modules["kikuchipy.pattern.chunk"] = chunk
        chunk.njit, chunk.np, chunk.correlate, chunk.Window = njit, np, correlate, window.Window
        chunk._rescale_with_min_max = stub._rescale_with_min_max
        want = ("_average_neighbour_patterns", "_rescale_neighbour_averaged_patterns")
        fns = [n for n in ast.parse(open(path).read()).body if isinstance(n, ast.FunctionDef) and n.name in want]
        exec(compile(ast.Module(body=fns, type_ignores=[]), path, "exec"), chunk.__dict__)
    ns = types.SimpleNamespace(Window=window.Window, fft_barnes=barnes, pattern=stub, chunk=sys.modules["kikuchipy.pattern.chunk"])
    return ns
