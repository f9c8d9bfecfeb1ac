"""``EBSD.refine_orientation()``, ``refine_projection_center()`` and
``refine_orientation_projection_center()`` on the GPU.

Mirrors /root/reference/src/kikuchipy/signals/ebsd.py:1986-2560 (the three public methods),
``indexing/_refinement/_refinement.py:340-870`` (set-up, trust-region bounds, result assembly,
printed messages) for the reference's DEFAULT optimiser, ``scipy.optimize.minimize`` with
``method="Nelder-Mead"``.  The per-pattern work - pattern preparation, projection of the master
pattern, NCC, the simplex search - is one call into ``libkdi`` (``kdi_refine``, csrc/kdi_refine.cu).
Every other SciPy optimiser the reference offers (``minimize`` with another ``method``,
``basinhopping``, ``differential_evolution``, ``dual_annealing``, ``shgo``;
``_refinement/__init__.py:32-60``) runs its own (host) search loop, as it does in the reference, and gets
its objective values from the device: ``kdi_refine_objective`` evaluates ``1 - NCC`` for a batch of
(pattern, parameters) pairs, and the searches of many patterns advance side by side so that the batches
are large (``_HostDriven``).  NLopt (``ln_neldermead``) is not installed here and raises like the
reference does without it.

Inputs are read by duck typing: kikuchipy signals / orix crystal maps / ``EBSDDetector`` work, and
so do plain arrays, :class:`~kikuchipy_b200.indexing.DictionaryIndexingResult` and
:class:`Detector` (the geometry an ``EBSDDetector`` contributes to this path).
"""

from __future__ import annotations

import sys
import time

import numpy as np

from . import _lib
from .indexing import _unwrap
from .master_pattern import direction_cosines as _direction_cosines


def sample_to_detector_matrix(sample_tilt=70.0, tilt=0.0, azimuthal=0.0, twist=0.0):
    """Rows = detector axes (X_d, Y_d, Z_d) in sample coordinates (``detectors/_ebsd_detector.py
    :100-149``): the basis ((0,1,0), (0,0,1), (1,0,0)) turned about its own first axis by
    ``-sample_tilt`` and ``+tilt``, about its second by ``-azimuthal`` and its third by ``-twist``
    (degrees)."""
    basis = np.array([[0, 1, 0], [0, 0, 1], [1, 0, 0]], dtype=np.float64)
    angles = np.deg2rad(np.array([-sample_tilt, tilt, -azimuthal, -twist], dtype=np.float64))
    for axis, angle in zip((0, 0, 1, 2), angles):
        u = basis[axis] / np.sqrt(np.sum(basis[axis] ** 2))
        c, s = np.cos(angle), np.sin(angle)
        for j in range(3):  # Rodrigues' formula, in place row by row like the reference
            v = basis[j].copy()
            basis[j] = v * c + np.cross(u, v) * s + u * np.dot(u, v) * (1.0 - c)
    return basis


class Detector:
    """The geometry an ``EBSDDetector`` contributes to refinement: ``shape`` (rows, columns), one
    projection centre or one per map point (Bruker convention) and the tilts in degrees."""

    def __init__(self, shape, pc=(0.5, 0.5, 0.5), sample_tilt=70.0, tilt=0.0, azimuthal=0.0, twist=0.0,
                 px_size=1.0, binning=1):
        self.shape = (int(shape[0]), int(shape[1]))
        self.nrows, self.ncols = self.shape
        self.pc = np.atleast_2d(np.asarray(pc, dtype=np.float64))
        self.sample_tilt, self.tilt, self.azimuthal, self.twist = sample_tilt, tilt, azimuthal, twist
        self.px_size, self.binning = px_size, binning  # carried for the file formats; not used by refinement

    @property
    def navigation_shape(self):
        return self.pc.shape[:-1]

    @property
    def navigation_size(self):
        return int(np.prod(self.navigation_shape))

    @property
    def pc_flattened(self):
        return self.pc.reshape(-1, 3)

    @property
    def pcx(self):
        return self.pc[..., 0]

    @property
    def pcy(self):
        return self.pc[..., 1]

    @property
    def pcz(self):
        return self.pc[..., 2]

    @property
    def om_detector_to_sample(self):
        return sample_to_detector_matrix(self.sample_tilt, self.tilt, self.azimuthal, self.twist).T

    @property
    def gnomonic_bounds(self):
        pcx, pcy, pcz = self.pc[..., 0], self.pc[..., 1], self.pc[..., 2]
        ar = self.ncols / self.nrows
        return np.stack([-ar * (pcx / pcz), ar * (1 - pcx) / pcz, -(1 - pcy) / pcz, pcy / pcz], axis=-1)

    def deepcopy(self):
        import copy

        return copy.deepcopy(self)


class RefinementResult:
    """What callers read from the refined ``CrystalMap`` (``_refinement.py:57-130``) when orix is
    not installed: ``rotations`` (quaternions, ``(n points, 4)``), ``euler`` (radians), ``prop``
    (``scores``, ``num_evals`` and, with pseudo-symmetry operators, ``pseudo_symmetry_index``),
    ``is_in_data`` (the refined points), ``shape``."""

    def __init__(self, shape, is_in_data, euler, prop):
        self.shape = tuple(int(s) for s in shape)
        self.is_in_data = is_in_data
        self.euler = euler
        self.rotations = euler_to_quaternion(euler)
        self.prop = prop
        self.rotations_per_point = 1

    @property
    def size(self):
        return int(self.is_in_data.sum())

    def __getattr__(self, name):
        prop = self.__dict__.get("prop", {})
        if name in prop:
            return prop[name]
        raise AttributeError(name)


# ---- rotations (what the path takes from orix: Rotation.to_euler / from_euler) -------------------

def quaternion_to_euler(q):
    """Bunge-Euler angles (radians, each in [0, 2 pi)) of unit quaternions ``(..., 4)`` - orix
    ``Rotation.to_euler`` (third party, not installed here), restated from the published
    algorithm (Rowenhorst et al. 2015, qu2eu with P = 1), the inverse of the reference's own
    ``rotation_from_euler`` (``_utils/numba.py:44-58``)."""
    q = np.asarray(q, dtype=np.float64)
    a, b, c, d = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    q03, q12 = a * a + d * d, b * b + c * c
    chi = np.sqrt(q03 * q12)
    with np.errstate(divide="ignore", invalid="ignore"):
        phi1 = np.arctan2((b * d - a * c) / chi, (-a * b - c * d) / chi)
        Phi = np.arctan2(2 * chi, q03 - q12)
        phi2 = np.arctan2((a * c + b * d) / chi, (c * d - a * b) / chi)
    flat = chi == 0
    only03 = flat & (q12 == 0)
    only12 = flat & ~only03
    phi1 = np.where(only03, np.arctan2(-2 * a * d, a * a - d * d), phi1)
    phi1 = np.where(only12, np.arctan2(2 * b * c, b * b - c * c), phi1)
    Phi = np.where(only03, 0.0, np.where(only12, np.pi, Phi))
    phi2 = np.where(flat, 0.0, phi2)
    eu = np.stack([phi1, Phi, phi2], axis=-1)
    return np.where(eu < 0, eu + 2 * np.pi, eu)


def euler_to_quaternion(eu):
    """``rotation_from_euler`` (``_utils/numba.py:44-58``) for arrays ``(..., 3)``."""
    eu = np.asarray(eu, dtype=np.float64)
    sigma, delta = 0.5 * (eu[..., 0] + eu[..., 2]), 0.5 * (eu[..., 0] - eu[..., 2])
    c, s = np.cos(0.5 * eu[..., 1]), np.sin(0.5 * eu[..., 1])
    q = np.stack([c * np.cos(sigma), -s * np.cos(delta), -s * np.sin(delta), -c * np.sin(sigma)], axis=-1)
    return np.where(q[..., :1] < 0, -q, q)


def _quaternion_multiply(p, q):
    a1, b1, c1, d1 = (p[..., i] for i in range(4))
    a2, b2, c2, d2 = (q[..., i] for i in range(4))
    return np.stack([a1 * a2 - b1 * b2 - c1 * c2 - d1 * d2, a1 * b2 + b1 * a2 + c1 * d2 - d1 * c2,
                     a1 * c2 - b1 * d2 + c1 * a2 + d1 * b2, a1 * d2 + b1 * c2 - c1 * b2 + d1 * a2], axis=-1)


# ---- input handling -----------------------------------------------------------------------------

def _detector_geometry(detector):
    nrows, ncols = (int(s) for s in detector.shape)
    pc = np.asarray(getattr(detector, "pc_flattened", None) if hasattr(detector, "pc_flattened") else detector.pc,
                    dtype=np.float64).reshape(-1, 3)
    if hasattr(detector, "om_detector_to_sample"):
        om = np.asarray(detector.om_detector_to_sample, dtype=np.float64)
    else:
        om = np.asarray((~detector.sample_to_detector).to_matrix(), dtype=np.float64).squeeze()
    return nrows, ncols, pc, np.ascontiguousarray(om.reshape(3, 3))


def _xmap_rotations(xmap, points_in_data):
    """Quaternions ``(n, 4)`` of the best match of the points to refine (``_refinement.py:966-970``)."""
    r = xmap.rotations if hasattr(xmap, "rotations") else xmap
    if r is None:
        raise ValueError("The crystal map has no rotations to refine (index against a dictionary with rotations)")
    r = np.asarray(r.data if hasattr(r, "data") and not isinstance(r, np.ndarray) else r, dtype=np.float64)
    if r.ndim == 3:
        r = r[:, 0]
    if r.shape[0] != points_in_data.size:
        raise ValueError(f"Crystal map has {r.shape[0]} rotations, expected {points_in_data.size}")
    return r[points_in_data]


def _master_pattern_arrays(master_pattern, energy):
    """``_get_master_pattern_data`` (``_refinement.py:1281-1320``): float32 hemispheres."""
    if isinstance(master_pattern, (tuple, list)):
        mpu, mpl = (np.asarray(m) for m in master_pattern)
    elif hasattr(master_pattern, "_get_master_pattern_arrays_from_energy"):
        mpu, mpl = master_pattern._get_master_pattern_arrays_from_energy(energy=energy)
    else:
        mpu = mpl = np.asarray(master_pattern)
    if mpu.ndim != 2 or mpu.shape != mpl.shape:
        raise ValueError("the master pattern must be given as two 2-D hemispheres of equal shape")
    if mpu.dtype != np.float32:
        # the reference rescales to float32 [-1, 1] (rescale_intensity); NCC is invariant to that
        # affine map and the bilinear interpolation is linear, so the values are passed on as they are
        mpu, mpl = mpu.astype(np.float32), mpl.astype(np.float32)
    return np.ascontiguousarray(mpu), np.ascontiguousarray(mpl)


_NM_NAMES = ("nelder-mead", "neldermead")

# _refinement/__init__.py:32-60
SUPPORTED_OPTIMIZATION_METHODS = {
    "minimize": {"type": "local", "supports_bounds": True, "package": "scipy"},
    "ln_neldermead": {"type": "local", "supports_bounds": True, "package": "nlopt"},
    "basinhopping": {"type": "global", "supports_bounds": False, "package": "scipy"},
    "differential_evolution": {"type": "global", "supports_bounds": True, "package": "scipy"},
    "dual_annealing": {"type": "global", "supports_bounds": True, "package": "scipy"},
    "shgo": {"type": "global", "supports_bounds": True, "package": "scipy"},
}


def _method_plan(method, method_kwargs):
    """``_RefinementSetup.set_optimization_parameters`` (``_refinement.py:1053-1139``): which search runs
    where.  Returns ``("device", options, name, "local")`` for SciPy's Nelder-Mead (the whole search on
    the GPU) or ``("host", (function name, keyword arguments), name, type)``."""
    method = "minimize" if method is None else str(method).lower()
    if method not in SUPPORTED_OPTIMIZATION_METHODS:
        raise ValueError(f"Method {method!r} not in the list of supported methods {list(SUPPORTED_OPTIMIZATION_METHODS)}")
    info = SUPPORTED_OPTIMIZATION_METHODS[method]
    if info["package"] == "nlopt":
        raise ImportError(f"Optimization method {method.upper()!r} requires nlopt, which is not installed")
    kw = dict(method_kwargs or {})
    if method == "minimize" and "method" not in kw:
        name = kw["method"] = "Nelder-Mead"
    elif "method" in kw:
        name = str(kw["method"])
    else:
        name = method
    if method == "basinhopping" and "minimizer_kwargs" not in kw:
        kw["minimizer_kwargs"] = {}
    if method == "minimize" and name.lower() in _NM_NAMES:
        try:
            opts, _ = _nelder_mead_options("minimize", kw)
            return "device", opts, name, info["type"], kw
        except NotImplementedError:
            pass  # an option the device search does not implement: SciPy's own loop, objective on the device
    return "host", (method, kw), name, info["type"], kw


def _supports_bounds(method):
    return SUPPORTED_OPTIMIZATION_METHODS["minimize" if method is None else str(method).lower()]["supports_bounds"]


class _HostDriven:
    """SciPy's optimisers on the host, the objective on the device.

    One Python thread per pattern in flight runs the optimiser exactly as the reference calls it
    (``_solvers.py:186-254, 300-345, 420-470``); its objective function parks the thread until the
    coordinator has gathered the requests of all threads that are currently waiting and evaluated them
    in one ``kdi_refine_objective`` launch.  Vectorised calls (``differential_evolution(vectorized=True)``
    hands over a whole population) become several points of one row."""

    def __init__(self, evaluate, n_workers):
        import threading

        self._evaluate = evaluate  # (pattern ids, list of (points, nv) arrays) -> list of (points,) arrays
        self._cv = threading.Condition()
        self._pending = []   # (pattern id, x (points, nv), slot)
        self._running = 0    # worker threads that are neither finished nor parked in the objective
        self._n_workers = n_workers
        self.launches = 0

    def objective(self, pattern_id):
        def f(x, *_):
            x = np.asarray(x, dtype=np.float64)
            pts = x.reshape(1, -1) if x.ndim == 1 else x.T  # vectorised callers pass (nv, points)
            slot = {}
            with self._cv:
                self._pending.append((pattern_id, np.ascontiguousarray(pts), slot))
                self._running -= 1
                self._cv.notify_all()
                while "y" not in slot and "err" not in slot:
                    self._cv.wait()
            if "err" in slot:
                raise slot["err"]
            return float(slot["y"][0]) if x.ndim == 1 else slot["y"]
        return f

    def run(self, jobs):
        """``jobs``: callables taking (pattern position) -> result; returns their results in order."""
        import queue
        import threading

        results = [None] * len(jobs)
        errors = []
        todo = queue.Queue()
        for i, job in enumerate(jobs):
            todo.put((i, job))
        n_threads = max(1, min(self._n_workers, len(jobs)))

        def worker():
            while True:
                try:
                    i, job = todo.get_nowait()
                except queue.Empty:
                    break
                try:
                    results[i] = job()
                except BaseException as e:  # noqa: BLE001
                    errors.append(e)
            with self._cv:
                self._running -= 1
                self._cv.notify_all()

        with self._cv:
            self._running = n_threads
        threads = [threading.Thread(target=worker, daemon=True) for _ in range(n_threads)]
        for t in threads:
            t.start()
        while True:
            with self._cv:
                while self._running > 0:
                    self._cv.wait()
                batch, self._pending = self._pending, []
                if not batch:
                    break  # every worker has finished
            try:
                ys = self._evaluate([b[0] for b in batch], [b[1] for b in batch])
                self.launches += 1
                with self._cv:
                    for (_, _, slot), y in zip(batch, ys):
                        slot["y"] = y
                    self._running += len(batch)
                    self._cv.notify_all()
            except BaseException as e:  # noqa: BLE001
                with self._cv:
                    for _, _, slot in batch:
                        slot["err"] = e
                    self._running += len(batch)
                    self._cv.notify_all()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        return results


def _nelder_mead_options(method, method_kwargs):
    """Tolerances and limits SciPy would use (``minimize`` + ``_minimize_neldermead``)."""
    if method is None or str(method).lower() != "minimize":
        raise NotImplementedError(
            f"Method {method!r} is not implemented on the GPU: only 'minimize' with method='Nelder-Mead' "
            "(the reference's default) is; there is no CPU fallback"
        )
    kw = dict(method_kwargs or {})
    name = str(kw.pop("method", "Nelder-Mead"))
    if name.lower() not in _NM_NAMES:
        raise NotImplementedError(f"scipy.optimize.minimize method {name!r} is not implemented on the GPU")
    tol = kw.pop("tol", None)
    options = dict(kw.pop("options", None) or {})
    if kw:
        raise NotImplementedError(f"method_kwargs {sorted(kw)} are not supported")
    xatol = options.pop("xatol", 1e-4 if tol is None else tol)
    fatol = options.pop("fatol", 1e-4 if tol is None else tol)
    maxiter, maxfev = options.pop("maxiter", None), options.pop("maxfev", None)
    adaptive = bool(options.pop("adaptive", False))
    options.pop("disp", None)
    if options:
        raise NotImplementedError(f"Nelder-Mead options {sorted(options)} are not supported")

    def limit(v):
        return -1 if v is None else (np.iinfo(np.int64).max if v == np.inf else int(v))

    return {"xatol": float(xatol), "fatol": float(fatol), "maxiter": limit(maxiter), "maxfev": limit(maxfev),
            "adaptive": adaptive}, name


def _bounds(x0, trust_region, mode):
    """``_RefinementSetup.get_bound_constraints`` (``_refinement.py:1192-1247``)."""
    if trust_region is None:
        return None, None
    leeway = np.deg2rad(5)
    eu_lo, eu_hi = 3 * [-leeway], [2 * np.pi + leeway, np.pi + leeway, 2 * np.pi + leeway]
    pc_lo, pc_hi = 3 * [-2], 3 * [2]
    tr = np.asarray(trust_region, dtype=np.float64).copy()
    if mode == "ori":
        tr = np.deg2rad(tr)
        lo, hi = eu_lo, eu_hi
    elif mode == "pc":
        lo, hi = pc_lo, pc_hi
    else:
        tr[:3] = np.deg2rad(tr[:3])
        lo, hi = eu_lo + pc_lo, eu_hi + pc_hi
    return np.fmax(x0 - tr, lo), np.fmin(x0 + tr, hi)


def _info_message(method_name, trust_region, method_kwargs, n_ps, kind="local", supports_bounds=True):
    """``_RefinementSetup.get_info_message`` (``_refinement.py:1249-1278``)."""
    info = f"Refinement information:\n  Method: {method_name} ({kind}) from SciPy"
    if supports_bounds:
        info += "\n  Trust region (+/-): " + np.array_str(np.asarray(trust_region), precision=5)
    info += f"\n  Keyword arguments passed to method: {method_kwargs}"
    if n_ps > 0:
        info += f"\n  No. pseudo-symmetry operators: {n_ps}"
    return info


class _Setup:
    """Everything the three entry points share: points to refine, patterns, masks, geometry."""

    def __init__(self, signal, xmap, detector, master_pattern, energy, navigation_mask, signal_mask, context,
                 sharded=False, group=None):
        self.sharded, self.group = False, group
        if sharded:
            import torch.distributed as dist

            self.sharded = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        data, nav_shape, sig_shape, _, _, _ = _unwrap(signal)
        self.nav_shape = tuple(nav_shape)
        self.nav_size = int(np.prod(nav_shape)) if len(nav_shape) else 1
        self.nrows, self.ncols, self.pc, self.om = _detector_geometry(detector)
        if (self.nrows, self.ncols) != tuple(sig_shape):
            raise ValueError(f"Detector shape {(self.nrows, self.ncols)} must be equal to the signal shape {tuple(sig_shape)}")
        if self.pc.shape[0] not in (1, self.nav_size):
            raise ValueError("Detector must have exactly one projection center (PC), or one PC per pattern in an "
                             f"array of shape signal.axes_manager.navigation_shape[::-1] + (3,), but was {self.pc.shape}")
        if signal_mask is not None and tuple(sig_shape) != np.shape(signal_mask):
            raise ValueError(f"Signal mask shape {np.shape(signal_mask)} and signal's signal shape "
                             f"{tuple(sig_shape)} must be the same shape")
        is_in_data = np.asarray(getattr(xmap, "is_in_data", np.ones(self.nav_size, dtype=bool)), dtype=bool)
        if is_in_data.size != self.nav_size:
            raise ValueError(f"Crystal map shape {getattr(xmap, 'shape', None)} and the signal's navigation shape "
                             f"{self.nav_shape} must be the same")
        points = is_in_data.copy()
        if navigation_mask is not None:
            if np.shape(navigation_mask) != self.nav_shape:
                raise ValueError(f"Navigation mask shape {np.shape(navigation_mask)} and crystal map shape "
                                 f"{self.nav_shape} must be the same")
            points &= ~np.asarray(navigation_mask, dtype=bool).ravel()
        if not points.any():
            raise ValueError("No points to refine")
        self.is_in_data = is_in_data
        self.points = points
        self.n = int(points.sum())
        self.quaternions = _xmap_rotations(xmap, points[is_in_data])
        pats = np.asarray(data).reshape(self.nav_size, self.nrows * self.ncols)
        self.patterns = np.ascontiguousarray(pats if points.all() else pats[points])
        self.rescale = self.patterns.dtype == np.float32  # _refinement.py:956
        self.unique_pc = self.pc.shape[0] > 1
        self.pcs = self.pc[points] if self.unique_pc else np.tile(self.pc[0], (self.n, 1))
        self.signal_mask = None if signal_mask is None else np.asarray(signal_mask, dtype=bool)
        self.ctx = context if context is not None else _lib.default_context()
        self.mpu, self.mpl = _master_pattern_arrays(master_pattern, energy)
        self.detector = detector

    def run(self, mode, x0, lower, upper, rotations, pcs, opts, fixed_dc, host_plan=None, trust_region_passed=False):
        """One device call for this process's patterns.  With ``sharded`` (one process per GPU,
        ``torch.distributed`` initialised) the patterns are split into contiguous balanced slices,
        every rank refines its own and the finished rows are all-gathered: the path partitions by
        pattern, there is no exchange step."""
        if self.sharded:
            import torch.distributed as dist

            from .distributed import gather_rows, shard_bounds

            world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
            bounds = [shard_bounds(self.n, world, r) for r in range(world)]
            a, b = bounds[rank]

            def cut(arr):
                return None if arr is None else arr[a:b]

            if host_plan is not None:
                whole, self.patterns = self.patterns, self.patterns[a:b]
                try:
                    local = self.run_host_driven(mode, x0[a:b], cut(lower), cut(upper), cut(rotations), cut(pcs),
                                                 host_plan, fixed_dc, trust_region_passed)
                finally:
                    self.patterns = whole
            else:
                local = self._run_local(mode, x0[a:b], cut(lower), cut(upper), cut(rotations), cut(pcs), opts, fixed_dc,
                                        self.patterns[a:b])
            return gather_rows(local, [e - s for s, e in bounds], self.group)
        if host_plan is not None:
            return self.run_host_driven(mode, x0, lower, upper, rotations, pcs, host_plan, fixed_dc, trust_region_passed)
        return self._run_local(mode, x0, lower, upper, rotations, pcs, opts, fixed_dc, self.patterns)

    def _master_pattern_on_device(self, fixed_dc):
        if fixed_dc:
            gb = [-(self.ncols / self.nrows) * (self.pc[0, 0] / self.pc[0, 2]),
                  (self.ncols / self.nrows) * (1 - self.pc[0, 0]) / self.pc[0, 2],
                  -(1 - self.pc[0, 1]) / self.pc[0, 2], self.pc[0, 1] / self.pc[0, 2]]
            if hasattr(self.detector, "gnomonic_bounds"):
                gb = np.asarray(self.detector.gnomonic_bounds, dtype=np.float64).reshape(-1, 4)[0]
            dc = _direction_cosines(gb, self.pc[0, 2], self.nrows, self.ncols, self.om)
        else:
            dc = np.zeros((self.nrows * self.ncols, 3))  # unused: computed per pattern on the device
        return self.ctx.master_pattern(self.mpu, self.mpl, dc)

    def run_host_driven(self, mode, x0, lower, upper, rotations, pcs, plan, fixed_dc, trust_region_passed,
                        n_workers=64):
        """The reference's ``_refine_*_solver_scipy`` for every pattern, with SciPy's optimiser on the host
        and the objective on the device (``_HostDriven``).  Same result rows as ``run``."""
        import copy

        import scipy.optimize

        fname, kwargs = plan
        func = getattr(scipy.optimize, fname)
        supports_bounds = SUPPORTED_OPTIMIZATION_METHODS[fname]["supports_bounds"]
        ctx = self.ctx
        n, n_starts, nv = x0.shape
        if n == 0:
            return np.zeros((0, 2 + nv + (1 if n_starts > 1 else 0)))
        ctx.set_signal_mask(self.signal_mask)
        try:
            mp = self._master_pattern_on_device(fixed_dc)
            pats = ctx.device_rows(self.patterns, n)  # uploaded once, read by every objective launch

            def evaluate(ids, xs):
                # one row per request; requests with several points (vectorised optimisers) are padded to
                # the longest of the batch with copies of their first point
                width = max(x.shape[0] for x in xs)
                x = np.stack([np.concatenate([xi, np.repeat(xi[:1], width - xi.shape[0], axis=0)]) for xi in xs])
                rows = np.asarray([i // n_starts for i in ids], dtype=np.int64)
                rot = None if rotations is None else np.repeat(rotations[rows][:, None, :], width, axis=1)
                pc = None if pcs is None else pcs[rows]
                y = ctx.refine_objective(mp, mode, pats, self.nrows, self.ncols, self.rescale, rows, x, rot, pc, self.om)
                return [y[k, :xs[k].shape[0]] for k in range(len(xs))]

            host = _HostDriven(evaluate, n_workers)

            def job(i, st):
                f = host.objective(i * n_starts + st)
                kw = copy.deepcopy(kwargs)
                if fname == "minimize":
                    if trust_region_passed:
                        kw["bounds"] = np.stack([lower[i, st], upper[i, st]], axis=1)
                    return lambda: func(fun=f, x0=x0[i, st], **kw)
                if supports_bounds:
                    return lambda: func(func=f, bounds=np.stack([lower[i, st], upper[i, st]], axis=1), **kw)
                return lambda: func(func=f, x0=x0[i, st], **kw)  # basinhopping

            if supports_bounds and fname != "minimize" and lower is None:
                raise ValueError(f"Method {fname!r} needs bounds: pass a trust_region")
            res = host.run([job(i, st) for i in range(n) for st in range(n_starts)])
        finally:
            ctx.set_signal_mask(None)
        out = np.zeros((n, 2 + nv + (1 if n_starts > 1 else 0)))
        for i in range(n):
            rs = res[i * n_starts:(i + 1) * n_starts]
            ncc = [1 - r.fun for r in rs]
            best = int(np.argmax(ncc))  # _solvers.py:236-254
            out[i, 0], out[i, 1], out[i, 2:2 + nv] = ncc[best], rs[best].nfev, rs[best].x
            if n_starts > 1:
                out[i, -1] = best
        self.objective_launches = host.launches
        return out

    def _run_local(self, mode, x0, lower, upper, rotations, pcs, opts, fixed_dc, patterns):
        ctx = self.ctx
        if x0.shape[0] == 0:
            return np.zeros((0, 2 + x0.shape[2] + (1 if x0.shape[1] > 1 else 0)))
        ctx.set_signal_mask(self.signal_mask)
        try:
            mp = self._master_pattern_on_device(fixed_dc)
            return ctx.refine(mp, mode, patterns, self.nrows, self.ncols, self.rescale, x0, lower, upper,
                              rotations, pcs, self.om, **opts)
        finally:
            ctx.set_signal_mask(None)


def _starts(setup, pseudo_symmetry_ops):
    """Euler start values ``(n, 1 + n ops, 3)`` (``_refinement.py:971-980``)."""
    q = setup.quaternions
    if pseudo_symmetry_ops is None:
        return quaternion_to_euler(q)[:, None, :], 0
    ops = np.asarray(getattr(pseudo_symmetry_ops, "data", pseudo_symmetry_ops), dtype=np.float64).reshape(-1, 4)
    # ops.outer(rot): operator applied first in orix's composition order (p * q)
    alt = _quaternion_multiply(ops[None, :, :], q[:, None, :])
    alt = np.where(alt[..., :1] < 0, -alt, alt)
    return quaternion_to_euler(np.concatenate([q[:, None, :], alt], axis=1)), ops.shape[0]


def _finish(setup, res, what, verbose, t0):
    if verbose:
        dt = max(time.time() - t0, 1e-12)
        print(f"Refinement speed: {setup.n / dt:.5f} patterns/s", file=sys.stdout)
    return res


def refine_orientation(signal, xmap, detector, master_pattern, energy=None, navigation_mask=None,
                       signal_mask=None, pseudo_symmetry_ops=None, method="minimize", method_kwargs=None,
                       trust_region=None, initial_step=None, rtol=1e-4, maxeval=None, compute=True,
                       rechunk=True, chunk_kwargs=None, *, context=None, verbose=True, sharded=False, group=None):
    """Refine orientations with fixed projection centres (``signals/ebsd.py:1986-2177``).

    Returns a :class:`RefinementResult` (an orix ``CrystalMap`` needs orix), or with
    ``compute=False`` the raw ``(n, 5 | 6)`` array of the reference (score, evaluations, Euler
    angles[, pseudo-symmetry index]) - already computed, the GPU call is not lazy.

    Patterns of up to 25 600 matched pixels (e.g. 160x160 unmasked) are staged in shared memory for the
    whole search; larger ones (240x240, 480x480) in global memory, L2-resident - the search is bound by the
    float64 projection geometry either way."""
    where, opts, name, kind, kw_shown = _method_plan(method, method_kwargs)
    setup = _Setup(signal, xmap, detector, master_pattern, energy, navigation_mask, signal_mask, context, sharded, group)
    x0, n_ps = _starts(setup, pseudo_symmetry_ops)
    lower, upper = _bounds(x0, trust_region, "ori")
    if verbose:
        print(_info_message(name, trust_region, kw_shown, n_ps, kind, _supports_bounds(method)))
        print(f"Refining {setup.n} orientation(s):", file=sys.stdout)
    t0 = time.time()
    res = setup.run(_lib.REFINE_ORI, x0, lower, upper, None, setup.pcs if setup.unique_pc else None, opts,
                    fixed_dc=not setup.unique_pc, host_plan=None if where == "device" else opts,
                    trust_region_passed=trust_region is not None)
    _finish(setup, res, "orientation", verbose, t0)
    if not compute:
        return res
    return _orientation_result(setup, res, 3, n_ps > 0)


def _orientation_result(setup, res, n_eu, with_ps):
    size = setup.nav_size
    euler = np.zeros((size, 3))
    prop = {"scores": np.zeros(size), "num_evals": np.zeros(size, dtype=np.int32)}
    euler[setup.points] = res[:, 2:5]
    prop["scores"][setup.points] = res[:, 0]
    prop["num_evals"][setup.points] = res[:, 1]
    if with_ps:
        prop["pseudo_symmetry_index"] = np.zeros(size, dtype=np.int32)
        prop["pseudo_symmetry_index"][setup.points] = res[:, -1]
    keep = setup.points
    return RefinementResult(setup.nav_shape, keep, euler[keep], {k: v[keep] for k, v in prop.items()})


def refine_projection_center(signal, xmap, detector, master_pattern, energy=None, navigation_mask=None,
                             signal_mask=None, method="minimize", method_kwargs=None, trust_region=None,
                             initial_step=None, rtol=1e-4, maxeval=None, compute=True, rechunk=True,
                             chunk_kwargs=None, *, context=None, verbose=True, sharded=False, group=None):
    """Refine projection centres with fixed orientations (``signals/ebsd.py:2179-2356``).  Returns
    ``(scores, detector with the refined PCs, num_evals)`` like the reference
    (``_refinement.py:133-200``), or the raw ``(n, 5)`` array with ``compute=False``."""
    where, opts, name, kind, kw_shown = _method_plan(method, method_kwargs)
    setup = _Setup(signal, xmap, detector, master_pattern, energy, navigation_mask, signal_mask, context, sharded, group)
    x0 = setup.pcs[:, None, :].copy()
    lower, upper = _bounds(x0, trust_region, "pc")
    if verbose:
        print(_info_message(name, trust_region, kw_shown, 0, kind, _supports_bounds(method)))
        print(f"Refining {setup.n} projection center(s):", file=sys.stdout)
    t0 = time.time()
    res = setup.run(_lib.REFINE_PC, x0, lower, upper, setup.quaternions, None, opts, fixed_dc=False,
                    host_plan=None if where == "device" else opts, trust_region_passed=trust_region is not None)
    _finish(setup, res, "pc", verbose, t0)
    if not compute:
        return res
    new_det = detector.deepcopy() if hasattr(detector, "deepcopy") else Detector((setup.nrows, setup.ncols))
    new_det.pc = res[:, 2:5].reshape(setup.nav_shape + (3,)) if setup.points.all() else res[:, 2:5]
    return res[:, 0].copy(), new_det, res[:, 1].astype(np.int32)


def refine_orientation_projection_center(signal, xmap, detector, master_pattern, energy=None,
                                         navigation_mask=None, signal_mask=None, pseudo_symmetry_ops=None,
                                         method="minimize", method_kwargs=None, trust_region=None,
                                         initial_step=None, rtol=1e-4, maxeval=None, compute=True, rechunk=True,
                                         chunk_kwargs=None, *, context=None, verbose=True, sharded=False, group=None):
    """Refine orientations and projection centres together (``signals/ebsd.py:2358-2560``).
    Returns ``(RefinementResult, detector with the refined PCs)``, or the raw ``(n, 8 | 9)``
    array with ``compute=False``."""
    where, opts, name, kind, kw_shown = _method_plan(method, method_kwargs)
    setup = _Setup(signal, xmap, detector, master_pattern, energy, navigation_mask, signal_mask, context, sharded, group)
    eu, n_ps = _starts(setup, pseudo_symmetry_ops)
    x0 = np.concatenate([eu, np.repeat(setup.pcs[:, None, :], eu.shape[1], axis=1)], axis=2)
    lower, upper = _bounds(x0, trust_region, "ori_pc")
    if verbose:
        print(_info_message(name, trust_region, kw_shown, n_ps, kind, _supports_bounds(method)))
        print(f"Refining {setup.n} orientation(s) and projection center(s):", file=sys.stdout)
    t0 = time.time()
    res = setup.run(_lib.REFINE_ORI_PC, x0, lower, upper, None, None, opts, fixed_dc=False,
                    host_plan=None if where == "device" else opts, trust_region_passed=trust_region is not None)
    _finish(setup, res, "ori_pc", verbose, t0)
    if not compute:
        return res
    new_det = detector.deepcopy() if hasattr(detector, "deepcopy") else Detector((setup.nrows, setup.ncols))
    new_det.pc = res[:, 5:8].reshape(setup.nav_shape + (3,)) if setup.points.all() else res[:, 5:8]
    return _orientation_result(setup, res, 3, n_ps > 0), new_det
