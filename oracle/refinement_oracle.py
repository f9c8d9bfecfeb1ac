"""CPU restatement (NumPy) of the reference's orientation / projection-centre REFINEMENT
(SURVEY.md section 8f.3).  TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``): imported by
``tests/``, ``tests/gpu_tools`` and golden generators, never by the product path.

Reference (paths relative to /root/reference/src/kikuchipy):
  indexing/_refinement/_solvers.py
    :50-73    _prepare_pattern (astype float32, rescale to [-1, 1] for float32 input, centre, squared norm)
    :79-254   _refine_orientation_solver_scipy (incl. pseudo-symmetry starts, argmax of the scores)
    :257-345  _refine_pc_solver_scipy, :348-470 _refine_orientation_pc_solver_scipy
  indexing/_refinement/_objective_functions.py:36-190  the three objective functions
  indexing/similarity_metrics/_normalized_cross_correlation.py:200-225
            _ncc_single_patterns_1d_float32_exp_centered (float32 arithmetic, returned as float64)
  _utils/numba.py:44-58  rotation_from_euler;  _utils/_gnonomic_bounds.py:23-62 get_gnomonic_bounds
  pattern/_pattern.py:97-139  _rescale_with_min_max, _rescale_without_min_max_1d_float32,
            _zero_mean_sum_square_1d_float32
Third party: the default optimiser is ``scipy.optimize.minimize(method="Nelder-Mead")``
(scipy >= 1.7 per the reference's pyproject; 1.18.1 installed here).  ``nelder_mead`` below restates
``scipy/optimize/_optimize.py::_minimize_neldermead`` (simplex construction, bound handling,
coefficient arithmetic, termination test, ``maxfev`` abort semantics); ``tests/test_refinement.py``
checks it step for step against the installed SciPy on the same objective.

Pinned: ``tests/golden/make_golden_refinement.py`` ran the reference's own solver functions in place
(``oracle/ref_loader.load_refinement``) and stored inputs and outputs in
``tests/golden/refinement.npz``.  The reference's Numba kernels are compiled with ``fastmath=True``
and sum in float32, so objective values agree to float32 rounding (~1e-7), not bit for bit; Nelder-Mead
trajectories are identical wherever no comparison falls inside that noise.
"""

from __future__ import annotations

import numpy as np

from . import projection_oracle as po


def prepare_pattern(pattern, rescale):
    """``_prepare_pattern`` (_solvers.py:50-73)."""
    p = np.asarray(pattern).astype(np.float32)
    if rescale:  # computed in float64 (float32 array / Python float), returned as float32
        imin, imax = p.min(), p.max()
        p = ((p.astype(np.float64) - np.float64(imin)) / float(imax - imin) * 2 + (-1)).astype(np.float32)
    p = p - np.float32(p.mean(dtype=np.float64))
    return p, np.float32(np.sum(np.square(p), dtype=np.float64))


def rotation_from_euler(alpha, beta, gamma):
    """``rotation_from_euler`` (_utils/numba.py:44-58)."""
    sigma = 0.5 * (alpha + gamma)
    delta = 0.5 * (alpha - gamma)
    c, s = np.cos(0.5 * beta), np.sin(0.5 * beta)
    rot = np.array([c * np.cos(sigma), -s * np.cos(delta), -s * np.sin(delta), -c * np.sin(sigma)])
    return -rot if rot[0] < 0 else rot


def gnomonic_bounds(nrows, ncols, pcx, pcy, pcz):
    """``get_gnomonic_bounds`` (_utils/_gnonomic_bounds.py:23-62)."""
    aspect = ncols / nrows
    return np.array([-aspect * (pcx / pcz), aspect * (1 - pcx) / pcz, -(1 - pcy) / pcz, pcy / pcz])


def ncc_exp_centered(exp, sim, exp_squared_norm):
    """``_ncc_single_patterns_1d_float32_exp_centered``: float32 arithmetic, float64 result."""
    sim = sim - np.float32(sim.mean(dtype=np.float64))
    s1 = np.float32(np.sum(exp * sim, dtype=np.float64))
    s2 = np.float32(np.sum(np.square(sim), dtype=np.float64))
    return float(s1 / np.sqrt(np.float32(exp_squared_norm) * s2))


class Problem:
    """Fixed parameters of one refinement (``_RefinementSetup.set_fixed_parameters``,
    _refinement.py:1155-1190): float32 master pattern hemispheres, detector shape, the pixels
    kept by the signal mask (``keep``: flat bool, True = use) and, per mode, fixed direction
    cosines or the detector-to-sample matrix."""

    def __init__(self, master_upper, master_lower, nrows, ncols, keep=None, direction_cosines=None,
                 om_detector_to_sample=None):
        self.mu = np.asarray(master_upper, dtype=np.float32)
        self.ml = np.asarray(master_lower, dtype=np.float32)
        self.npy, self.npx = self.mu.shape
        self.scale = (self.npx - 1) / 2
        self.nrows, self.ncols = int(nrows), int(ncols)
        self.keep = np.ones(nrows * ncols, dtype=bool) if keep is None else np.asarray(keep, dtype=bool).ravel()
        self.dc = direction_cosines  # (n kept pixels, 3) or None
        self.om = om_detector_to_sample

    def dc_from_pc(self, pcx, pcy, pcz):
        return po.direction_cosines_fixed_pc(gnomonic_bounds(self.nrows, self.ncols, pcx, pcy, pcz), pcz,
                                             self.nrows, self.ncols, self.om, self.keep)

    def simulate(self, quaternion, dc):
        return po.project_single_pattern(quaternion, dc, self.mu, self.ml, self.npx, self.npy, self.scale,
                                         False, 0, 1, np.float32)

    # the three objective functions (_objective_functions.py:36-190)
    def objective_ori(self, x, exp, sqnorm, dc):
        return 1 - ncc_exp_centered(exp, self.simulate(rotation_from_euler(*x), dc), sqnorm)

    def objective_pc(self, x, exp, sqnorm, quaternion):
        return 1 - ncc_exp_centered(exp, self.simulate(quaternion, self.dc_from_pc(*x)), sqnorm)

    def objective_ori_pc(self, x, exp, sqnorm):
        return 1 - ncc_exp_centered(exp, self.simulate(rotation_from_euler(*x[:3]), self.dc_from_pc(*x[3:])), sqnorm)


class _MaxFev(Exception):
    pass


def nelder_mead(fun, x0, bounds=None, xatol=1e-4, fatol=1e-4, maxiter=None, maxfev=None, adaptive=False):
    """``_minimize_neldermead`` of SciPy restated.  ``bounds``: ``(n, 2)`` array of (min, max) or
    ``None``.  Returns ``(x, fun, nfev, nit)``."""
    x0 = np.asarray(x0, dtype=np.float64).ravel()
    n = x0.size
    if adaptive:
        dim = float(n)
        rho, chi, psi, sigma = 1, 1 + 2 / dim, 0.75 - 1 / (2 * dim), 1 - 1 / dim
    else:
        rho, chi, psi, sigma = 1, 2, 0.5, 0.5
    lb = ub = None
    if bounds is not None:
        b = np.asarray(bounds, dtype=np.float64)
        lb, ub = b[:, 0], b[:, 1]
        x0 = np.clip(x0, lb, ub)
    sim = np.empty((n + 1, n))
    sim[0] = x0
    for k in range(n):
        y = x0.copy()
        y[k] = (1 + 0.05) * y[k] if y[k] != 0 else 0.00025
        sim[k + 1] = y
    if maxiter is None and maxfev is None:
        maxiter = maxfev = n * 200
    elif maxiter is None:
        maxiter = n * 200 if maxfev == np.inf else np.inf
    elif maxfev is None:
        maxfev = n * 200 if maxiter == np.inf else np.inf
    if bounds is not None:
        sim = np.where(sim > ub, 2 * ub - sim, sim)
        sim = np.clip(sim, lb, ub)
    fsim = np.full(n + 1, np.inf)
    calls = [0]

    def f(x):
        if calls[0] >= maxfev:
            raise _MaxFev
        calls[0] += 1
        return fun(np.copy(x))

    def clip(x):
        return x if bounds is None else np.clip(x, lb, ub)

    try:
        for k in range(n + 1):
            fsim[k] = f(sim[k])
    except _MaxFev:
        pass
    ind = np.argsort(fsim, kind="stable")
    sim, fsim = sim[ind], fsim[ind]
    it = 1
    while calls[0] < maxfev and it < maxiter:
        try:
            if np.max(np.abs(sim[1:] - sim[0])) <= xatol and np.max(np.abs(fsim[0] - fsim[1:])) <= fatol:
                break
            xbar = np.add.reduce(sim[:-1], 0) / n
            xr = clip((1 + rho) * xbar - rho * sim[-1])
            fxr = f(xr)
            shrink = False
            if fxr < fsim[0]:
                xe = clip((1 + rho * chi) * xbar - rho * chi * sim[-1])
                fxe = f(xe)
                if fxe < fxr:
                    sim[-1], fsim[-1] = xe, fxe
                else:
                    sim[-1], fsim[-1] = xr, fxr
            elif fxr < fsim[-2]:
                sim[-1], fsim[-1] = xr, fxr
            elif fxr < fsim[-1]:
                xc = clip((1 + psi * rho) * xbar - psi * rho * sim[-1])
                fxc = f(xc)
                if fxc <= fxr:
                    sim[-1], fsim[-1] = xc, fxc
                else:
                    shrink = True
            else:
                xcc = clip((1 - psi) * xbar + psi * sim[-1])
                fxcc = f(xcc)
                if fxcc < fsim[-1]:
                    sim[-1], fsim[-1] = xcc, fxcc
                else:
                    shrink = True
            if shrink:
                for j in range(1, n + 1):
                    sim[j] = clip(sim[0] + sigma * (sim[j] - sim[0]))
                    fsim[j] = f(sim[j])
            it += 1
        except _MaxFev:
            pass
        ind = np.argsort(fsim, kind="stable")
        sim, fsim = sim[ind], fsim[ind]
    return sim[0], float(np.min(fsim)), calls[0], it


def _solve(fun, starts, bounds, nm_kwargs):
    res = [nelder_mead(fun, x0, None if bounds is None else bounds[i], **nm_kwargs) for i, x0 in enumerate(starts)]
    ncc = [1 - r[1] for r in res]
    best = int(np.argmax(ncc))
    out = [ncc[best], res[best][2], *res[best][0]]
    if len(starts) > 1:
        out.append(best)
    return out


def refine_orientation(problem, patterns, eulers, rescale, bounds=None, pcs=None, **nm_kwargs):
    """``_refine_orientation_chunk_scipy`` (_refinement.py:437-500) with Nelder-Mead: ``patterns``
    ``(n, kept pixels)``, ``eulers`` ``(n, starts, 3)``, ``bounds`` ``None`` or
    ``(n, starts, 3, 2)``, ``pcs`` ``None`` (fixed direction cosines) or ``(n, 3)``.  Rows of the
    result: score, number of evaluations, phi1, Phi, phi2[, index of the best start]."""
    out = []
    for i in range(len(patterns)):
        exp, sq = prepare_pattern(patterns[i], rescale)
        dc = problem.dc if pcs is None else problem.dc_from_pc(*pcs[i])
        out.append(_solve(lambda x: problem.objective_ori(x, exp, sq, dc), eulers[i],
                          None if bounds is None else bounds[i], nm_kwargs))
    return np.array(out, dtype=np.float64)


def refine_pc(problem, patterns, quaternions, pcs, rescale, bounds=None, **nm_kwargs):
    """``_refine_pc_chunk_scipy``: rows score, number of evaluations, PCx, PCy, PCz."""
    out = []
    for i in range(len(patterns)):
        exp, sq = prepare_pattern(patterns[i], rescale)
        out.append(_solve(lambda x: problem.objective_pc(x, exp, sq, quaternions[i]), [pcs[i]],
                          None if bounds is None else [bounds[i]], nm_kwargs))
    return np.array(out, dtype=np.float64)


def refine_orientation_pc(problem, patterns, euler_pcs, rescale, bounds=None, **nm_kwargs):
    """``_refine_orientation_pc_chunk_scipy``: ``euler_pcs`` ``(n, starts, 6)``; rows score, number
    of evaluations, phi1, Phi, phi2, PCx, PCy, PCz[, index of the best start]."""
    out = []
    for i in range(len(patterns)):
        exp, sq = prepare_pattern(patterns[i], rescale)
        out.append(_solve(lambda x: problem.objective_ori_pc(x, exp, sq), euler_pcs[i],
                          None if bounds is None else bounds[i], nm_kwargs))
    return np.array(out, dtype=np.float64)


# ---- synthetic refinement problems (tests, golden generator, timing tools) ---------------------

def euler_to_quaternion_batch(eulers):
    return np.array([rotation_from_euler(*e) for e in np.asarray(eulers).reshape(-1, 3)])


def synthetic_case(n=8, nrows=24, ncols=32, mp_size=201, seed=0, dtype=np.uint8, noise=0.05, perturb_deg=1.0,
                   pc=(0.42, 0.21, 0.51), pc_spread=0.0, circular_mask=False):
    """Patterns projected from a synthetic master pattern at known orientations (+ noise), and
    start values a little away from them, as dictionary indexing would deliver."""
    rng = np.random.default_rng(seed)
    mu, ml = po.synthetic_master_pattern(mp_size, seed=seed + 5)
    om = po.tilted_detector_matrix(70.0 - 0.0)
    keep = np.ones(nrows * ncols, dtype=bool)
    if circular_mask:
        r, c = np.mgrid[:nrows, :ncols]
        keep = (np.sqrt((r - nrows // 2) ** 2 + (c - ncols // 2) ** 2) <= max(nrows // 2, ncols // 2) * 0.9).ravel()
    prob = Problem(mu, ml, nrows, ncols, keep=keep, om_detector_to_sample=om)
    true_eu = np.stack([rng.uniform(0.2, 6.0, n), rng.uniform(0.2, 2.9, n), rng.uniform(0.2, 6.0, n)], axis=1)
    pcs = np.asarray(pc)[None, :] + rng.normal(scale=pc_spread, size=(n, 3)) if pc_spread else np.tile(pc, (n, 1))
    full = Problem(mu, ml, nrows, ncols, om_detector_to_sample=om)
    pats = np.empty((n, nrows * ncols), dtype=dtype)
    for i in range(n):
        sim = full.simulate(rotation_from_euler(*true_eu[i]), full.dc_from_pc(*pcs[i])).astype(np.float64)
        sim = sim + rng.normal(scale=noise * sim.std(), size=sim.shape)
        sim = (sim - sim.min()) / (sim.max() - sim.min())
        pats[i] = np.round(sim * 255).astype(np.uint8) if np.dtype(dtype) == np.uint8 else sim.astype(dtype)
    start_eu = true_eu + np.deg2rad(rng.uniform(-perturb_deg, perturb_deg, (n, 3)))
    prob.dc = prob.dc_from_pc(*pc)
    return {"problem": prob, "patterns": pats, "true_eulers": true_eu, "start_eulers": start_eu, "pcs": pcs,
            "pc": np.asarray(pc, dtype=np.float64), "om": om, "keep": keep, "mu": mu, "ml": ml}
