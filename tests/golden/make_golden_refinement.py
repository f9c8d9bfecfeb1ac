"""Generate tests/golden/refinement.npz by running the REFERENCE's own refinement solvers in place
(build container only: needs /root/reference, numba and scipy).

  python tests/golden/make_golden_refinement.py

Stored: the inputs of small synthetic refinement problems (``oracle.refinement_oracle.synthetic_case``:
patterns, start values, projection centres, master pattern seed) and what the reference returns for
them - ``_prepare_pattern``, the three objective functions at the start values
(indexing/_refinement/_objective_functions.py:36-190) and the three ``*_solver_scipy`` functions with
``scipy.optimize.minimize(method="Nelder-Mead")`` (indexing/_refinement/_solvers.py:79-470): without
and with a trust region, with pseudo-symmetry starts, with one PC per pattern, with float32 patterns
(rescaled) - plus ``_sample_to_detector_matrix`` (detectors/_ebsd_detector.py:100-149).
"""
import ast
import os
import sys

import numpy as np
import scipy.optimize as so

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402
from oracle import refinement_oracle as ro  # noqa: E402

S, MP = ref_loader.load_refinement()
OF = sys.modules["kikuchipy.indexing._refinement._objective_functions"]
NM = dict(method="Nelder-Mead")
out = {}


def fixed(prob, pc_modes=False):
    base = (prob.mu, prob.ml, prob.npx, prob.npy, prob.scale)
    return base + (prob.keep, prob.nrows, prob.ncols, np.ascontiguousarray(prob.om)) if pc_modes else base


def bounds_for(x0, tr):
    lo, hi = x0 - tr, x0 + tr
    return np.stack([lo, hi], axis=-1)


# ---- case A: uint8 patterns, circular mask, fixed PC -------------------------------------------
A = ro.synthetic_case(n=8, seed=1, circular_mask=True)
prob, keep = A["problem"], A["keep"]
pats = A["patterns"][:, keep]
out["A_patterns"], out["A_keep"], out["A_start"], out["A_true"], out["A_pc"] = A["patterns"], keep, A["start_eulers"], A["true_eulers"], A["pc"]
out["A_dc"] = prob.dc
prep = [S._prepare_pattern(p.copy(), False) for p in pats]
out["A_prepared"] = np.array([p[0] for p in prep])
out["A_sqnorm"] = np.array([p[1] for p in prep], dtype=np.float32)
out["A_objective_start"] = np.array([OF._refine_orientation_objective_function(A["start_eulers"][i], prep[i][0], prob.dc, *fixed(prob), prep[i][1]) for i in range(8)])
res, res_b, res_ps = [], [], []
tr = np.deg2rad([2.0, 2.0, 2.0])
for i in range(8):
    x0 = A["start_eulers"][i][None]
    res.append(S._refine_orientation_solver_scipy(pattern=pats[i].copy(), rotation=x0, bounds=np.zeros((1, 3, 2)), signal_mask=keep, rescale=False,
                                                  method=so.minimize, method_kwargs=dict(NM), trust_region_passed=False, fixed_parameters=fixed(prob), direction_cosines=prob.dc))
    res_b.append(S._refine_orientation_solver_scipy(pattern=pats[i].copy(), rotation=x0, bounds=bounds_for(x0, tr), signal_mask=keep, rescale=False,
                                                    method=so.minimize, method_kwargs=dict(NM), trust_region_passed=True, fixed_parameters=fixed(prob), direction_cosines=prob.dc))
    # pseudo-symmetry: two more starts, one far off (a wrong variant) and one on the other side
    x3 = np.stack([A["start_eulers"][i], A["start_eulers"][i] + [0.5, -0.3, 0.2], A["true_eulers"][i] + np.deg2rad([0.3, -0.2, 0.4])])
    res_ps.append(S._refine_orientation_solver_scipy(pattern=pats[i].copy(), rotation=x3, bounds=np.zeros((3, 3, 2)), signal_mask=keep, rescale=False,
                                                     method=so.minimize, method_kwargs=dict(NM), trust_region_passed=False, fixed_parameters=fixed(prob), direction_cosines=prob.dc,
                                                     n_pseudo_symmetry_ops=2))
    out.setdefault("A_ps_starts", []).append(x3)
out["A_ps_starts"] = np.array(out["A_ps_starts"])
out["A_result"], out["A_result_bounded"], out["A_result_ps"] = np.array(res), np.array(res_b), np.array(res_ps)
out["A_trust_region_deg"] = np.array([2.0, 2.0, 2.0])

# ---- case B: one PC per pattern (orientation), PC refinement, orientation + PC -----------------
B = ro.synthetic_case(n=4, seed=2, pc_spread=0.01, nrows=20, ncols=20)
prob, keep = B["problem"], B["keep"]
pats = B["patterns"][:, keep]
out["B_patterns"], out["B_start"], out["B_true"], out["B_pcs"], out["B_om"] = B["patterns"], B["start_eulers"], B["true_eulers"], B["pcs"], B["om"]
quats = ro.euler_to_quaternion_batch(B["true_eulers"])
pc_start = B["pcs"] + np.array([0.004, -0.003, 0.005])
out["B_quats"], out["B_pc_start"] = quats, pc_start
r_vpc, r_pc, r_pcb, r_opc = [], [], [], []
tr_pc = np.array([0.02, 0.02, 0.02])
obj_pc, obj_opc = [], []
for i in range(4):
    x0 = B["start_eulers"][i][None]
    r_vpc.append(S._refine_orientation_solver_scipy(pattern=pats[i].copy(), rotation=x0, bounds=np.zeros((1, 3, 2)), signal_mask=keep, rescale=False,
                                                    method=so.minimize, method_kwargs=dict(NM), trust_region_passed=False, fixed_parameters=fixed(prob),
                                                    pcx=float(B["pcs"][i, 0]), pcy=float(B["pcs"][i, 1]), pcz=float(B["pcs"][i, 2]), nrows=prob.nrows, ncols=prob.ncols,
                                                    om_detector_to_sample=np.ascontiguousarray(prob.om)))
    r_pc.append(S._refine_pc_solver_scipy(pattern=pats[i].copy(), rotation=quats[i], pc=pc_start[i], bounds=np.zeros((3, 2)), rescale=False, method=so.minimize,
                                          method_kwargs=dict(NM), fixed_parameters=fixed(prob, True), trust_region_passed=False))
    r_pcb.append(S._refine_pc_solver_scipy(pattern=pats[i].copy(), rotation=quats[i], pc=pc_start[i], bounds=bounds_for(pc_start[i], tr_pc), rescale=False,
                                           method=so.minimize, method_kwargs=dict(NM), fixed_parameters=fixed(prob, True), trust_region_passed=True))
    x6 = np.concatenate([B["start_eulers"][i], pc_start[i]])[None]
    r_opc.append(S._refine_orientation_pc_solver_scipy(pattern=pats[i].copy(), rot_pc=x6, bounds=np.zeros((1, 6, 2)), rescale=False, method=so.minimize,
                                                       method_kwargs=dict(NM), fixed_parameters=fixed(prob, True), trust_region_passed=False))
    e, sq = S._prepare_pattern(pats[i].copy(), False)
    obj_pc.append(OF._refine_pc_objective_function(pc_start[i], e, quats[i], *fixed(prob, True), sq))
    obj_opc.append(OF._refine_orientation_pc_objective_function(x6[0], e, *fixed(prob, True), sq))
out["B_result_varying_pc"], out["B_result_pc"], out["B_result_pc_bounded"], out["B_result_ori_pc"] = map(np.array, (r_vpc, r_pc, r_pcb, r_opc))
out["B_pc_trust_region"] = tr_pc
out["B_objective_pc_start"], out["B_objective_ori_pc_start"] = np.array(obj_pc), np.array(obj_opc)

# ---- case C: float32 patterns (rescaled to [-1, 1] by _prepare_pattern) ------------------------
C = ro.synthetic_case(n=4, seed=3, dtype=np.float32, nrows=18, ncols=26)
prob = C["problem"]
out["C_patterns"], out["C_start"], out["C_pc"], out["C_dc"] = C["patterns"], C["start_eulers"], C["pc"], prob.dc
prep = [S._prepare_pattern(p.copy(), True) for p in C["patterns"]]
out["C_prepared"] = np.array([p[0] for p in prep])
out["C_sqnorm"] = np.array([p[1] for p in prep], dtype=np.float32)
out["C_result"] = np.array([
    S._refine_orientation_solver_scipy(pattern=C["patterns"][i].copy(), rotation=C["start_eulers"][i][None], bounds=np.zeros((1, 3, 2)), signal_mask=C["keep"],
                                       rescale=True, method=so.minimize, method_kwargs=dict(NM, options=dict(xatol=1e-5, fatol=1e-6, maxfev=150)),
                                       trust_region_passed=False, fixed_parameters=fixed(prob), direction_cosines=prob.dc) for i in range(4)])
out["C_options"] = np.array([1e-5, 1e-6, 150])

# ---- detector matrix -------------------------------------------------------------------------
path = os.path.join(ref_loader._SRC, "detectors", "_ebsd_detector.py")
fn = [n for n in ast.parse(open(path).read()).body if isinstance(n, ast.FunctionDef) and n.name == "_sample_to_detector_matrix"]
import numba as nb  # noqa: E402

ns = {"np": np, "nb": nb}
exec(compile(ast.Module(body=fn, type_ignores=[]), path, "exec"), ns)
angles = np.array([[70, 0, 0, 0], [70, 5, 3, 1], [65.5, -10, 12, -3]], dtype=np.float64)
out["det_angles_deg"] = angles
out["det_matrices"] = np.array([ns["_sample_to_detector_matrix"](*np.deg2rad(a)) for a in angles])

np.savez_compressed(os.path.join(ROOT, "tests", "golden", "refinement.npz"), **out)
for k in ("A_result", "A_result_bounded", "A_result_ps", "B_result_varying_pc", "B_result_pc", "B_result_pc_bounded", "B_result_ori_pc", "C_result"):
    print(k, np.array2string(out[k][:2], precision=6, max_line_width=200))
