"""How close the oracle's refinement is to the reference's own solver over many patterns (build
container only: runs the reference in place).  Writes profiles/r1_refine_oracle_vs_reference.json.

  python tests/golden/compare_refinement_with_reference.py [n_patterns]
"""
import json
import os
import sys
import time

import numpy as np
import scipy.optimize as so

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402
from oracle import refinement_oracle as ro  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
S, _ = ref_loader.load_refinement()
case = ro.synthetic_case(n=n, seed=21, nrows=40, ncols=40, mp_size=401, noise=0.05, perturb_deg=1.5, circular_mask=True)
prob, keep = case["problem"], case["keep"]
pats = case["patterns"][:, keep]
fixed = (prob.mu, prob.ml, prob.npx, prob.npy, prob.scale)
out = {}
for name, tr in (("unbounded", None), ("trust_region_2deg", np.deg2rad([2.0, 2.0, 2.0]))):
    t0 = time.time()
    ref = []
    for i in range(n):
        x0 = case["start_eulers"][i][None]
        b = np.zeros((1, 3, 2)) if tr is None else np.stack([x0 - tr, x0 + tr], axis=-1)
        ref.append(S._refine_orientation_solver_scipy(pattern=pats[i].copy(), rotation=x0, bounds=b, signal_mask=keep, rescale=False,
                                                      method=so.minimize, method_kwargs=dict(method="Nelder-Mead"),
                                                      trust_region_passed=tr is not None, fixed_parameters=fixed, direction_cosines=prob.dc))
    t_ref = time.time() - t0
    ref = np.array(ref)
    x0 = case["start_eulers"][:, None, :]
    t0 = time.time()
    got = ro.refine_orientation(prob, pats, x0, False, bounds=None if tr is None else np.stack([x0 - tr, x0 + tr], axis=-1))
    t_orc = time.time() - t0
    q_ref, q_got, q_true = (ro.euler_to_quaternion_batch(e) for e in (ref[:, 2:5], got[:, 2:5], case["true_eulers"]))
    mis = np.degrees(2 * np.arccos(np.clip(np.abs(np.sum(q_ref * q_got, axis=1)), 0, 1)))
    mis_true_ref = np.degrees(2 * np.arccos(np.clip(np.abs(np.sum(q_ref * q_true, axis=1)), 0, 1)))
    mis_true_got = np.degrees(2 * np.arccos(np.clip(np.abs(np.sum(q_got * q_true, axis=1)), 0, 1)))
    out[name] = {
        "patterns": n, "detector": [40, 40], "kept_pixels": int(keep.sum()),
        "max_abs_dscore": float(np.abs(ref[:, 0] - got[:, 0]).max()), "median_abs_dscore": float(np.median(np.abs(ref[:, 0] - got[:, 0]))),
        "identical_evaluation_counts": float(np.mean(ref[:, 1] == got[:, 1])),
        "mean_evaluations": {"reference": float(ref[:, 1].mean()), "oracle": float(got[:, 1].mean())},
        "misorientation_between_results_deg": {"median": float(np.median(mis)), "max": float(mis.max())},
        "median_misorientation_to_truth_deg": {"reference": float(np.median(mis_true_ref)), "oracle": float(np.median(mis_true_got))},
        "patterns_per_s": {"reference_numba_1core": round(n / t_ref, 1), "oracle_numpy_1core": round(n / t_orc, 1)},
    }
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
with open(os.path.join(ROOT, "profiles", "r1_refine_oracle_vs_reference.json"), "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps(out, indent=1))
