"""Generate tests/golden/projection.npz by running the REFERENCE's own Numba projection functions
in place (build container only: needs /root/reference and numba).

  python tests/golden/make_golden_projection.py

Stored: inputs (master patterns, direction cosines, rotations, geometry) and the reference's
outputs of _get_direction_cosines_for_fixed_pc and
_project_patterns_from_master_pattern_with_fixed_pc (signals/util/_master_pattern.py:133,299)
for (a) a float32 master pattern without rescaling (what EBSDMasterPattern.get_patterns does when
dtype_out equals the master pattern's dtype) and (b) a uint8 master pattern rescaled to the
float32 range [-1, 1] (dtype differs -> rescale, signals/ebsd_master_pattern.py:222-233).
"""
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
from oracle import projection_oracle as po  # noqa: E402
from oracle import ref_loader  # noqa: E402

mp = ref_loader.load_master_pattern()
nrows, ncols = 16, 20
gb = np.array([-0.95, 0.81, -0.62, 0.77])
pcz = 0.52
om = po.tilted_detector_matrix(70.0)
dc = mp._get_direction_cosines_for_fixed_pc(gb, pcz, nrows, ncols, np.ascontiguousarray(om), np.ones(nrows * ncols, bool))
mask = np.ones(nrows * ncols, bool); mask[::7] = False
dc_masked = mp._get_direction_cosines_for_fixed_pc(gb, pcz, nrows, ncols, np.ascontiguousarray(om), mask)
# a few hand-made directions: both poles, the x/y diagonal, axis-aligned vectors
extra = np.array([[0, 0, 1], [0, 0, -1], [1, 1, 0.2], [-1, 1, -0.2], [1, 0, 0], [0, -1, 0], [1e-9, 1e-9, 1]], dtype=np.float64)
extra /= np.linalg.norm(extra, axis=1, keepdims=True)
dc_all = np.ascontiguousarray(np.vstack([dc, extra]))
rot = po.random_rotations(5, seed=4)
rot = np.vstack([rot, [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [np.sqrt(0.5), 0, np.sqrt(0.5), 0]]])
n = 101
mu32, ml32 = po.synthetic_master_pattern(n, seed=5, dtype=np.float32)
mu8, ml8 = po.synthetic_master_pattern(n, seed=6, dtype=np.uint8)
scale = float((n - 1) / 2)
out_f32 = mp._project_patterns_from_master_pattern_with_fixed_pc(rot, dc_all, mu32, ml32, n, n, scale, False, 1, 2, np.float32)
out_u8 = mp._project_patterns_from_master_pattern_with_fixed_pc(rot, dc_all, mu8, ml8, n, n, scale, True, -1.0, 1.0, np.float32)
np.savez_compressed(
    os.path.join(ROOT, "tests", "golden", "projection.npz"),
    gnomonic_bounds=gb, pcz=pcz, nrows=nrows, ncols=ncols, om=om, dc=dc, dc_mask=mask, dc_masked=dc_masked,
    dc_all=dc_all, rotations=rot, mu32=mu32, ml32=ml32, mu8=mu8, ml8=ml8, scale=scale, out_f32=out_f32, out_u8=out_u8,
)
# the restatement must agree with what was just generated
o1 = po.project_patterns(rot, dc_all, mu32, ml32)
o2 = po.project_patterns(rot, dc_all, mu8, ml8, rescale=True, out_min=-1.0, out_max=1.0)
d = po.direction_cosines_fixed_pc(gb, pcz, nrows, ncols, om)
print("dc max diff", np.abs(d - dc).max(), "masked", np.abs(po.direction_cosines_fixed_pc(gb, pcz, nrows, ncols, om, mask) - dc_masked).max())
print("f32 max diff", np.abs(o1 - out_f32).max(), "identical", np.mean(o1 == out_f32))
print("u8  max diff", np.abs(o2 - out_u8).max(), "identical", np.mean(o2 == out_u8))
print("lower-hemisphere fraction", np.mean([(po.rotate_vector(r, dc_all)[:, 2] < 0).mean() for r in rot]))
