"""Generate tests/golden/projection_varying_pc.npz by running the REFERENCE's own Numba functions in
place (build container only): _get_direction_cosines_for_varying_pc and
_project_patterns_from_master_pattern_with_varying_pc (signals/util/_master_pattern.py:207-296,
:374-445) for seven rotations, each with its own projection centre, float32 master pattern without
rescaling and uint8 master pattern rescaled to [-1, 1].

  python tests/golden/make_golden_projection_varying_pc.py
"""
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
from oracle import projection_oracle as po  # noqa: E402
from oracle import ref_loader  # noqa: E402

mp = ref_loader.load_master_pattern()
nrows, ncols, n = 14, 18, 7
rng = np.random.default_rng(12)
pcs = np.array([0.42, 0.21, 0.51]) + rng.normal(scale=0.02, size=(n, 3))
om = po.tilted_detector_matrix(70.0)
rot = po.random_rotations(n, seed=8)
gb = np.array([po.gnomonic_bounds(nrows, ncols, *pc) for pc in pcs])
dc = mp._get_direction_cosines_for_varying_pc(gb, np.ascontiguousarray(pcs[:, 2]), nrows, ncols, np.ascontiguousarray(om),
                                              np.ones(nrows * ncols, bool))
mu32, ml32 = po.synthetic_master_pattern(101, seed=5, dtype=np.float32)
mu8, ml8 = po.synthetic_master_pattern(101, seed=6, dtype=np.uint8)
out32 = mp._project_patterns_from_master_pattern_with_varying_pc(rot, dc, mu32, ml32, 101, 101, 50.0, False, 1, 2, np.float32)
out8 = mp._project_patterns_from_master_pattern_with_varying_pc(rot, dc, mu8, ml8, 101, 101, 50.0, True, -1.0, 1.0, np.float32)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "projection_varying_pc.npz"), nrows=nrows, ncols=ncols, pcs=pcs, om=om,
                    rotations=rot, dc=dc, out_f32=out32, out_u8=out8)
o32 = po.project_patterns_varying_pc(rot, pcs, nrows, ncols, om, mu32, ml32)
o8 = po.project_patterns_varying_pc(rot, pcs, nrows, ncols, om, mu8, ml8, rescale=True, out_min=-1.0, out_max=1.0)
print("f32 max diff", np.abs(o32 - out32).max(), "identical", np.mean(o32 == out32))
print("u8  max diff", np.abs(o8 - out8).max(), "identical", np.mean(o8 == out8))
