"""cProfile of the host side of one public-API call (pinned host inputs; GEN=1: generated dictionary)."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import kikuchipy_b200 as kb
from kikuchipy_b200 import synthetic as po

ctx = kb.default_context(0)
exp = ctx.pinned_empty((10000, 60, 60), np.uint8); exp[:] = np.random.default_rng(1).integers(0, 256, exp.shape, dtype=np.uint8)
if os.environ.get("GEN", "0") == "1":
    mu, ml = po.synthetic_master_pattern(1001, seed=5)
    dc = kb.direction_cosines([-0.9, 0.85, -0.7, 0.95], 0.5, 60, 60, po.tilted_detector_matrix(70.0))
    dic = kb.get_patterns(mu, ml, po.random_rotations(100000, seed=4), direction_cosines=dc, detector_shape=(60, 60), context=ctx)
else:
    dic = ctx.pinned_empty((100000, 60, 60), np.float32); dic[:] = np.random.default_rng(2).random(dic.shape, dtype=np.float32)
call = lambda: kb.dictionary_indexing(exp, dic, metric="ncc", keep_n=20, verbose=False, context=ctx)
for _ in range(3):
    call()
t0 = time.perf_counter()
for _ in range(5):
    call()
print(f"wall per call: {(time.perf_counter() - t0) / 5 * 1e3:.3f} ms; library total_ms: {ctx.timings()['total_ms']:.3f}")
pr = cProfile.Profile(); pr.enable()
for _ in range(5):
    call()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
