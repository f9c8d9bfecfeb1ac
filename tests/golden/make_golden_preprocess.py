"""Generate tests/golden/preprocess.npz by running the REFERENCE's own background-removal and
neighbour-averaging functions in place (build container only: needs /root/reference, numba, scipy).

  python tests/golden/make_golden_preprocess.py

Inputs: the nine 60x60 uint8 patterns of kp.data.nickel_ebsd_small (as a 3x3 map), random uint8 /
uint16 / float32 patterns.  Outputs of pattern/_pattern.py's _remove_static_background_subtract /
_divide (with and without scale_bg), _remove_dynamic_background (frequency and spatial domain,
subtract and divide, default and explicit std), pattern/chunk.py's _average_neighbour_patterns
(circular 3x3, rectangular 3x3, Gaussian 5x5 windows), and the Window arrays themselves.
"""
import os
import sys

import numpy as np
from scipy.ndimage import correlate, gaussian_filter

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
from oracle import preprocess_oracle as pp  # noqa: E402
from oracle import ref_loader  # noqa: E402

R = ref_loader.load_preprocessing()
P, W = R.pattern, R.Window
rng = np.random.default_rng(0)
ni = ref_loader.nickel_ebsd_small()  # (3, 3, 60, 60) uint8
out = {"ni": ni}
sets = {
    "ni": ni.reshape(9, 60, 60),
    "u8": rng.integers(0, 256, (6, 24, 32), dtype=np.uint8),
    "u16": rng.integers(0, 65536, (4, 20, 20)).astype(np.uint16),
    "f32": rng.random((4, 20, 28), dtype=np.float32),
}
report = {}


def same(name, got, want):
    report[name] = (int(np.sum(got != want)), want.size)


for key, pats in sets.items():
    dt = pats.dtype.type
    omin, omax = pp.DTYPE_RANGE[pats.dtype]
    bg = pats.mean(axis=0).astype(pats.dtype) if key != "f32" else (pats.mean(axis=0) + np.float32(0.05)).astype(np.float32)
    if key in ("u8", "u16"):
        bg = np.maximum(bg, 1).astype(pats.dtype)
    out[f"{key}_patterns"], out[f"{key}_static_bg"] = pats, bg
    for op, fn in (("subtract", P._remove_static_background_subtract), ("divide", P._remove_static_background_divide)):
        for scale in (False, True):
            ref = np.stack([fn(p, bg.astype(np.float32), dt, omin, omax, scale) for p in pats])
            name = f"{key}_static_{op}_{'scaled' if scale else 'plain'}"
            out[name] = ref
            same(name, pp.remove_static_background(pats, bg, op, scale), ref)
    for dom in ("frequency", "spatial"):
        for op in ("subtract", "divide"):
            for std in (None, 2.5):
                s = pats.shape[2] / 8 if std is None else std
                if dom == "frequency":
                    kw = {}
                    (kw["fft_shape"], kw["window_shape"], kw["transfer_function"], kw["offset_before_fft"],
                     kw["offset_after_ifft"]) = P._dynamic_background_frequency_space_setup(pats.shape[1:], s, 4.0)
                    ff = R.fft_barnes._fft_filter
                else:
                    kw = {"sigma": s, "truncate": 4.0}
                    ff = gaussian_filter
                ref = np.stack([P._remove_dynamic_background(p, ff, op, dt, omin, omax, **kw) for p in pats])
                name = f"{key}_dynamic_{dom}_{op}_{'default' if std is None else 'std2p5'}"
                out[name] = ref
                same(name, pp.remove_dynamic_background(pats, op, dom, std, 4.0), ref)

# neighbour averaging on the 3 x 3 nickel map and a 4 x 5 random map
maps = {"ni": ni, "u8map": rng.integers(0, 256, (4, 5, 12, 12), dtype=np.uint8), "f32map": rng.random((3, 4, 10, 10), dtype=np.float32)}
windows = {"circular3": W("circular", (3, 3)), "rect3": W("rectangular", (3, 3)), "gauss5": W("gaussian", (5, 5), std=1.0),
           "circular5": W("circular", (5, 5))}
for wname, w in windows.items():
    out[f"window_{wname}"] = np.asarray(w)
for mname, m in maps.items():
    out[f"{mname}_map"] = m
    omin, omax = pp.DTYPE_RANGE[m.dtype]
    for wname, w in windows.items():
        warr = np.asarray(w)
        sums = correlate(np.ones(m.shape[:2], dtype=int), weights=warr, mode="constant")
        ref = R.chunk._average_neighbour_patterns(m, sums[..., None, None], warr.reshape(warr.shape + (1, 1)), m.dtype.type, omin, omax)
        name = f"{mname}_average_{wname}"
        out[name] = ref
        same(name, pp.average_neighbour_patterns(m, warr), ref)
for wname in ("circular3", "circular5"):
    n = int(wname[-1])
    report[f"window_{wname}"] = (int(np.sum(pp.circular_window((n, n)) != out[f"window_{wname}"])), n * n)

np.savez_compressed(os.path.join(ROOT, "tests", "golden", "preprocess.npz"), **out)
bad = {k: v for k, v in report.items() if v[0]}
print(f"{len(report)} outputs; oracle differs in: {bad if bad else 'none'}")
