"""Synthetic inputs for benchmarks, diagnostics and tests of the dictionary-generation step: a
smooth two-hemisphere master pattern, random unit quaternions and a tilted-detector matrix.
Plain NumPy input generators - not part of the indexing path."""

from __future__ import annotations

import numpy as np


def synthetic_master_pattern(n: int = 401, seed: int = 5, dtype=np.float32):
    """A smooth synthetic (upper, lower) master pattern pair of shape ``(n, n)``: a sum of a few
    random low-frequency cosines, rescaled to [0, 1] (float) or [0, 255] (uint8)."""
    rng = np.random.default_rng(seed)
    y, x = np.meshgrid(np.linspace(-1, 1, n), np.linspace(-1, 1, n), indexing="ij")
    out = []
    for _ in range(2):
        m = np.zeros((n, n))
        for _ in range(24):
            fx, fy = rng.uniform(-14, 14, 2)
            m += rng.uniform(0.3, 1.0) * np.cos(fx * x + fy * y + rng.uniform(0, 2 * np.pi))
        m = (m - m.min()) / (m.max() - m.min())
        out.append((m * 255).astype(np.uint8) if np.dtype(dtype) == np.uint8 else m.astype(dtype))
    return out[0], out[1]


def random_rotations(n: int, seed: int = 4) -> np.ndarray:
    """``n`` random unit quaternions ``(a, b, c, d)``, float64."""
    rng = np.random.default_rng(seed)
    q = rng.normal(size=(n, 4))
    return q / np.linalg.norm(q, axis=1, keepdims=True)


def tilted_detector_matrix(tilt_deg: float = 70.0) -> np.ndarray:
    """A detector-to-sample orientation matrix: rotation by ``tilt_deg`` about x (a stand-in for
    ``(~EBSDDetector.sample_to_detector).to_matrix()``)."""
    t = np.deg2rad(tilt_deg)
    return np.array([[1, 0, 0], [0, np.cos(t), -np.sin(t)], [0, np.sin(t), np.cos(t)]], dtype=np.float64)
