"""The certificate's worst-case bound (kdi_certificate_bound, include/kdi.h KDI_OPT_CERT_STRICT) against a
software model of what the tensor-core pass computes - CPU only.

The model restates, in exact integer-valued float64 arithmetic, the rule measured on the B200 by
tools/probes/mma_accumulate_probe.py (profiles/r2_mma_accumulate_probe.txt): operands = RN-even fp16 of
256 x the float32 row; per 16-deep step the 16 exact products and the float32 accumulator are aligned to
the largest exponent among them, truncated towards zero at 2 guard bits below that exponent's float32 ulp,
summed, and the sum is truncated towards zero to float32.  The bound must cover |model - exact| / (|e| |d|)
for ADVERSARIAL rows - every element rounded by almost the full half ulp, the rounding errors parallel to the
other operand, all products of one sign so that every truncation pulls the same way - as well as for random
ones; the adversarial case also shows how much of the bound is reachable at all."""
import json
import math
import os

import numpy as np
import pytest

from kikuchipy_b200 import _lib


def _trunc_to(x: float, quantum_log2: int) -> float:
    q = math.ldexp(1.0, quantum_log2)
    return math.trunc(x / q) * q


def _tc_model(e16: np.ndarray, d16: np.ndarray) -> float:
    """float32-accumulated dot product of two fp16 operand rows by the measured rule."""
    acc = 0.0
    for k0 in range(0, e16.size, 16):
        prods = [float(a) * float(b) for a, b in zip(e16[k0:k0 + 16], d16[k0:k0 + 16])]  # exact in float64
        big = max([abs(acc)] + [abs(p) for p in prods])
        if big == 0.0:
            continue
        quantum = math.frexp(big)[1] - 1 - 23 - 2  # 2 guard bits below the float32 ulp of the largest addend
        total = _trunc_to(acc, quantum) + sum(_trunc_to(p, quantum) for p in prods)
        if total != 0.0:
            total = _trunc_to(total, math.frexp(total)[1] - 1 - 23)  # back to float32, towards zero
        acc = total
    return acc


def _operands(row32: np.ndarray, kp: int) -> np.ndarray:
    out = np.zeros(kp, dtype=np.float16)
    out[:row32.size] = (row32.astype(np.float32) * np.float32(256.0)).astype(np.float16)  # cvt.rn.f16.f32
    return out


def _ratio(e32, d32, kp):
    exact = float(np.dot(e32.astype(np.float64), d32.astype(np.float64)))
    approx = _tc_model(_operands(e32, kp), _operands(d32, kp)) / 65536.0
    scale = float(np.linalg.norm(e32.astype(np.float64)) * np.linalg.norm(d32.astype(np.float64)))
    return abs(approx - exact) / scale


@pytest.mark.parametrize("s_eff", [900, 3600])
def test_bound_covers_the_modelled_tensor_core_error(s_eff):
    lib = _lib.load()
    kp = (s_eff + 63) // 64 * 64
    bound = lib.kdi_certificate_bound(0, s_eff)
    u = 2.0 ** -11
    rng = np.random.default_rng(s_eff)
    # adversarial: every element sits just below the midpoint between two fp16 values right above a power of
    # two (relative rounding error ~ -u), all elements and products positive: rounding errors parallel to the
    # other operand, every truncation downwards
    m = math.floor(math.log2(1.0 / math.sqrt(s_eff)))
    val = np.float32(math.ldexp(1.0, m) * (1.0 + 0.998 * u))
    e_adv = np.full(s_eff, val, dtype=np.float32)
    d_adv = np.full(s_eff, val, dtype=np.float32)
    r_adv = _ratio(e_adv, d_adv, kp)
    assert r_adv <= bound, (r_adv, bound)
    assert r_adv >= 0.6 * bound, (r_adv, bound)  # the bound is not loose by more than its safety factors
    # mixed signs with the errors still aligned (e rounds down in magnitude, d has e's signs)
    sgn = rng.choice([-1.0, 1.0], s_eff).astype(np.float32)
    assert _ratio(e_adv * sgn, d_adv * sgn, kp) <= bound
    # random rows of the benchmark's kind, centred (NCC) and uncentred (NDP)
    worst = 0.0
    for _ in range(6):
        x = rng.integers(0, 256, s_eff).astype(np.float64)
        y = rng.random(s_eff)
        for centre in (True, False):
            a = x - x.mean() if centre else x
            b = y - y.mean() if centre else y
            a32 = (a / np.linalg.norm(a)).astype(np.float32)
            b32 = (b / np.linalg.norm(b)).astype(np.float32)
            worst = max(worst, _ratio(a32, b32, kp))
    assert worst <= 0.2 * bound, (worst, bound)


def test_model_reproduces_the_hardware_probe():
    """tests/golden/mma_accumulate_probe_b200.jsonl: what a B200 returned for the 324 operand pairs of
    tools/probes/mma_accumulate_probe.py (probes A and B; session 69, through kdi_debug_gemm16).  The probe's
    rows are rebuilt here (NDP prepare = x / float32(norm), operands = fp16 of 256 x), which the recorded exact
    sums confirm, and the model must return the tensor core's float32 result in every case, to the last bit."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mma_accumulate_probe_b200.jsonl")
    rec = [json.loads(line) for line in open(path)]
    S, exps = 1024, list(range(6, 15))

    def prepare(x):
        n = np.sqrt(np.sum(x.astype(np.float64) ** 2, axis=1)).astype(np.float32)
        return (x / n[:, None]).astype(np.float32)

    k = 0
    for label, pos in (("A_one_step", list(range(1, 16))), ("B_across_steps", [16 * m for m in range(1, 64)])):
        e = np.zeros((2 * len(exps), S), dtype=np.float32)
        d = np.zeros((len(exps), S), dtype=np.float32)
        for r, a in enumerate(exps):
            for sgn in (0, 1):
                e[2 * r + sgn, 0] = 1.0
                e[2 * r + sgn, pos] = (1.0 if sgn == 0 else -1.0) * 2.0 ** -a
            d[r, 0] = 1.0
            d[r, pos] = 2.0 ** -a
        oe = np.stack([_operands(row, S) for row in prepare(e)])
        od = np.stack([_operands(row, S) for row in prepare(d)])
        for i in range(e.shape[0]):
            for j in range(d.shape[0]):
                r = rec[k]
                k += 1
                assert r["probe"] == label
                ref = float(np.dot(oe[i].astype(np.float64), od[j].astype(np.float64)))
                big = float(oe[i, 0]) * float(od[j, 0])
                ulp = 2.0 ** (math.floor(math.log2(abs(ref))) - 23)
                assert abs((ref - big) / ulp - r["exact_minus_big_ulps"]) < 1e-3, (label, i, j)   # same operands
                assert abs((_tc_model(oe[i], od[j]) - big) / ulp - r["tc_minus_big_ulps"]) < 1e-3, (label, i, j, r)
    assert k == len(rec) == 324


def test_model_matches_the_probe_statistics_on_random_rows():
    """Probe C on the B200 (profiles/r2_mma_accumulate_probe.txt): over 131 072 random pairs of 3 600 values the
    accumulation error in units of 2^-23 * sum |e'_k d'_k| was 189.6 on average (204 at most) for the all-positive
    NDP rows - the truncations of 228 steps pulling the same way - and 0.63 on average (4.7 at most) for NCC.
    The model on a few pairs of the same kind must land in those ranges."""
    rng = np.random.default_rng(5)
    ev = rng.integers(0, 256, (3, 3600)).astype(np.float64)
    dv = rng.random((2, 3600))
    for centre, lo, hi in ((True, 0.0, 4.8), (False, 170.0, 210.0)):
        for a in ev:
            for b in dv:
                a32 = ((a - a.mean() if centre else a) / np.linalg.norm(a - a.mean() if centre else a)).astype(np.float32)
                b32 = ((b - b.mean() if centre else b) / np.linalg.norm(b - b.mean() if centre else b)).astype(np.float32)
                oe, od = _operands(a32, 3648), _operands(b32, 3648)
                ref = float(np.dot(oe.astype(np.float64), od.astype(np.float64)))
                mag = float(np.dot(np.abs(oe.astype(np.float64)), np.abs(od.astype(np.float64))))
                err = abs(_tc_model(oe, od) - ref) / (mag * 2.0 ** -23)
                assert lo <= err <= hi, (centre, err)
