"""How the tensor core accumulates: the constant of the accumulation term in the certificate's worst-case
bound (kdi_internal.cuh: kdi_cert_bound assumes <= 1 ulp of the largest magnitude per addend of every 16-deep
step, 18 ulp per step).  Through kdi_debug_gemm16 (the library's own TMA / tcgen05 pipeline, raw float32
accumulators) with operands whose 16-bit values are known exactly (read back, rounded on the host the way
the prepare kernel rounds them), against the float64 sum of the exact products.

A: one large product and 15 tiny ones of relative size 2^-s inside ONE 16-deep step - down to which s do the
   tiny addends still count (guard bits of the alignment)?
B: the large product in the first step and one tiny product in each of the 63 later steps - the same
   question for the accumulator carried from step to step.
C: random unit rows of 3 600 values: accumulation error in units of 2^-23 * sum |e'_k d'_k|, against the
   18 * K / 16 the bound allows."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import kikuchipy_b200 as kb
from kikuchipy_b200 import _lib

ctx = kb.default_context(0)
S = 1024
A_EXP = list(range(6, 15))
B_EXP = list(range(6, 15))


def operands(p):
    """the fp16 operand rows of a prepared set, as float64 (prepare kernel: RN of row * 256)"""
    return (np.asarray(p).astype(np.float32) * np.float32(256.0)).astype(np.float16).astype(np.float64)


def run(e_rows, d_rows, label, n_tiny):
    e = ctx.patterns(np.ascontiguousarray(e_rows, dtype=np.float32), e_rows.shape[0], _lib.KDI_NDP)
    d = ctx.patterns(np.ascontiguousarray(d_rows, dtype=np.float32), d_rows.shape[0], _lib.KDI_NDP)
    tc = ctx.debug_gemm16(e, d).astype(np.float64)
    oe, od = operands(e), operands(d)
    ref = oe @ od.T
    big = oe[:, :1] * od[None, :, 0]
    ulp = 2.0 ** (np.floor(np.log2(np.abs(ref))) - 23)
    for i in range(e_rows.shape[0]):
        for j in range(d_rows.shape[0]):
            tiny = abs(oe[i, pos_tiny[0]] * od[j, pos_tiny[0]])
            s = int(round(np.log2(big[i, j] / tiny))) if tiny > 0 else None
            print(json.dumps({"probe": label, "rel_size_log2": -s if s is not None else None,
                              "sign": "+" if oe[i, pos_tiny[0]] > 0 else "-", "n_tiny": n_tiny,
                              "exact_minus_big_ulps": round(float((ref[i, j] - big[i, j]) / ulp[i, j]), 4),
                              "tc_minus_big_ulps": round(float((tc[i, j] - big[i, j]) / ulp[i, j]), 4),
                              "error_ulps": round(float((tc[i, j] - ref[i, j]) / ulp[i, j]), 4)}), flush=True)
    e.close(); d.close()


for label, pos_tiny in (("A_one_step", list(range(1, 16))), ("B_across_steps", [16 * m for m in range(1, 64)])):
    e_rows = np.zeros((2 * len(A_EXP), S), dtype=np.float32)
    d_rows = np.zeros((len(B_EXP), S), dtype=np.float32)
    for r, a in enumerate(A_EXP):
        for sgn in (0, 1):
            e_rows[2 * r + sgn, 0] = 1.0
            e_rows[2 * r + sgn, pos_tiny] = (1.0 if sgn == 0 else -1.0) * 2.0 ** -a
    for r, b in enumerate(B_EXP):
        d_rows[r, 0] = 1.0
        d_rows[r, pos_tiny] = 2.0 ** -b
    run(e_rows, d_rows, label, len(pos_tiny))

# C: random rows of the benchmark's kind
rng = np.random.default_rng(5)
for metric, name in ((_lib.KDI_NCC, "ncc"), (_lib.KDI_NDP, "ndp")):
    ev = rng.integers(0, 256, (256, 3600)).astype(np.float32)
    dv = rng.random((512, 3600), dtype=np.float32)
    e = ctx.patterns(ev, 256, metric); d = ctx.patterns(dv, 512, metric)
    tc = ctx.debug_gemm16(e, d).astype(np.float64)
    oe, od = operands(e), operands(d)
    ref = oe @ od.T
    mag = np.abs(oe) @ np.abs(od).T
    unit = mag * 2.0 ** -23
    err = np.abs(tc - ref) / unit
    print(json.dumps({"probe": "C_random_3600", "metric": name, "pairs": int(err.size),
                      "max_error_in_units_of_2^-23_sum_abs_products": round(float(err.max()), 4),
                      "mean": round(float(err.mean()), 4), "bound_allows": 18 * 3648 / 16,
                      "max_abs_error_of_score": float(np.abs(tc - ref).max() / 65536.0),
                      "accumulation_term_of_bound": 18 * 3648 / 16 * 2.0 ** -23}), flush=True)
    e.close(); d.close()
