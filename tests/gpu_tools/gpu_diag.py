"""Step-by-step GPU diagnostic of libkdi against the NumPy oracle (run on the B200 box).

Prints error statistics per stage so that a wrong descriptor / layout shows up with enough detail
to be fixed without another round trip.  Usage: python tools/gpu_diag.py [quick|full]
"""

import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)

from kikuchipy_b200 import _lib  # noqa: E402
from oracle import di_oracle as orc  # noqa: E402  (checker only)


def f16_operand(a32, bf16=False):
    x = (a32.astype(np.float32) * np.float32(256.0)).astype(np.float32)
    if not bf16:
        return x.astype(np.float16).astype(np.float64)
    u = x.view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16  # round to nearest even
    return r.astype(np.uint32).view(np.float32).astype(np.float64)


def stage(name):
    print(f"\n=== {name} ===", flush=True)


def run(ok, name, fn):
    try:
        t0 = time.time()
        r = fn()
        print(f"[{'PASS' if r else 'FAIL'}] {name} ({time.time() - t0:.2f}s)", flush=True)
        ok.append((name, bool(r)))
    except Exception as e:  # noqa: BLE001
        print(f"[ERROR] {name}: {type(e).__name__}: {e}", flush=True)
        traceback.print_exc()
        ok.append((name, False))


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "quick"
    cgs = tuple(int(c) for c in sys.argv[2].split(",")) if len(sys.argv) > 2 else (1, 2)
    ok = []
    ctx = _lib.Context(0)
    print("device:", ctx.device_info(), flush=True)
    rng = np.random.default_rng(0)

    stage("K1 normalise")

    def norm_case(dtype_name, metric, mask, S=(60, 60), rows=37):
        if dtype_name == "u8":
            src = rng.integers(0, 256, (rows,) + S, dtype=np.uint8)
        elif dtype_name == "u16":
            src = rng.integers(0, 65535, (rows,) + S).astype(np.uint16)
        elif dtype_name == "f64":
            src = rng.random((rows,) + S)
        else:
            src = rng.random((rows,) + S, dtype=np.float32)
        smask = orc.circular_signal_mask(S) if mask else None
        ctx.set_signal_mask(smask)
        p = ctx.patterns(src, rows, _lib.KDI_NCC if metric == "ncc" else _lib.KDI_NDP)
        got = np.asarray(p)
        ref = orc.prepare_dictionary(src.reshape(rows, -1), metric, smask)
        err = np.max(np.abs(got - ref))
        print(f"  {dtype_name} {metric} mask={mask} S={S}: shape {got.shape} max|d|={err:.3e}")
        ctx.set_signal_mask(None)
        return got.shape == ref.shape and err < 2e-7

    for dt in ("u8", "f32", "f64", "u16"):
        for metric in ("ncc", "ndp"):
            for mask in (False, True):
                run(ok, f"normalise {dt} {metric} mask={mask}", lambda: norm_case(dt, metric, mask))
    run(ok, "normalise f32 3x3", lambda: norm_case("f32", "ncc", False, S=(3, 3), rows=9))
    run(ok, "normalise f32 120x120", lambda: norm_case("f32", "ncc", False, S=(120, 120), rows=11))
    run(ok, "normalise f32 61x59 (odd)", lambda: norm_case("f32", "ndp", False, S=(61, 59), rows=5))

    def gemm_case(cg, M, N, S, bf16=False, strip=0, sb=0):
        ctx.set_option(_lib.OPT_CTA_GROUP, cg)
        ctx.set_option(_lib.OPT_COMPUTE_DTYPE, 1 if bf16 else 0)
        ctx.set_option(_lib.OPT_STRIP_TILES, strip)
        ctx.set_option(_lib.OPT_SUPERBLOCK, sb)
        a = rng.random((M, S), dtype=np.float32)
        b = rng.random((N, S), dtype=np.float32)
        pa = ctx.patterns(a, M, _lib.KDI_NCC)
        pb = ctx.patterns(b, N, _lib.KDI_NCC)
        got = ctx.debug_gemm16(pa, pb).astype(np.float64)
        A = f16_operand(np.asarray(pa), bf16)
        B = f16_operand(np.asarray(pb), bf16)
        ref = A @ B.T
        err = np.abs(got - ref)
        scale = np.max(np.abs(ref)) + 1e-30
        rel = err.max() / scale
        print(f"  cg={cg} M={M} N={N} S={S} bf16={bf16} strip={strip} sb={sb}: max|d|/max|ref|={rel:.3e}"
              f" (max|ref|={scale:.3e})")
        good = rel < 2e-5
        if not good:
            bad = err > 1e-4 * scale
            print(f"    bad fraction {bad.mean():.4f}; bad rows {np.unique(np.nonzero(bad)[0])[:16]}"
                  f" bad cols {np.unique(np.nonzero(bad)[1])[:16]}")
            print("    got[:4,:6]\n", got[:4, :6], "\n    ref[:4,:6]\n", ref[:4, :6])
            nanfrac = np.mean(~np.isfinite(got))
            print(f"    non-finite fraction {nanfrac:.4f}")
        ctx.set_option(_lib.OPT_STRIP_TILES, 0)
        ctx.set_option(_lib.OPT_SUPERBLOCK, 0)
        return good

    for cg in cgs:
        stage(f"K2 tensor-core block, cta_group={cg}")
        cases = [(128, 256, 64), (128, 256, 128), (128, 256, 3600), (9, 1000, 3600), (300, 1000, 100),
                 (257, 513, 640), (1000, 3000, 2819)]
        for (M, N, S) in cases:
            run(ok, f"gemm16 cg={cg} {M}x{N}x{S}", lambda: gemm_case(cg, M, N, S))
        run(ok, f"gemm16 cg={cg} bf16", lambda: gemm_case(cg, 300, 1000, 3600, bf16=True))
        run(ok, f"gemm16 cg={cg} strip=1 sb=1", lambda: gemm_case(cg, 700, 2100, 256, strip=1, sb=1))
        run(ok, f"gemm16 cg={cg} strip=3 sb=2", lambda: gemm_case(cg, 700, 2100, 256, strip=3, sb=2))

    ctx.set_option(_lib.OPT_COMPUTE_DTYPE, 0)

    def topk_case(cg, M, N, S, k, metric="ncc", planted=False, bf16=False, exact=False, mask=False):
        ctx.set_option(_lib.OPT_CTA_GROUP, cg)
        ctx.set_option(_lib.OPT_COMPUTE_DTYPE, 1 if bf16 else 0)
        ctx.set_option(_lib.OPT_FORCE_EXACT, 1 if exact else 0)
        dic = orc.synthetic_dictionary(N, S, seed=2)
        exp = orc.planted_experimental(dic, M, seed=3)[0] if planted else orc.synthetic_experimental(M, S, seed=1)
        smask = orc.circular_signal_mask(S) if mask else None
        ctx.set_signal_mask(smask)
        t0 = time.time()
        idx, sc = ctx.dictionary_indexing(exp, M, dic, N, _lib.KDI_NCC if metric == "ncc" else _lib.KDI_NDP, k)
        dt = time.time() - t0
        tm = ctx.timings()
        ridx, rsc = orc.dictionary_indexing(exp, dic, metric=metric, keep_n=k, signal_mask=smask)
        r = orc.compare_topk(ridx, rsc, idx, sc)
        print(f"  cg={cg} {M}x{N}x{S} k={k} {metric} planted={planted} bf16={bf16} exact={exact} mask={mask}: "
              f"{r} flagged={tm['flagged_rows']} gemm={tm['gemm_topk_ms']:.3f}ms rescore={tm['rescore_ms']:.3f}ms"
              f" total={tm['total_ms']:.3f}ms wall={dt * 1e3:.1f}ms")
        ctx.set_option(_lib.OPT_FORCE_EXACT, 0)
        ctx.set_option(_lib.OPT_COMPUTE_DTYPE, 0)
        ctx.set_signal_mask(None)
        good = r["tie_ok"] and r["scores_ok"]
        if planted:
            good = good and r["exact_rows"] == 1.0
        return good

    stage("exact path (no tensor cores)")
    run(ok, "exact 64x4096 k=20", lambda: topk_case(1, 64, 4096, (60, 60), 20, exact=True))
    run(ok, "exact 9x100 k=100", lambda: topk_case(1, 9, 100, (12, 12), 100, exact=True))
    for cg in cgs:
        stage(f"fused match + top-k, cta_group={cg}")
        run(ok, f"topk cg={cg} 64x4096 k=20", lambda: topk_case(cg, 64, 4096, (60, 60), 20))
        run(ok, f"topk cg={cg} planted", lambda: topk_case(cg, 64, 4096, (60, 60), 20, planted=True))
        run(ok, f"topk cg={cg} k=50", lambda: topk_case(cg, 300, 5000, (60, 60), 50))
        run(ok, f"topk cg={cg} k=1 ndp mask", lambda: topk_case(cg, 200, 3000, (60, 60), 1, metric="ndp", mask=True))
        run(ok, f"topk cg={cg} bf16", lambda: topk_case(cg, 200, 3000, (60, 60), 20, bf16=True))
        run(ok, f"topk cg={cg} tiny dict", lambda: topk_case(cg, 9, 20, (60, 60), 20))
        if mode == "full":
            run(ok, f"topk cg={cg} 1000x20000", lambda: topk_case(cg, 1000, 20000, (60, 60), 20))

    stage("summary")
    bad = [n for n, g in ok if not g]
    print(f"{len(ok) - len(bad)}/{len(ok)} passed")
    for n in bad:
        print("  FAILED:", n)
    return 0 if not bad else 1


if __name__ == "__main__":
    sys.exit(main())
