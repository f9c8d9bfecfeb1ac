"""cuBLAS bf16 / fp16 GEMM rate on THIS box next to the library's GEMM span at BASELINE config 2, interleaved:
how much of the gap to MEASURED_PEAKS.json is the box (power cap) and how much the kernel."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import kikuchipy_b200 as kb
from kikuchipy_b200 import _lib

dev = torch.device("cuda", 0)
ctx = kb.default_context(0)
M, N, S, K = 10000, 100000, 3600, 20
g = torch.Generator(device=dev); g.manual_seed(1)
exp = torch.randint(0, 256, (M, 60, 60), dtype=torch.uint8, device=dev, generator=g)
dic = torch.rand((N, 60, 60), dtype=torch.float32, device=dev, generator=g)
idx = torch.empty((M, K), dtype=torch.int64, device=dev); sc = torch.empty((M, K), dtype=torch.float32, device=dev)
a = torch.randn((M, 3648), dtype=torch.float16, device=dev); b = torch.randn((N, 3648), dtype=torch.float16, device=dev)
a8 = torch.randn((8192, 8192), dtype=torch.bfloat16, device=dev); b8 = torch.randn((8192, 8192), dtype=torch.bfloat16, device=dev)
out = torch.empty((2048, N), dtype=torch.float16, device=dev)

def ev_ms(fn, reps):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

def cublas_same_shape():  # the same M x N x K product in row blocks (the full fp16 output would be 2 GB)
    for r0 in range(0, M, 2048):
        torch.matmul(a[r0:r0 + 2048], b.T, out=out[: min(2048, M - r0)])

res = {"ours_gemm_ms": [], "ours_step_ms": [], "cublas_same_shape_ms": [], "cublas_8192_ms": []}
for rnd in range(4):
    for _ in range(3):
        ctx.dictionary_indexing(exp, M, dic, N, _lib.KDI_NCC, K, out=(idx, sc))
    t = ctx.timings(); res["ours_gemm_ms"].append(t["gemm_topk_ms"]); res["ours_step_ms"].append(t["total_ms"])
    res["cublas_same_shape_ms"].append(ev_ms(cublas_same_shape, 3))
    res["cublas_8192_ms"].append(ev_ms(lambda: torch.matmul(a8, b8), 10))
flops = 2.0 * M * N * S
summary = {
    "ours_gemm_tflops": [round(flops / (x * 1e-3) / 1e12, 1) for x in res["ours_gemm_ms"]],
    "cublas_same_shape_tflops_padded_k": [round(2.0 * M * N * 3648 / (x * 1e-3) / 1e12, 1) for x in res["cublas_same_shape_ms"]],
    "cublas_same_shape_tflops_algorithmic": [round(flops / (x * 1e-3) / 1e12, 1) for x in res["cublas_same_shape_ms"]],
    "cublas_8192_tflops": [round(2.0 * 8192 ** 3 / (x * 1e-3) / 1e12, 1) for x in res["cublas_8192_ms"]],
    "note": "rounds interleaved on one box; cuBLAS writes the full fp16 score block (2 GB per step) that the library never materialises",
}
print(json.dumps({**res, **summary}))
