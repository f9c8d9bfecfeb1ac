"""Target for an ncu capture of the cuBLAS kernels (8192^3 bf16, and 2048 x 100000 x 3648 fp16)."""
import torch
dev = torch.device("cuda", 0)
a8 = torch.randn((8192, 8192), dtype=torch.bfloat16, device=dev); b8 = torch.randn((8192, 8192), dtype=torch.bfloat16, device=dev)
a = torch.randn((2048, 3648), dtype=torch.float16, device=dev); b = torch.randn((100000, 3648), dtype=torch.float16, device=dev)
out = torch.empty((2048, 100000), dtype=torch.float16, device=dev)
for _ in range(12):
    torch.matmul(a8, b8)
for _ in range(12):
    torch.matmul(a, b.T, out=out)
torch.cuda.synchronize()
