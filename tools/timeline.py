"""Per-launch timeline of one device-resident BASELINE-config-2 call (run with KDI_TIMELINE=1)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import kikuchipy_b200 as kb
from kikuchipy_b200 import _lib
M, N, SIG = int(os.environ.get("M", "10000")), int(os.environ.get("N", "100000")), (60, 60)
ctx = kb.default_context(0)
for name, opt in (("OVERLAP", _lib.OPT_OVERLAP), ("SUPERBLOCK", _lib.OPT_SUPERBLOCK), ("MAX_STAGES", _lib.OPT_MAX_STAGES)):
    if os.environ.get(name):
        ctx.set_option(opt, int(os.environ[name]))
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(1)
exp = torch.randint(0, 256, (M,) + SIG, dtype=torch.uint8, device=dev, generator=g)
dic = torch.rand((N,) + SIG, dtype=torch.float32, device=dev, generator=g)
idx = torch.empty((M, 20), dtype=torch.int64, device=dev); sc = torch.empty((M, 20), dtype=torch.float32, device=dev)
torch.cuda.synchronize()
for i in range(4):
    if i == 3:
        sys.stderr.write("==== %s\n" % {k: os.environ.get(k) for k in ("OVERLAP", "KDI_CARVEOUT", "SUPERBLOCK", "MAX_STAGES")})
    else:
        sys.stderr.flush()
    ctx.dictionary_indexing(exp, M, dic, N, _lib.KDI_NCC, 20, out=(idx, sc))
print({k: round(v, 3) if isinstance(v, float) else v for k, v in ctx.timings().items()})
