"""Per-launch timeline of one device-resident BASELINE-config-2 call (run with KDI_TIMELINE=1)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import kikuchipy_b200 as kb
from kikuchipy_b200 import _lib
M, N, SIG = int(os.environ.get("M", "10000")), int(os.environ.get("N", "100000")), (60, 60)
KEEP = int(os.environ.get("KEEP", "20"))
ctx = kb.default_context(0)
for name, opt in (("OVERLAP", _lib.OPT_OVERLAP), ("SUPERBLOCK", _lib.OPT_SUPERBLOCK), ("MAX_STAGES", _lib.OPT_MAX_STAGES)):
    if os.environ.get(name):
        ctx.set_option(opt, int(os.environ[name]))
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(1)
exp = torch.randint(0, 256, (M,) + SIG, dtype=torch.uint8, device=dev, generator=g)
dic = torch.rand((N,) + SIG, dtype=torch.float32, device=dev, generator=g)
idx = torch.empty((M, KEEP), dtype=torch.int64, device=dev); sc = torch.empty((M, KEEP), dtype=torch.float32, device=dev)
torch.cuda.synchronize()
gen = None
if os.environ.get("GEN"):
    from kikuchipy_b200 import synthetic as po
    mu, ml = po.synthetic_master_pattern(1001, seed=5)
    dcs = kb.direction_cosines([-0.9, 0.85, -0.7, 0.95], 0.5, 60, 60, po.tilted_detector_matrix(70.0))
    rot = torch.from_numpy(po.random_rotations(N, seed=4)).cuda()
    gen = ctx.master_pattern(mu, ml, dcs)
for i in range(4):
    if i == 3:
        sys.stderr.write("==== %s\n" % {k: os.environ.get(k) for k in ("OVERLAP", "KDI_CARVEOUT", "KDI_GEMM_CARVEOUT", "SUPERBLOCK", "MAX_STAGES", "GEN")})
    else:
        sys.stderr.flush()
    if gen is not None:
        ctx.dictionary_indexing_projected(exp, M, gen, rot, _lib.KDI_NCC, KEEP, out=(idx, sc))
    else:
        ctx.dictionary_indexing(exp, M, dic, N, _lib.KDI_NCC, KEEP, out=(idx, sc))
print({k: round(v, 3) if isinstance(v, float) else v for k, v in ctx.timings().items()})
