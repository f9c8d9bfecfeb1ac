"""Sharded refinement over NCCL (run with torchrun on N GPUs of the GPU box):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      tests/gpu_tools/refine_sharded_check.py [n_patterns]

Every rank refines its slice of the patterns and all-gathers the finished rows; rank 0 compares the
result with an unsharded run on its own GPU (must be identical) and with the oracle on a sample,
and prints one JSON line with the device time (max over ranks, CUDA events) and patterns/s."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))
import kikuchipy_b200 as kb  # noqa: E402
from kikuchipy_b200 import refinement as rf  # noqa: E402
from oracle import refinement_oracle as ro  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl")
base = ro.synthetic_case(n=256, nrows=60, ncols=60, mp_size=501, seed=11, noise=0.05, perturb_deg=1.0)
reps = -(-n // 256)
pats = np.tile(base["patterns"], (reps, 1))[:n].reshape(n, 60, 60)
quat = rf.euler_to_quaternion(np.tile(base["start_eulers"], (reps, 1))[:n])


class Det:
    shape = (60, 60)
    pc = base["pc"][None]
    om_detector_to_sample = base["om"]


ctx = kb.default_context(local)
kw = dict(compute=False, verbose=False, context=ctx)
kb.refine_orientation(pats[:512], quat[:512], Det, (base["mu"], base["ml"]), sharded=True, **kw)  # warm-up
dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
res = kb.refine_orientation(pats, quat, Det, (base["mu"], base["ml"]), sharded=True, **kw)
torch.cuda.synchronize()
dist.barrier()
wall = time.perf_counter() - t0
ms = torch.tensor([ctx.timings()["total_ms"]], device="cuda")
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    single = kb.refine_orientation(pats, quat, Det, (base["mu"], base["ml"]), **kw)
    x0 = rf.quaternion_to_euler(quat[:8])[:, None, :]
    want = ro.refine_orientation(base["problem"], pats[:8].reshape(8, -1), x0, False)
    print(json.dumps({"n_gpus": world, "patterns": n, "kernel_ms_max_over_ranks": round(float(ms), 3),
                      "patterns_per_s_kernel": round(n / float(ms) * 1e3), "patterns_per_s_call": round(n / wall),
                      "identical_to_one_gpu": bool(np.array_equal(res, single)),
                      "max_dscore_vs_oracle_sample": float(np.abs(res[:8, 0] - want[:, 0]).max())}))
dist.destroy_process_group()
