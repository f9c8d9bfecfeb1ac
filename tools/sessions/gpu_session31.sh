#!/bin/bash
# sharded refinement over NCCL on 2 GPUs
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/gpu_tools/refine_sharded_check.py 20000 2>&1 | tail -3 | tee gpurun_out/refine_sharded_n2.json
