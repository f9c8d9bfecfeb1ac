"""KDI_TIMELINE=1 per-launch timeline of one BASELINE configuration (tools/di_configs.py) on one GPU,
optionally scaled down, with library options from the environment (OPTS="18=0,9=1": option=value)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import kikuchipy_b200 as kb
from tools import di_configs as dc

ctx = kb.default_context(0)
for kv in filter(None, os.environ.get("OPTS", "").split(",")):
    o, v = kv.split("=")
    ctx.set_option(int(o), float(v))
r = dc.run_config(int(os.environ.get("CONFIG", "3")), ctx, 0, 1, torch.device("cuda", 0), steps=2, warmup=1,
                  sample64=int(os.environ.get("SAMPLE64", "64")), scale=float(os.environ.get("SCALE", "1.0")))
print(json.dumps(r))
