#!/usr/bin/env python
"""One transform launch for ncu (round 2): METHOD=direct|staged, SORT=1|0, N samples (default 131072) of the
UKB-scale plan on random plane contents.  Two warm launches, then the ones ncu should look at (-s 2)."""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from haptools_b200 import engine, synth  # noqa: E402

n = int(os.environ.get("N", 131072))
method = os.environ.get("METHOD", "direct")
plan = synth.synth_plan(50_000, 100_000, seed=2, sort=os.environ.get("SORT", "1") == "1")
dplan = engine.DevicePlan(plan)
planes = engine.alloc_planes(n, plan.n_keys, "cuda")
planes.random_(-(2**31), 2**31 - 1)
out = torch.empty((n, engine.out_pitch(plan.n_haps)), dtype=torch.uint8, device="cuda")
for _ in range(int(os.environ.get("LAUNCHES", 3))):
    engine.transform_u8(dplan, planes, n, out.data_ptr(), out.stride(0), method=method)
torch.cuda.synchronize()
print("done", method, dplan.stage_lines)
