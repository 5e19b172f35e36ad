#!/usr/bin/env python
"""A few ancestry pack launches for ncu: N samples (default 32768) x P variants (default 200000) x H haplotypes,
5 populations; HT_PACK_ANC_KERNEL selects the form."""
import os, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from haptools_b200 import _cuda, engine, synth
n, p, H = int(os.environ.get("N", 32768)), int(os.environ.get("P", 200_000)), int(os.environ.get("H", 50_000))
dev = torch.device("cuda"); lib = _cuda.lib()
plan = synth.synth_plan(p, H, seed=4, n_pops=5)
dp = engine.DevicePlan(plan)
gt = torch.empty((n, p, 2), dtype=torch.uint8, device=dev)
_cuda.check(lib.ht_synth_gt(gt.data_ptr(), *gt.stride(), n, p, 0, 5, 1, 0, 0), "synth")
anc = torch.empty((n, p, 2), dtype=torch.uint8, device=dev)
_cuda.check(lib.ht_synth_ancestry(anc.data_ptr(), *anc.stride(), n, p, 0, 6, 5, 2000, 0), "synth_anc")
planes = engine.alloc_planes(n, plan.n_keys, dev)
for _ in range(int(os.environ.get("LAUNCHES", 3))):
    engine.pack(dp, gt.data_ptr(), gt.stride(), n, planes, anc.data_ptr(), anc.stride())
torch.cuda.synchronize()
print("done", plan.algorithmic_bytes(n)["pack"] / 1e9, "GB algorithmic")
