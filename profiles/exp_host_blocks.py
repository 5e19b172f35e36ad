#!/usr/bin/env python
"""Host copy engine block sizes (HT_HOST_BLOCK_KB: upload blocks of a pageable source; HT_HOST_EXPAND_BLOCK_KB:
bit-packed result blocks before widening): do smaller blocks keep the page-locked staging slots in the last-level
cache and take their traffic off the host's DRAM?  transform_host, UKB-scale plan, 16,384 samples, result='bits'."""
import json, os, sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from haptools_b200 import _cuda, engine, synth
ne, p, H = 16384, 50_000, 100_000
plan = synth.synth_plan(p, H, seed=2)
dplan = engine.DevicePlan(plan)
base = synth.synth_gt(256, p, 3, mode=1)
pinned_in = engine.pinned_empty((ne, p, 2)); pinned_in[:] = base[np.arange(ne) % 256]
pageable_in = np.array(pinned_in)
out = np.empty((ne, H, 2), dtype=np.uint8); out[:] = 0
def best(gt, reps=4):
    ts = []
    for i in range(reps + 1):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        engine.transform_host(gt, plan, out=out, dplan=dplan, result="bits")
        torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    return min(ts[1:]) * 1e3
res = {}
for up_kb, ex_kb in ((4096, 1024), (8192, 512), (4096, 768), (4096, 512), (4096, 384), (8192, 384), (4096, 512), (4096, 1024)):
    os.environ["HT_HOST_BLOCK_KB"], os.environ["HT_HOST_EXPAND_BLOCK_KB"] = str(up_kb), str(ex_kb)
    r = {"pinned_in_ms": best(pinned_in), "pageable_in_ms": best(pageable_in)}
    res[f"upload {up_kb} KB, expand {ex_kb} KB"] = r
    print(up_kb, ex_kb, r, flush=True)
Path("gpurun_out/host_blocks.json").write_text(json.dumps(res, indent=1))
