#!/usr/bin/env python
"""Two-phase ("local") transform with an L2-RESIDENT intermediate: ht_transform_u8_local processes groups of
workspace_bytes / (H * 256) tiles; with 1-4 tiles per group the bit-packed records of a group (25.6 MB per tile at
100 k haplotypes) can stay in the 126 MB L2 between the phase that writes them (haplotypes in key order: few L2->SM
plane reads) and the phase that expands them to bytes.  UKB-scale plan in generation (unsorted) order, N samples."""
import json, os, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from haptools_b200 import engine, synth
n = int(os.environ.get("N", 200 * 1024))
def timeit(fn, reps=4):
    fn(); fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
plan = synth.synth_plan(50_000, 100_000, seed=2)
dplan = engine.DevicePlan(plan)
planes = engine.alloc_planes(n, plan.n_keys, "cuda"); planes.random_(-(2**31), 2**31 - 1)
pitch = engine.out_pitch(plan.n_haps)
out = torch.empty((n, pitch), dtype=torch.uint8, device="cuda")
res = {"n": n}
ms = timeit(lambda: engine.transform_u8(dplan, planes, n, out.data_ptr(), pitch, method="direct"))
ref = out.clone(); res["direct"] = ms; print("direct", ms, flush=True)
for g in (1, 2, 3, 4, 8, 64):
    ws = torch.empty(g * plan.n_haps * 256 // 4, dtype=torch.int32, device="cuda")
    ms = timeit(lambda: engine.transform_u8(dplan, planes, n, out.data_ptr(), pitch, method="local", workspace=ws))
    same = bool(torch.equal(out, ref))
    res[f"local, {g} tiles per group"] = {"ms": ms, "same": same}; print("local", g, ms, same, flush=True)
Path("gpurun_out/local_l2.json").write_text(json.dumps(res, indent=1))
