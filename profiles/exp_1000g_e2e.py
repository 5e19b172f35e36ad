#!/usr/bin/env python
"""1000G-scale end-to-end call (variant-major page-locked host array, only the referenced rows uploaded by the SM
gather over PCIe) as a function of the sample chunking of transform_host."""
import json, os, sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from haptools_b200 import _cuda, engine, synth
def best(fn, reps=5):
    fn(); ts = []
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    return min(ts)
p, H, n = 1_000_000, 10_000, 2_504
plan = synth.synth_plan(p, H, seed=1)
pinned = engine.pinned_empty((p, n, 2))
pat = synth.synth_gt(n, 512, 2, mode=1).transpose(1, 0, 2)
for v0 in range(0, p, 512):
    pinned[v0:v0 + 512] = pat[: min(512, p - v0)]
view = pinned.transpose(1, 0, 2)
res = {}
ref = engine.transform_host(view, plan).copy()
for chunk in (None, 2504, 1280):
    for result in ("auto", "u8"):
        t = best(lambda: engine.transform_host(view, plan, chunk_samples=chunk, result=result))
        same = bool(np.array_equal(engine.transform_host(view, plan, chunk_samples=chunk, result=result), ref))
        res[f"chunk={chunk} result={result}"] = {"ms": t * 1e3, "same": same}
        print(f"chunk={chunk} result={result}", res[f"chunk={chunk} result={result}"], flush=True)
Path("gpurun_out/e2e_1000g_chunks.json").write_text(json.dumps(res, indent=1))
