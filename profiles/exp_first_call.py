"""One-shot cost of the result buffer (a fresh process, as `haptools transform` is): page-locked allocation
against a plain numpy array that the staged download fills (first touch included)."""
import os, sys, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from haptools_b200 import engine, synth
p, H, ne = 50_000, 100_000, 16_384
plan = synth.synth_plan(p, H, seed=2)
dplan = engine.DevicePlan(plan)
host_gt = engine.pinned_empty((ne, p, 2))
host_gt[:] = synth.synth_gt(256, p, 3, mode=1)[np.arange(ne) % 256]
engine.transform_host(host_gt[:2048], plan, dplan=dplan)  # CUDA context, kernels, small buffers
torch.cuda.synchronize()
mode = sys.argv[1]
t0 = time.perf_counter()
out = engine.pinned_empty((ne, H, 2)) if mode == "pinned" else np.empty((ne, H, 2), dtype=np.uint8)
t1 = time.perf_counter()
engine.transform_host(host_gt, plan, out=out, dplan=dplan)
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"{mode}: allocate {1e3*(t1-t0):.0f} ms + first call {1e3*(t2-t1):.0f} ms = {1e3*(t2-t0):.0f} ms for {out.nbytes/1e9:.2f} GB of result", flush=True)
