"""End to end with the result in PAGEABLE host memory (what transform_host returns when the result is larger
than 4 GiB, or when the caller passes its own numpy array) against a page-locked result."""
import os, sys, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from haptools_b200 import engine, synth
p, H, ne = 50_000, 100_000, 16_384
plan = synth.synth_plan(p, H, seed=2)
dplan = engine.DevicePlan(plan)
host_gt = engine.pinned_empty((ne, p, 2))
host_gt[:] = synth.synth_gt(256, p, 3, mode=1)[np.arange(ne) % 256]
outs = {"pinned": engine.pinned_empty((ne, H, 2)), "pageable": np.empty((ne, H, 2), dtype=np.uint8)}
outs["pageable"][:] = 0  # touch the pages
for name, out in outs.items():
    for _ in range(2):
        engine.transform_host(host_gt, plan, out=out, dplan=dplan)
    ts = []
    for _ in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        engine.transform_host(host_gt, plan, out=out, dplan=dplan)
        torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    dt = min(ts)
    print(f"result {name}: {dt*1e3:7.1f} ms  {ne*H*2/dt:.3e} hap-calls/s  D2H {out.nbytes/dt/1e9:.1f} GB/s", flush=True)
assert np.array_equal(outs["pinned"], outs["pageable"])
