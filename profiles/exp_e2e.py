"""Experiment: end-to-end (pinned host -> device -> pinned host) throughput vs pipeline chunk size."""
import os, sys, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from haptools_b200 import engine, synth
p, H, ne = 50_000, 100_000, 16_384
plan = synth.synth_plan(p, H, seed=2)
dplan = engine.DevicePlan(plan)
host_gt = engine.pinned_empty((ne, p, 2))
host_gt[:] = synth.synth_gt(256, p, 3, mode=1)[np.arange(ne) % 256]
host_out = engine.pinned_empty((ne, H, 2))
for chunk in (None, 512, 1024, 2048, 4096, 16384):  # None = the default schedule (ramped lead-in)
    for _ in range(2):
        engine.transform_host(host_gt, plan, out=host_out, dplan=dplan, chunk_samples=chunk)
    ts = []
    for _ in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        engine.transform_host(host_gt, plan, out=host_out, dplan=dplan, chunk_samples=chunk)
        torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    dt = min(ts)
    print(f"chunk {str(chunk):>6s}: {dt*1e3:7.2f} ms  {ne*H*2/dt:.3e} calls/s  H2D+D2H {(host_gt.nbytes+host_out.nbytes)/dt/1e9:.1f} GB/s")
