"""Experiment: does the one-kernel transform get faster when every plane read is a near-L2 hit?
Same haplotype count and list lengths (K = 1.1 M reads per tile), shrinking key space: with few keys
each record is read hundreds of times per tile, so first-touch (far-partition / DRAM) reads vanish."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from haptools_b200 import engine, synth
dev = torch.device("cuda")
n, H = 200_000, 100_000
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
out = torch.empty((n, 2 * H), dtype=torch.uint8, device=dev)
for p in (50_000, 10_000, 2_000, 500, 100):
    plan = synth.synth_plan(p, H, seed=2)
    dp = engine.DevicePlan(plan)
    planes = engine.alloc_planes(n, plan.n_keys, dev); planes.random_(0, 2**31 - 1)
    ms = timeit(lambda: engine.transform_u8(dp, planes, n, out.data_ptr(), 2 * H, method="direct"))
    print(f"p={p:6d} keys={plan.n_keys:6d} ({plan.n_keys*256/1e6:6.2f} MB of records per tile) K={plan.n_refs}: {ms:.2f} ms")
