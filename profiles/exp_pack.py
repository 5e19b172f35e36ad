"""Experiment: sample-major pack bandwidth vs row width (DRAM locality) at constant bytes."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from haptools_b200 import engine, _cuda
from haptools_b200.plan import plan_from_refs
dev = torch.device('cuda')
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
total = 8 << 30
for p in (512, 2048, 8192, 50_000, 200_000):
    n = total // (2 * p) // 1024 * 1024
    # every column referenced with both alleles: one haplotype per (column, allele)
    cols = np.repeat(np.arange(p), 2); al = np.tile([0, 1], p)
    plan = plan_from_refs(cols, al, np.arange(2 * p + 1), p)
    dp = engine.DevicePlan(plan)
    gt = torch.randint(0, 2, (n, p, 2), dtype=torch.uint8, device=dev)
    planes = engine.alloc_planes(n, plan.n_keys, dev)
    ms = timeit(lambda: engine.pack(dp, gt.data_ptr(), gt.stride(), n, planes))
    b = plan.algorithmic_bytes(n)["pack"]
    print(f"p={p:7d} n={n:8d}: {ms:7.3f} ms  {b/ms/1e6:6.0f} GB/s (gt {n*p*2/1e9:.1f} GB + planes {b/1e9 - n*p*2/1e9:.1f} GB)")
    del gt, planes
