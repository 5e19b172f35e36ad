import sys, time, numpy as np, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from haptools_b200 import engine, synth, _cuda
from haptools_b200.plan import plan_from_refs
dev = torch.device('cuda')
def timeit(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
GB = 1 << 30
a = torch.empty(48 * GB, dtype=torch.uint8, device=dev)
b = torch.empty(48 * GB, dtype=torch.uint8, device=dev)
ms = timeit(lambda: a.fill_(1)); print(f"fill_ 48GiB: {ms:.2f} ms  {a.numel()/ms/1e6:.0f} GB/s write")
ms = timeit(lambda: b.copy_(a)); print(f"copy_ 48GiB: {ms:.2f} ms  {2*a.numel()/ms/1e6:.0f} GB/s r+w")
ms = timeit(lambda: a.sum(dtype=torch.int64) if False else torch.sum(a.view(torch.int64)[: a.numel()//8])); print(f"sum 48GiB: {ms:.2f} ms  {a.numel()/ms/1e6:.0f} GB/s read")
del a, b
torch.cuda.empty_cache()
n, H, p = 200_000, 100_000, 50_000
# (1) empty haplotypes: pure write through my store path
plan0 = plan_from_refs(np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64), np.zeros(H + 1, dtype=np.int64), p)
d0 = engine.DevicePlan(plan0)
out = torch.empty((n, 2 * H), dtype=torch.uint8, device=dev)
planes = engine.alloc_planes(n, 1, dev)
ms = timeit(lambda: engine.transform_u8(d0, planes, n, out.data_ptr(), 2 * H)); print(f"transform empty haps: {ms:.2f} ms  {n*2*H/ms/1e6:.0f} GB/s write")
# (2) single-key haplotypes / (3) real plan
for name, kw in (("1 allele/hap", dict(min_alleles=1, max_alleles=1)), ("2-20 alleles", {}), ("20 alleles", dict(min_alleles=20, max_alleles=20))):
    plan = synth.synth_plan(p, H, seed=2, **kw)
    dp = engine.DevicePlan(plan)
    planes = engine.alloc_planes(n, plan.n_keys, dev)
    planes.random_(0, 2**31 - 1)
    ms = timeit(lambda: engine.transform_u8(dp, planes, n, out.data_ptr(), 2 * H))
    print(f"transform {name}: K={plan.n_refs} A={plan.n_keys} {ms:.2f} ms  out {n*2*H/ms/1e6:.0f} GB/s")
# (4) smaller H -> narrower rows
for H2 in (12_800, 1280):
    plan = synth.synth_plan(p, H2, seed=2)
    dp = engine.DevicePlan(plan)
    planes = engine.alloc_planes(n, plan.n_keys, dev); planes.random_(0, 2**31 - 1)
    ms = timeit(lambda: engine.transform_u8(dp, planes, n, out.data_ptr(), 2 * H2))
    print(f"transform H={H2}: {ms:.2f} ms  out {n*2*H2/ms/1e6:.0f} GB/s")
