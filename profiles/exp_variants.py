"""Time the one-kernel transform of whatever library HT_LIB_PATH names (build variants)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from haptools_b200 import engine, synth
dev = torch.device("cuda")
n, p, H = 200_000, 50_000, 100_000
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
out = torch.empty((n, 2 * H), dtype=torch.uint8, device=dev)
res = []
for name, kw in (("2-20", {}), ("20", dict(min_alleles=20, max_alleles=20)), ("2-20 sorted", dict(sort=True))):
    plan = synth.synth_plan(p, H, seed=2, **kw)
    dp = engine.DevicePlan(plan)
    planes = engine.alloc_planes(n, plan.n_keys, dev); planes.random_(0, 2**31 - 1)
    ms = timeit(lambda: engine.transform_u8(dp, planes, n, out.data_ptr(), 2 * H, method="direct"))
    res.append(f"{name}: {ms:.2f} ms")
print(os.environ.get("HT_LIB_PATH", "default"), " | ".join(res))
