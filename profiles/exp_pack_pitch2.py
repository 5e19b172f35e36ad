import os, sys, torch
sys.path.insert(0, os.getcwd())
from haptools_b200 import _cuda, engine, synth
dev = torch.device("cuda"); lib = _cuda.lib()
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for (n, p, H) in ((400_000, 50_000, 25_000), (100_000, 200_000, 100_000), (40_000, 500_000, 250_000), (20_000, 1_000_000, 500_000)):
    plan = synth.synth_plan(p, H, seed=4)
    dp = engine.DevicePlan(plan)
    gt = torch.empty((n, p, 2), dtype=torch.uint8, device=dev)
    _cuda.check(lib.ht_synth_gt(gt.data_ptr(), *gt.stride(), n, p, 0, 5, 1, 0, 0), "synth")
    planes = engine.alloc_planes(n, plan.n_keys, dev)
    ab = plan.algorithmic_bytes(n)["pack"]
    ms = timeit(lambda: engine.pack(dp, gt.data_ptr(), gt.stride(), n, planes))
    print(f"n={n} p={p} pitch={2*p/1e3:.0f} KB keys={plan.n_keys}: {ms:.2f} ms, {ab/1e9:.1f} GB, {ab/ms/1e6:.0f} GB/s", flush=True)
    del gt, planes; torch.cuda.empty_cache()
