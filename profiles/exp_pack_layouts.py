import sys, torch, numpy as np
sys.path.insert(0, ".")
from haptools_b200 import engine, synth, _cuda
dev = torch.device("cuda")
n, p, H = 200_000, 50_000, 100_000
plan = synth.synth_plan(p, H, seed=2)
dp = engine.DevicePlan(plan)
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
planes = engine.alloc_planes(n, plan.n_keys, dev)
for name, gt in (("sample-major", torch.empty((n, p, 2), dtype=torch.uint8, device=dev)),
                 ("variant-major", torch.empty((p, n, 2), dtype=torch.uint8, device=dev).permute(1, 0, 2))):
    _cuda.check(_cuda.lib().ht_synth_gt(gt.data_ptr(), *gt.stride(), n, p, 0, 3, 1, 0, 0), "synth")
    ms = timeit(lambda: engine.pack(dp, gt.data_ptr(), gt.stride(), n, planes))
    ab = plan.algorithmic_bytes(n)["pack"]
    print(f"pack {name}: {ms:.2f} ms  {ab/ms/1e6:.0f} GB/s")
    del gt
