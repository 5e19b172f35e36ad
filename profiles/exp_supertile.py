"""Experiment: pack -> transform per super-tile of S tiles (planes stay in L2) vs whole-cohort launches."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from haptools_b200 import engine, synth, _cuda
dev = torch.device('cuda')
n, p, H = 200_000, 50_000, 100_000
plan = synth.synth_plan(p, H, seed=2)
dp = engine.DevicePlan(plan)
gt = torch.empty((n, p, 2), dtype=torch.uint8, device=dev)
_cuda.check(_cuda.lib().ht_synth_gt(gt.data_ptr(), *gt.stride(), n, p, 0, 3, 1, 0, 0), "synth")
out = torch.empty((n, 2 * H), dtype=torch.uint8, device=dev)
planes = engine.alloc_planes(n, plan.n_keys, dev)
def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
def whole():
    engine.pack(dp, gt.data_ptr(), gt.stride(), n, planes)
    engine.transform_u8(dp, planes, n, out.data_ptr(), 2 * H)
ms = timeit(whole); print(f"whole: {ms:.2f} ms -> {n*H*2/ms/1e6/1e3:.1f} G calls/s x1e3")
ref = out.clone()
for S in (1, 2, 3, 4, 8):
    rows = S * 1024
    pl = [engine.alloc_planes(rows, plan.n_keys, dev) for _ in range(2)]
    def tiled():
        for i, s0 in enumerate(range(0, n, rows)):
            m = min(rows, n - s0)
            b = pl[i & 1]
            engine.pack(dp, gt.data_ptr() + s0 * gt.stride(0), gt.stride(), m, b)
            engine.transform_u8(dp, b, m, out.data_ptr() + s0 * 2 * H, 2 * H)
    out.zero_()
    ms = timeit(tiled)
    print(f"S={S}: {ms:.2f} ms  equal={torch.equal(out, ref)}")
