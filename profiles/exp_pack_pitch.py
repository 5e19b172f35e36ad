"""
Sample-major pack kernels as a function of the row pitch and of the number of input arrays: is the
ancestry pack's 0.68-0.73 of the HBM peak a property of its access pattern (two arrays, 400 KB between
consecutive rows) or of the kernel?  Plain pack (one array) at p = 50 k / 200 k variants with the same
bytes, ancestry pack (two arrays) at the same shapes, both forms (HT_PACK_ANC_KERNEL=warp|line).
"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from haptools_b200 import _cuda, engine, synth
dev = torch.device("cuda"); lib = _cuda.lib()

def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

for (n, p, H, pops) in ((400_000, 50_000, 25_000, 0), (100_000, 200_000, 100_000, 0), (400_000, 50_000, 12_500, 5), (100_000, 200_000, 50_000, 5)):
    plan = synth.synth_plan(p, H, seed=4, n_pops=pops)
    dp = engine.DevicePlan(plan)
    gt = torch.empty((n, p, 2), dtype=torch.uint8, device=dev)
    _cuda.check(lib.ht_synth_gt(gt.data_ptr(), *gt.stride(), n, p, 0, 5, 1, 0, 0), "synth")
    anc = None
    if pops:
        anc = torch.empty((n, p, 2), dtype=torch.uint8, device=dev)
        _cuda.check(lib.ht_synth_ancestry(anc.data_ptr(), *anc.stride(), n, p, 0, 6, pops, 2000, 0), "synth_anc")
    planes = engine.alloc_planes(n, plan.n_keys, dev)
    ab = plan.algorithmic_bytes(n)["pack"]
    ms = timeit(lambda: engine.pack(dp, gt.data_ptr(), gt.stride(), n, planes, anc.data_ptr() if pops else 0, anc.stride() if pops else (0, 0, 0)))
    print(f"{os.environ.get('HT_PACK_ANC_KERNEL', 'warp')}: n={n} p={p} keys={plan.n_keys} pops={pops}: {ms:.2f} ms, {ab/1e9:.1f} GB algorithmic, {ab/ms/1e6:.0f} GB/s", flush=True)
    del gt, anc, planes
    torch.cuda.empty_cache()
