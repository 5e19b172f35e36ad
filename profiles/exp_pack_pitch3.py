"""Ancestry pack (line form) with the tile mapping (CTA = 1,024 rows x 64 columns) and the slab mapping
(128 rows x 512 columns) as a function of the row pitch.  HT_PACK_SM2_KERNEL=tile|slab."""
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
for (n, p, H) in ((200_000, 50_000, 12_500), (50_000, 200_000, 50_000), (20_000, 500_000, 125_000), (10_000, 1_000_000, 250_000)):
    plan = synth.synth_plan(p, H, seed=4, n_pops=5)
    dp = engine.DevicePlan(plan)
    gt = torch.empty((n, p, 2), dtype=torch.uint8, device=dev)
    anc = torch.empty((n, p, 2), dtype=torch.uint8, device=dev)
    _cuda.check(lib.ht_synth_gt(gt.data_ptr(), *gt.stride(), n, p, 0, 5, 1, 0, 0), "synth")
    _cuda.check(lib.ht_synth_ancestry(anc.data_ptr(), *anc.stride(), n, p, 0, 6, 5, 2000, 0), "synth_anc")
    planes = engine.alloc_planes(n, plan.n_keys, dev)
    ab = plan.algorithmic_bytes(n)["pack"]
    ms = timeit(lambda: engine.pack(dp, gt.data_ptr(), gt.stride(), n, planes, anc.data_ptr(), anc.stride()))
    print(f"{os.environ.get('HT_PACK_SM2_KERNEL', 'default')}: n={n} p={p} pitch={2*p/1e3:.0f} KB keys={plan.n_keys}: {ms:.2f} ms, {ab/1e9:.1f} GB, {ab/ms/1e6:.0f} GB/s", flush=True)
    del gt, anc, planes; torch.cuda.empty_cache()
