"""Experiment: transform speed when haplotypes are position-sorted (as .hap files are) vs random order."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from haptools_b200 import engine, synth
from haptools_b200.plan import plan_from_refs
dev = torch.device('cuda')
n, p, H = 200_000, 50_000, 100_000
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ref_col, ref_allele, hap_off, _ = synth.synth_haplotype_refs(p, H, 2)
out = torch.empty((n, 2 * H), dtype=torch.uint8, device=dev)
for name in ("random order", "sorted by first variant"):
    if name.startswith("sorted"):
        k = np.diff(hap_off); first = ref_col[hap_off[:-1]]
        order = np.argsort(first, kind="stable")
        idx = np.concatenate([np.arange(hap_off[h], hap_off[h + 1]) for h in order])
        rc, ra = ref_col[idx], ref_allele[idx]; ho = np.concatenate([[0], np.cumsum(k[order])])
    else:
        rc, ra, ho = ref_col, ref_allele, hap_off
    plan = plan_from_refs(rc, ra, ho, p)
    dp = engine.DevicePlan(plan)
    planes = engine.alloc_planes(n, plan.n_keys, dev); planes.random_(0, 2**31 - 1)
    ms = timeit(lambda: engine.transform_u8(dp, planes, n, out.data_ptr(), 2 * H))
    print(f"{name:26s}: {ms:.2f} ms  out {n*2*H/ms/1e6:.0f} GB/s")
