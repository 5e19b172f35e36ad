"""Bit-packed transform (ht_transform_bits / ht_transform_bits_local) at n = 200 k of the UKB-scale plan,
next to the uint8 form: what the writers' path (VCF text, PGEN batches, all-gather) costs."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from haptools_b200 import engine, synth
dev = torch.device("cuda")
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
n, p, H = 200_000, 50_000, 100_000
tiles = (n + 1023) // 1024
for name, kw in (("2-20", {}), ("2-20 sorted", dict(sort=True))):
    plan = synth.synth_plan(p, H, seed=2, **kw)
    dp = engine.DevicePlan(plan)
    planes = engine.alloc_planes(n, plan.n_keys, dev); planes.random_(-2**31, 2**31 - 1)
    bits = torch.empty((tiles, H, 32, 2), dtype=torch.int32, device=dev)
    out = torch.empty((n, 2 * H), dtype=torch.uint8, device=dev)
    res = [f"u8 {timeit(lambda: engine.transform_u8(dp, planes, n, out.data_ptr(), 2 * H, method='direct')):.2f} ms"]
    for m in ("direct", "local"):
        res.append(f"bits/{m} {timeit(lambda: engine.transform_bits(dp, planes, n, bits, method=m)):.2f} ms")
    K = len(plan.hap_key)
    print(name, " | ".join(res), f"| record reads {K * tiles * 256 / 1e9:.1f} GB, bits {bits.numel() * 4 / 1e9:.1f} GB", flush=True)
