"""Experiment: one-kernel transform vs the two-phase local form (phase 1 / phase 2 apart)."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from haptools_b200 import engine, synth, _cuda
dev = torch.device('cuda')
n, p, H = int(os.environ.get("N", 200_000)), 50_000, 100_000
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
plan = synth.synth_plan(p, H, seed=2)
out = torch.empty((n, 2 * H), dtype=torch.uint8, device=dev)
planes = engine.alloc_planes(n, plan.n_keys, dev); planes.random_(0, 2**31 - 1)
tiles = (n + 1023) // 1024
for dmax, mb in ((224, 128), (288, 128)):
    dp = engine.DevicePlan(plan, local_dmax=dmax, local_max_block=mb)
    lt, _ = dp.local()
    if dmax == 224:
        ms = timeit(lambda: engine.transform_u8(dp, planes, n, out.data_ptr(), 2 * H, method="direct"))
        print(f"direct: {ms:.2f} ms  out {n*2*H/ms/1e6:.0f} GB/s")
        ref = out.clone() if n <= 200_000 else None
    for g in (64, 16, tiles):
        ws = torch.empty(min(g, tiles) * H * 64, dtype=torch.int32, device=dev)
        ms = timeit(lambda: engine.transform_u8(dp, planes, n, out.data_ptr(), 2 * H, method="local", workspace=ws))
        print(f"local dmax={dmax} blocks={lt.n_blocks} staged_reads={lt.staged_reads} group={g}: {ms:.2f} ms  out {n*2*H/ms/1e6:.0f} GB/s", "same" if ref is None else bool(torch.equal(ref, out)))
        del ws
    bits = torch.empty((tiles, H, 32, 2), dtype=torch.int32, device=dev)
    ms1 = timeit(lambda: engine.transform_bits(dp, planes, n, bits, method="local"))
    ms2 = timeit(lambda: engine.unpack_bits_u8(bits, n, H, out.data_ptr(), 2 * H))
    print(f"   phase 1 alone {ms1:.2f} ms, phase 2 alone {ms2:.2f} ms")
    del bits
ms = timeit(lambda: engine.transform_bits(dp, planes, n, torch.empty((tiles, H, 32, 2), dtype=torch.int32, device=dev), method="direct"))
print(f"bits direct {ms:.2f} ms")
