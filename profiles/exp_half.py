"""
Half-tile transform kernel (csrc/ht_transform_half.cu) against the tile form: bit-for-bit comparison on
ragged shapes (rows, haplotype groups, empty haplotypes, counts), then timings at n = 200 k of the
UKB-scale plan.  HT_TRANSFORM_KERNEL=tile forces the tile form inside the same library;
HT_LIB_PATH names a build variant.
"""
import os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from haptools_b200 import engine, synth
dev = torch.device("cuda")

def run(dp, planes, n, H, pitch, kernel, counts=False):
    os.environ["HT_TRANSFORM_KERNEL"] = kernel
    out = torch.full((n, pitch), 7, dtype=torch.uint8, device=dev)
    cnt = torch.zeros(H, dtype=torch.int64, device=dev) if counts else None
    engine.transform_u8(dp, planes, n, out.data_ptr(), pitch, method="direct", counts=cnt)
    torch.cuda.synchronize()
    return out, cnt

bad = 0
for (n, p, H, kw) in ((1, 64, 1, {}), (513, 300, 129, {}), (1500, 300, 250, dict(min_alleles=0, max_alleles=5)),
                      (4096 + 37, 2000, 1000, {}), (2048, 500, 136, dict(min_alleles=20, max_alleles=60)),
                      (3000, 4000, 128 * 3 + 5, dict(min_alleles=0, max_alleles=300))):
    try:
        plan = synth.synth_plan(p, H, seed=n, **kw)
    except Exception as e:  # window too small for the longest haplotypes
        print("skip", n, p, H, kw, e); continue
    dp = engine.DevicePlan(plan)
    planes = engine.alloc_planes(n, plan.n_keys, dev)
    planes.random_(-2**31, 2**31 - 1)
    for pitch in (engine.out_pitch(H), engine.out_pitch(H) + 32):
        a, ca = run(dp, planes, n, H, pitch, "tile", counts=True)
        b, cb = run(dp, planes, n, H, pitch, "half", counts=True)
        same = torch.equal(a, b) and torch.equal(ca, cb)
        bad += not same
        print(f"n={n} H={H} keys={plan.n_keys} K={len(plan.hap_key)} pitch={pitch}: {'same' if same else 'DIFFERENT'}"
              + ("" if same else f" ({int((a != b).sum())} bytes, counts {int((ca != cb).sum())})"))
print("PARITY", "ok" if bad == 0 else f"FAILED ({bad})")

def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

n, p, H = 200_000, 50_000, 100_000
out = torch.empty((n, 2 * H), dtype=torch.uint8, device=dev)
for name, kw in (("2-20", {}), ("20", dict(min_alleles=20, max_alleles=20)), ("0 (store only)", dict(min_alleles=0, max_alleles=0)),
                 ("2-20 sorted", dict(sort=True))):
    plan = synth.synth_plan(p, H, seed=2, **kw)
    dp = engine.DevicePlan(plan)
    planes = engine.alloc_planes(n, max(plan.n_keys, 1), dev); planes.random_(0, 2**31 - 1)
    res = []
    for kernel in ("tile", "half"):
        os.environ["HT_TRANSFORM_KERNEL"] = kernel
        ms = timeit(lambda: engine.transform_u8(dp, planes, n, out.data_ptr(), 2 * H, method="direct"))
        res.append(f"{kernel} {ms:.2f} ms")
    print(os.environ.get("HT_LIB_PATH", "default"), name, " | ".join(res), flush=True)
