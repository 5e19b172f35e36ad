"""Variant-major pack (pack_vm_kernel): time per launch on the 1000G-scale shape and at 200 k samples.
Run with HT_VM_WAVES=1|2|4|8|16 (grid = that many waves of resident CTAs) or HT_LIB_PATH=<another build>;
`small` as the first argument: only the 1000G-scale shape with 20 launches (for ncu captures)."""
import sys, torch
sys.path.insert(0, ".")
from haptools_b200 import engine, synth, _cuda
dev = torch.device("cuda")
def timeit(fn, reps):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
CASES = ((2_504, 1_000_000, 10_000, 200), (200_000, 50_000, 100_000, 5))
if len(sys.argv) > 1 and sys.argv[1] == "small":
    CASES = ((2_504, 1_000_000, 10_000, 20),)
for n, p, H, reps in CASES:
    plan = synth.synth_plan(p, H, seed=1)
    dp = engine.DevicePlan(plan)
    planes = engine.alloc_planes(n, plan.n_keys, dev)
    gt = torch.empty((p, n, 2), dtype=torch.uint8, device=dev).permute(1, 0, 2)
    _cuda.check(_cuda.lib().ht_synth_gt(gt.data_ptr(), *gt.stride(), n, p, 0, 3, 1, 0, 0), "synth")
    ms = timeit(lambda: engine.pack(dp, gt.data_ptr(), gt.stride(), n, planes), reps)
    ab = plan.algorithmic_bytes(n)["pack"]
    print(f"n={n} p={p} H={H}: pack_vm {ms*1e3:.1f} us  {ab/ms/1e6:.0f} GB/s", flush=True)
    del gt, planes
