"""
Ancestry pack at the BASELINE config's shape (100 k samples x 200 k variants x 50 k haplotypes, 5 populations):
register-load form (HT_PACK_ANC_KERNEL=line) against the TMA-fed form (tma) for several item orders
(HT_TMA_GROUP = column blocks per group, HT_TMA_CHUNK = consecutive items per CTA).  Results: gpurun_out/pack_anc_tma.json
"""
import json, os, sys, torch
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

n, p, H, pops = (int(x) for x in (sys.argv[1:5] + [100_000, 200_000, 50_000, 5][len(sys.argv) - 1:]))
plan = synth.synth_plan(p, H, seed=4, n_pops=pops)
dp = engine.DevicePlan(plan)
gt = torch.empty((n, p, 2), dtype=torch.uint8, device=dev)
_cuda.check(lib.ht_synth_gt(gt.data_ptr(), *gt.stride(), n, p, 0, 5, 1, 0, 0), "synth")
anc = torch.empty((n, p, 2), dtype=torch.uint8, device=dev)
_cuda.check(lib.ht_synth_ancestry(anc.data_ptr(), *anc.stride(), n, p, 0, 6, pops, 2000, 0), "synth_anc")
ab = plan.algorithmic_bytes(n)["pack"]
ref = None
out = {"shape": [n, p, H, pops], "keys": int(plan.n_keys), "algorithmic_GB": ab / 1e9, "runs": []}
configs = [("line", None, None)] + [("tma", g, c) for g in (64, 256) for c in (1, 4, 8)]
if os.environ.get("EXP_FULL"): configs = [("line", None, None)] + [("tma", g, c) for g in (0, 64, 16, 256) for c in (1, 2, 4)]
for form, g, c in configs:
    os.environ["HT_PACK_ANC_KERNEL"] = form
    for k, v in (("HT_TMA_GROUP", g), ("HT_TMA_CHUNK", c)):
        if v is None: os.environ.pop(k, None)
        else: os.environ[k] = str(v)
    planes = engine.alloc_planes(n, plan.n_keys, dev)
    planes.zero_()
    run = lambda: engine.pack(dp, gt.data_ptr(), gt.stride(), n, planes, anc.data_ptr(), anc.stride())
    ms = timeit(run)
    same = None
    if ref is None: ref = planes
    else:
        same = bool(torch.equal(ref, planes)); del planes
    r = {"form": form, "group": g, "chunk": c, "ms": ms, "GBps": ab / ms / 1e6, "frac_of_6531": ab / ms / 1e6 / 6531.3, "same_planes_as_line": same}
    out["runs"].append(r); print(r, flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/pack_anc_tma%s.json" % os.environ.get("EXP_TAG", ""), "w"), indent=1)
