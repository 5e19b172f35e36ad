#!/usr/bin/env python
"""The FIRST Haplotypes.transform of a process (what `haptools transform` is): where its time goes.  UKB-scale plan,
65,536 samples, page-locked gts.data; each stage timed by running the pieces by hand in a fresh process, then the
real call in another fresh process (argv[1] = "pieces" | "call")."""
import json, os, sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
t_imp = time.perf_counter()
from haptools_b200 import _cuda, engine, synth, integrate
from haptools_b200.plan import plan_from_haplotypes
ne, p, H = int(os.environ.get("E2E_SAMPLES", 65536)), 50_000, 100_000
haps, gts = synth.synth_objects(p, H, 2)
base = synth.synth_gt(256, p, 3, mode=1)
res = {}
def T(name, fn):
    torch.cuda.synchronize() if torch.cuda.is_initialized() else None
    t0 = time.perf_counter(); r = fn()
    if torch.cuda.is_initialized(): torch.cuda.synchronize()
    res[name] = (time.perf_counter() - t0) * 1e3; print(f"{name}: {res[name]:.1f} ms", flush=True); return r
T("cuda context (torch.zeros on the device)", lambda: torch.zeros(1, device="cuda"))
gt = T("page-locked gts.data (6.5 GB: not part of the call)", lambda: engine.pinned_empty((ne, p, 2)))
gt[:] = base[np.arange(ne) % 256]
gts.samples = tuple(map(str, range(ne))); gts.data = gt
if sys.argv[1] == "pieces":
    hap_list = list(haps.data.values())
    plan = T("planner (plan_from_haplotypes)", lambda: plan_from_haplotypes(hap_list, gts, ancestry=False))
    T("quad_tables + block_spans + fast_tables (numpy)", lambda: (plan.quad_tables(), plan.block_spans(), plan.fast_tables()))
    dplan = T("DevicePlan (table upload)", lambda: engine.device_plan(plan, torch.device("cuda", 0)))
    T("library load + host copy engine start (first tiny transform_host)", lambda: engine.transform_host(gt[:1024], plan, dplan=dplan))
    out = T("result array np.empty (13 GB, untouched)", lambda: np.empty((ne, H, 2), dtype=np.uint8))
    T("transform_host into the fresh result (first touch of 13 GB)", lambda: engine.transform_host(gt, plan, dplan=dplan, out=out))
    T("transform_host again (pages touched)", lambda: engine.transform_host(gt, plan, dplan=dplan, out=out))
    T("hap_gts.variants", lambda: integrate._result_variants(haps, hap_list, np.dtype([("id", "U50"), ("chrom", "U10"), ("pos", np.uint32), ("alleles", object)])))
else:
    r = T("first Haplotypes.transform of the process", lambda: haps.transform(gts))
    del r
    r = T("second call", lambda: haps.transform(gts))
Path(f"gpurun_out/cold_call_{sys.argv[1]}.json").write_text(json.dumps(res, indent=1))
