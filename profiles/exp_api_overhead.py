#!/usr/bin/env python
"""Where the class API's time beyond the engine's goes (UKB-scale plan, 65,536 samples, page-locked gts.data):
engine with a reused page-locked result / engine with the recycled pageable result the API uses / the class API
with and without the helper thread's work."""
import json, os, sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from haptools_b200 import _cuda, engine, synth, integrate
ne, p, H = int(os.environ.get("E2E_SAMPLES", 65536)), 50_000, 100_000
haps, gts = synth.synth_objects(p, H, 2)
base = synth.synth_gt(256, p, 3, mode=1)
gt = engine.pinned_empty((ne, p, 2)); gt[:] = base[np.arange(ne) % 256]
gts.samples = tuple(map(str, range(ne))); gts.data = gt
def timed(fn, reps=4):
    ts = []
    for _ in range(reps + 1):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0); del r
    return {"ms_best": min(ts[1:]) * 1e3, "ms_mean": float(np.mean(ts[1:])) * 1e3}
res = {}
r0 = haps.transform(gts); del r0
plan = haps.__dict__[integrate._CACHE_ATTR][2]
dplan = engine.device_plan(plan, torch.device("cuda", 0))
out_pinned = engine.pinned_empty((ne, H, 2))
res["engine, reused page-locked out"] = timed(lambda: engine.transform_host(gt, plan, out=out_pinned, dplan=dplan))
res["engine, out=None (recycled pageable)"] = timed(lambda: engine.transform_host(gt, plan))
res["class API"] = timed(lambda: haps.transform(gts))
orig = integrate._result_variants
cached = orig(haps, list(haps.data.values()), r0.variants.dtype) if False else None
integrate._result_variants = lambda haps_obj, hl, dtype: np.empty(0, dtype=dtype)
res["class API, variants array not built"] = timed(lambda: haps.transform(gts))
integrate._result_variants = orig
import sys as _s
for si in (5e-3, 1e-3, 2e-4, 5e-5):
    class _S(integrate._ShortSwitchInterval):
        def __enter__(self, si=si):
            if self._active:
                self._old = _s.getswitchinterval(); _s.setswitchinterval(si)
            return self
    integrate._ShortSwitchInterval, keep = _S, integrate._ShortSwitchInterval
    res[f"class API, switch interval {si}"] = timed(lambda: haps.transform(gts))
    integrate._ShortSwitchInterval = keep
for k, v in res.items(): print(k, v, flush=True)
Path("gpurun_out/api_overhead.json").write_text(json.dumps(res, indent=1))
