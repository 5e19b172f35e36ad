#!/usr/bin/env python
"""Round 2: (1) transform_host(devices=[0..k-1]) on PAGEABLE arrays from ONE process, k = 1, 2, 4, 8 (the host copy
engine serves all devices at once: does a box's worth of GPUs help a pageable caller?); (2) the 1000G-scale
end-to-end call (variant-major host array, only referenced rows uploaded): row gather by the SMs over PCIe
(page-locked source) vs by the host copy engine's threads (HT_GATHER=bounce, or a pageable source)."""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from haptools_b200 import _cuda, engine, synth  # noqa: E402

res = {"gpus": torch.cuda.device_count(), "workers": _cuda.lib().ht_host_workers(), "cores": os.cpu_count()}


def best(fn, reps=3):
    fn()
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return min(ts)


# (1) UKB-scale plan, 32,768 samples, pageable source and a fresh (recycled) pageable result
p, H, n = 50_000, 100_000, 32_768
plan = synth.synth_plan(p, H, seed=2)
base = synth.synth_gt(256, p, 3, mode=1)
gt = np.empty((n, p, 2), dtype=np.uint8)
for s0 in range(0, n, 256):
    gt[s0:s0 + 256] = base
ref = None
k = 1
while k <= torch.cuda.device_count():
    devs = list(range(k))
    t = best(lambda: engine.transform_host(gt, plan, devices=devs))
    out = engine.transform_host(gt, plan, devices=devs)
    chk = int(out[:64].sum()) + int(out[-64:].sum())
    ref = chk if ref is None else ref
    res[f"pageable, {k} device(s)"] = {"ms": t * 1e3, "hapcalls_per_s": n * H * 2 / t, "same_result": chk == ref}
    print(f"pageable, {k} device(s)", res[f"pageable, {k} device(s)"], flush=True)
    del out
    k *= 2
del gt

# (2) 1000G-scale end to end
p, H, n = 1_000_000, 10_000, 2_504
plan = synth.synth_plan(p, H, seed=1)
pinned = engine.pinned_empty((p, n, 2))
pat = synth.synth_gt(n, 512, 2, mode=1).transpose(1, 0, 2)
for v0 in range(0, p, 512):
    pinned[v0:v0 + 512] = pat[: min(512, p - v0)]
pageable = np.array(pinned)
for name, arr, env in (("page-locked source, SM gather over PCIe", pinned, None), ("page-locked source, host copy engine gather", pinned, "bounce"),
                       ("pageable source, host copy engine gather", pageable, None)):
    if env:
        os.environ["HT_GATHER"] = env
    else:
        os.environ.pop("HT_GATHER", None)
    view = arr.transpose(1, 0, 2)
    t = best(lambda: engine.transform_host(view, plan), reps=5)
    res["1000g e2e: " + name] = {"ms": t * 1e3, "hapcalls_per_s": n * H * 2 / t, "upload_bytes": engine.upload_bytes(view, plan)}
    print("1000g e2e:", name, res["1000g e2e: " + name], flush=True)
Path("gpurun_out").mkdir(exist_ok=True)
Path(f"gpurun_out/host_multi_g{torch.cuda.device_count()}.json").write_text(json.dumps(res, indent=1))
