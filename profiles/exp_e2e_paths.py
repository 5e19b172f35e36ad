#!/usr/bin/env python
"""End-to-end paths of transform_host on the UKB-scale plan (round 2): how the result leaves the device
("u8" bytes by DMA vs "bits" + host widening), page-locked vs pageable source / result, fresh vs reused result.
Run once per HT_HOST_COPY_THREADS value (the pool is sized at first use)."""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from haptools_b200 import _cuda, engine, synth  # noqa: E402

ne = int(os.environ.get("E2E_SAMPLES", 16384))
p, H = 50_000, 100_000
plan = synth.synth_plan(p, H, seed=2)
dplan = engine.DevicePlan(plan)
lib = _cuda.lib()
base = synth.synth_gt(256, p, 3, mode=1)
pinned_in = engine.pinned_empty((ne, p, 2))
pinned_in[:] = base[np.arange(ne) % 256]
pageable_in = np.array(pinned_in)
pinned_out = engine.pinned_empty((ne, H, 2))
pageable_out = np.empty((ne, H, 2), dtype=np.uint8)
pageable_out[:] = 0
res = {"workers": lib.ht_host_workers(), "isa": lib.ht_host_expand_isa(-1), "samples": ne, "cores": os.cpu_count()}
ref = None


def run(name, gt, out, result, reps=4, **kw):
    global ref
    times = []
    for i in range(reps + 1):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = engine.transform_host(gt, plan, out=out() if callable(out) else out, dplan=dplan, result=result, **kw)
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    if ref is None:
        ref = r[:64].copy(), r[-64:].copy()
    ok = bool(np.array_equal(r[:64], ref[0]) and np.array_equal(r[-64:], ref[1]))
    res[name] = {"ms_first": times[0] * 1e3, "ms_best": min(times[1:]) * 1e3, "ms_mean": float(np.mean(times[1:])) * 1e3,
                 "hapcalls_per_s": ne * H * 2 / min(times[1:]), "same_result": ok}
    print(name, res[name], flush=True)


for result in ("u8", "bits"):
    run(f"{result}: pinned in, pinned out (reused)", pinned_in, pinned_out, result)
    run(f"{result}: pinned in, pageable out (reused)", pinned_in, pageable_out, result)
    run(f"{result}: pinned in, fresh np.empty out", pinned_in, None, result)
    run(f"{result}: pageable in, pageable out (reused)", pageable_in, pageable_out, result)
    run(f"{result}: pageable in, fresh np.empty out", pageable_in, None, result)
for chunk in (1024, 4096):
    run(f"bits: pinned in, pinned out, chunk {chunk}", pinned_in, pinned_out, "bits", chunk_samples=chunk)
# the pieces alone: host widening of a device-resident row-bit matrix, and a pool upload
pitch = int(lib.ht_rowbits_pitch(H))
rb = torch.randint(0, 255, (ne, pitch), dtype=torch.uint8, device="cuda")
for name, out in (("pinned", pinned_out), ("pageable", pageable_out)):
    best = 1e9
    for _ in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        t = lib.ht_host_submit_download_expand(out.ctypes.data, 2 * H, 2 * H, rb.data_ptr(), pitch, ne, 0)
        _cuda.check(lib.ht_host_wait(t), "wait")
        best = min(best, time.perf_counter() - t0)
    res[f"download_expand alone -> {name}"] = {"ms": best * 1e3, "GBps_out": ne * 2 * H / best / 1e9}
    print(f"download_expand alone -> {name}", res[f"download_expand alone -> {name}"], flush=True)
dbuf = torch.empty((ne, 2 * p), dtype=torch.uint8, device="cuda")
best = 1e9
for _ in range(4):
    t0 = time.perf_counter()
    t = lib.ht_host_submit_upload(dbuf.data_ptr(), 2 * p, pageable_in.ctypes.data, 2 * p, 2 * p, ne, None, 0)
    _cuda.check(lib.ht_host_wait(t), "wait")
    best = min(best, time.perf_counter() - t0)
res["pool upload alone (pageable)"] = {"ms": best * 1e3, "GBps": ne * 2 * p / best / 1e9}
print("pool upload alone", res["pool upload alone (pageable)"], flush=True)
out_path = Path("gpurun_out") / f"e2e_paths_t{lib.ht_host_workers()}.json"
out_path.parent.mkdir(exist_ok=True)
out_path.write_text(json.dumps(res, indent=1))
