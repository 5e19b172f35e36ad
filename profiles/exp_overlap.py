#!/usr/bin/env python
"""Does the pack of the next sample chunk fit under the transform of the current one?  The pack is DRAM-bound
(6.8 TB/s, little L2->SM pressure), the transform L2->SM-bound at 59 % of the DRAM bandwidth.  UKB-scale plan,
N samples resident; sequential (one stream) against two streams with the cohort cut into C chunks."""
import json, os, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from haptools_b200 import _cuda, engine, synth
n, p, H = int(os.environ.get("N", 262_144)), 50_000, 100_000
dev = torch.device("cuda"); lib = _cuda.lib()
plan = synth.synth_plan(p, H, seed=2, sort=os.environ.get("SORT", "0") == "1")
dp = engine.DevicePlan(plan)
gt = torch.empty((n, p, 2), dtype=torch.uint8, device=dev)
_cuda.check(lib.ht_synth_gt(gt.data_ptr(), *gt.stride(), n, p, 0, 5, 1, 0, 0), "synth")
planes = engine.alloc_planes(n, plan.n_keys, dev)
pitch = engine.out_pitch(H)
out = torch.empty((n, pitch), dtype=torch.uint8, device=dev)
rec = plan.n_keys * 256  # bytes of planes per tile
def run(chunks, two_streams):
    s_main = torch.cuda.current_stream()
    sa, sb = (torch.cuda.Stream(), torch.cuda.Stream()) if two_streams else (s_main, s_main)
    if two_streams:
        sa.wait_stream(s_main); sb.wait_stream(s_main)
    per = (n // 1024 + chunks - 1) // chunks * 1024
    evs = []
    for c in range(chunks):
        s0, s1 = c * per, min(n, (c + 1) * per)
        if s0 >= s1: break
        pl = planes.view(torch.uint8)[(s0 // 1024) * rec:].view(planes.dtype)
        engine.pack(dp, gt.data_ptr() + s0 * gt.stride(0), gt.stride(), s1 - s0, pl, stream=sa)
        ev = sa.record_event(); evs.append(ev)
        sb.wait_event(ev)
        engine.transform_u8(dp, pl, s1 - s0, out.data_ptr() + s0 * pitch, pitch, stream=sb)
    if two_streams:
        s_main.wait_stream(sa); s_main.wait_stream(sb)
def timeit(fn, reps=4):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
res = {"n": n}
run(1, False); torch.cuda.synchronize(); ref = out.clone()
for chunks, two in ((1, False), (4, False), (2, True), (4, True), (8, True), (16, True), (32, True)):
    ms = timeit(lambda: run(chunks, two))
    same = bool(torch.equal(out, ref))
    res[f"chunks={chunks} two_streams={two}"] = {"ms": ms, "same": same}
    print(chunks, two, round(ms, 3), same, flush=True)
Path("gpurun_out/overlap_%s.json" % os.environ.get("SORT", "0")).write_text(json.dumps(res, indent=1))
