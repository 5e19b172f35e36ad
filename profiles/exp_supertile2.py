"""Experiment: 2-stream pipeline over super-tiles: pack(i+1) overlaps transform(i); plane ring of 3."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from haptools_b200 import engine, synth, _cuda
dev = torch.device('cuda')
n, p, H = 200_000, 50_000, 100_000
plan = synth.synth_plan(p, H, seed=2)
dp = engine.DevicePlan(plan)
gt = torch.empty((n, p, 2), dtype=torch.uint8, device=dev)
_cuda.check(_cuda.lib().ht_synth_gt(gt.data_ptr(), *gt.stride(), n, p, 0, 3, 1, 0, 0), "synth")
out = torch.empty((n, 2 * H), dtype=torch.uint8, device=dev)
planes = engine.alloc_planes(n, plan.n_keys, dev)
def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
def whole():
    engine.pack(dp, gt.data_ptr(), gt.stride(), n, planes)
    engine.transform_u8(dp, planes, n, out.data_ptr(), 2 * H)
ms = timeit(whole); print(f"whole: {ms:.2f} ms")
ref = out.clone()
for S, prio in ((8, 0), (16, 0), (32, 0), (8, -1), (16, -1), (32, -1), (16, 1), (32, 1)):
    # prio -1: the pack stream has priority (its CTAs are scheduled first whenever a slot frees); +1: the transform stream
    sp, st = torch.cuda.Stream(priority=min(prio, 0)), torch.cuda.Stream(priority=min(-prio, 0))
    rows = S * 1024
    NB = 3
    pl = [engine.alloc_planes(rows, plan.n_keys, dev) for _ in range(NB)]
    def piped():
        cur = torch.cuda.current_stream()
        sp.wait_stream(cur); st.wait_stream(cur)
        ev_p = [None] * NB; ev_t = [None] * NB
        for i, s0 in enumerate(range(0, n, rows)):
            m = min(rows, n - s0); b = i % NB
            if ev_t[b] is not None: sp.wait_event(ev_t[b])
            engine.pack(dp, gt.data_ptr() + s0 * gt.stride(0), gt.stride(), m, pl[b], stream=sp)
            ev_p[b] = sp.record_event()
            st.wait_event(ev_p[b])
            engine.transform_u8(dp, pl[b], m, out.data_ptr() + s0 * 2 * H, 2 * H, stream=st)
            ev_t[b] = st.record_event()
        cur.wait_stream(sp); cur.wait_stream(st)
    out.zero_()
    ms = timeit(piped)
    print(f"S={S} prio={prio}: {ms:.2f} ms  equal={torch.equal(out, ref)}", flush=True)
