#!/usr/bin/env python
"""Peer-to-peer bandwidth between two GPUs of the box from ONE process: DMA copy (torch copy_) one way and both
ways at once, and SM stores into peer memory (ht_transform_bits_gather with only the PEER as destination would
need IPC; here: a torch elementwise copy kernel writing into the peer's tensor)."""
import json, sys, time
from pathlib import Path
import torch
n = 4 << 30
a0 = torch.empty(n, dtype=torch.uint8, device="cuda:0"); a1 = torch.empty(n, dtype=torch.uint8, device="cuda:1")
b0 = torch.empty(n, dtype=torch.uint8, device="cuda:0"); b1 = torch.empty(n, dtype=torch.uint8, device="cuda:1")
res = {"can_access_peer": torch.cuda.can_device_access_peer(0, 1)}
def timed(fn, reps=3):
    fn(); torch.cuda.synchronize(0); torch.cuda.synchronize(1)
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(0); torch.cuda.synchronize(1); best = min(best, time.perf_counter() - t0)
    return best
t = timed(lambda: a1.copy_(a0, non_blocking=True)); res["dma 0->1 GBps"] = n / t / 1e9
s0, s1 = torch.cuda.Stream(0), torch.cuda.Stream(1)
def both():
    with torch.cuda.stream(s0): a1.copy_(a0, non_blocking=True)
    with torch.cuda.stream(s1): b0.copy_(b1, non_blocking=True)
t = timed(both); res["dma both ways GBps per direction"] = n / t / 1e9
# SM stores into peer memory: an elementwise kernel on device 0 whose output tensor lives on device 1
x0 = a0.view(torch.int32)
y1 = a1.view(torch.int32)
def sm_store():
    with torch.cuda.device(0):
        torch.add(x0, 1, out=y1) if False else y1.copy_(x0 + 0)
try:
    with torch.cuda.device(0):
        t = timed(lambda: torch.bitwise_not(x0, out=y1))
    res["sm stores 0->1 GBps (torch elementwise kernel, out= on the peer)"] = n / t / 1e9
except Exception as e:
    res["sm stores"] = f"{type(e).__name__}: {e}"[:200]
print(json.dumps(res, indent=1))
Path("gpurun_out/p2p.json").write_text(json.dumps(res, indent=1))
