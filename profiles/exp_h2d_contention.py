#!/usr/bin/env python
"""Round 2: does the input's H2D slow down because the result's widening hammers the same host memory, and do more
copy streams (more DMA reads in flight) win it back?  H2D of 1.64 GB page-locked (the input of 16,384 UKB samples)
alone, the host copy engine's download+widen of 3.28 GB alone, and both at once with the H2D split over 1 / 2 / 4
streams."""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from haptools_b200 import _cuda, engine  # noqa: E402

lib = _cuda.lib()
ne, p, H = 16384, 50_000, 100_000
src = engine.pinned_empty((ne, 2 * p))
src[:] = 1
dst = torch.empty((ne, 2 * p), dtype=torch.uint8, device="cuda")
pitch = int(lib.ht_rowbits_pitch(H))
rb = torch.randint(0, 255, (ne, pitch), dtype=torch.uint8, device="cuda")
out = np.empty((ne, 2 * H), dtype=np.uint8)
out[:] = 0
streams = [torch.cuda.Stream() for _ in range(4)]
res = {}


def h2d(k):
    rows = ne // k
    for i in range(k):
        with torch.cuda.stream(streams[i]):
            dst[i * rows:(i + 1) * rows].copy_(torch.from_numpy(src[i * rows:(i + 1) * rows]), non_blocking=True)


def run(name, k, expand):
    best = None
    for _ in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        t = lib.ht_host_submit_download_expand(out.ctypes.data, 2 * H, 2 * H, rb.data_ptr(), pitch, ne, 0) if expand else 0
        if k:
            h2d(k)
        t_h2d = None
        if k:
            for s in streams[:k]:
                s.synchronize()
            t_h2d = time.perf_counter() - t0
        if expand:
            _cuda.check(lib.ht_host_wait(t), "wait")
        dt = time.perf_counter() - t0
        if best is None or dt < best[0]:
            best = (dt, t_h2d)
    res[name] = {"ms_total": best[0] * 1e3, "ms_h2d": best[1] * 1e3 if best[1] else None,
                 "h2d_GBps": src.nbytes / best[1] / 1e9 if best[1] else None,
                 "widen_GBps_out": out.nbytes / best[0] / 1e9 if expand else None}
    print(name, res[name], flush=True)


run("h2d alone, 1 stream", 1, False)
run("h2d alone, 2 streams", 2, False)
run("download+widen alone", 0, True)
for k in (1, 2, 4):
    run(f"both, h2d on {k} stream(s)", k, True)
Path("gpurun_out").mkdir(exist_ok=True)
Path("gpurun_out/h2d_contention.json").write_text(json.dumps(res, indent=1))
