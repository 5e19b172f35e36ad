#!/usr/bin/env python
"""Row gather out of page-locked host memory over PCIe (ht_copy_rows_h2d, gather_rows_kernel): the 1000G-scale
upload (104 k of 1 M variant rows of 5,008 bytes) for several grids and loads in flight per lane, next to a
plain DMA of the same number of bytes."""
import json, os, sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from haptools_b200 import _cuda, engine, synth
lib = _cuda.lib()
p, H, n = 1_000_000, 10_000, 2_504
plan = synth.synth_plan(p, H, seed=1)
rows = np.ascontiguousarray(plan.var_col, dtype=np.uint32)
width = 2 * n
pinned = engine.pinned_empty((p, width))
pinned[:] = 1
dst = torch.empty((len(rows), (width + 15) // 16 * 16), dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream()
def run():
    _cuda.check(lib.ht_copy_rows_h2d(dst.data_ptr(), dst.stride(0), pinned.ctypes.data, width, width, rows.ctypes.data, len(rows), st.cuda_stream), "rows")
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(reps):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
nbytes = len(rows) * width
out = {"rows": int(len(rows)), "row_bytes": width, "MB": nbytes / 1e6, "runs": []}
for unroll in (4, 8, 16):
    for ctas in (148, 296, 592, 1184, 2368):
        os.environ["HT_GATHER_UNROLL"], os.environ["HT_GATHER_CTAS"] = str(unroll), str(ctas)
        ms = timeit(run)
        r = {"unroll": unroll, "ctas": ctas, "ms": ms, "GBps": nbytes / ms / 1e6}
        out["runs"].append(r); print(r, flush=True)
flat = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
src = torch.from_numpy(pinned.reshape(-1)[:nbytes])
ms = timeit(lambda: flat.copy_(src, non_blocking=True))
out["dma_same_bytes"] = {"ms": ms, "GBps": nbytes / ms / 1e6}; print(out["dma_same_bytes"])
json.dump(out, open("gpurun_out/gather.json", "w"), indent=1)
