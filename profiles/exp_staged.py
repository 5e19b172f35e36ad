#!/usr/bin/env python
"""Round 2: the shared-memory staged transform (ht_transform_u8_staged) on the UKB-scale plan with position-sorted
haplotypes, against the L2 path, for several stage budgets; and on the unsorted plan (must not get slower).
HT_LIB_PATH selects a build variant (profiles/variants/)."""
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from haptools_b200 import _cuda, engine, synth  # noqa: E402

n = int(os.environ.get("N", 200_000))
p, H = 50_000, 100_000
dev = torch.device("cuda")
res = {"n": n, "lib": os.environ.get("HT_LIB_PATH", "product")}


def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


out = torch.empty((n, engine.out_pitch(H)), dtype=torch.uint8, device=dev)
for sort in (True, False):
    plan = synth.synth_plan(p, H, seed=2, sort=sort)
    planes = engine.alloc_planes(n, plan.n_keys, dev)
    planes.random_(-(2**31), 2**31 - 1)
    lo, cnt = plan.block_spans()
    ref = None
    for lines, hpc in ([(0, 1), (0, 2), (256, 1), (256, 2), (288, 2), (352, 2)] if sort else [(0, 1), (0, 2)]):
        os.environ["HT_HALF_PER_CTA"] = str(hpc)
        dplan = engine.DevicePlan(plan, dev)
        dplan.stage_lines = lines
        if lines:
            dplan.blk_key_lo = torch.from_numpy(lo.view(np.int32)).to(dev)
            dplan.blk_key_cnt = torch.from_numpy(cnt.view(np.int32)).to(dev)
        method = "staged" if lines else "direct"
        ms = timeit(lambda: engine.transform_u8(dplan, planes, n, out.data_ptr(), out.stride(0), method=method))
        chk = int(out[:4096].sum().item()) + int(out[-4096:].sum().item())
        ref = chk if ref is None else ref
        ab = plan.algorithmic_bytes(n)["transform"]
        key = f"{'sorted' if sort else 'unsorted'} stage_lines={lines} half_tiles_per_cta={hpc}"
        res[key] = {"ms": ms, "GBps": ab / ms / 1e6, "frac_of_6531": ab / ms / 1e6 / 6531.3, "same_result": chk == ref,
                    "staged_block_fraction": float(((cnt > 0) & (cnt <= lines)).mean()) if lines else 0.0,
                    "ms_at_500k": ms * 500_000 / n}
        print(key, res[key], flush=True)
    del planes
Path("gpurun_out").mkdir(exist_ok=True)
tag = Path(os.environ.get("HT_LIB_PATH", "product")).stem
Path(f"gpurun_out/staged_{tag}.json").write_text(json.dumps(res, indent=1))
