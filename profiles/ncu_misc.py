#!/usr/bin/env python
"""Launches for ncu of the round-2 kernels outside the default bench: ht_transform_bits_gather (a world of one rank:
its own buffer is the only destination) at 65,536 samples of the UKB-scale plan, and ht_pack_i32 (narrow + variant-major
pack) on one 1,024-variant chunk of 200 k samples."""
import sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from haptools_b200 import _cuda, engine, synth
from haptools_b200 import dist as hdist
lib = _cuda.lib()
n = 65536
plan = synth.synth_plan(50_000, 100_000, seed=2)
dplan = engine.DevicePlan(plan)
planes = engine.alloc_planes(n, plan.n_keys, "cuda"); planes.random_(-(2**31), 2**31 - 1)
pg = hdist.PeerGather(n, plan.n_haps)
for _ in range(3):
    pg.run(dplan, planes, n)
torch.cuda.synchronize(); pg.close(); del planes
# one int32 chunk
ns, nv = 200_000, 1024
p2 = synth.synth_plan(nv, 100, seed=3)
al = torch.randint(0, 2, (nv, 2 * ns), dtype=torch.int32, device="cuda")
pl = engine.alloc_planes(ns, p2.n_keys, "cuda")
rows = torch.from_numpy(p2.var_col.astype(np.int32)).cuda()
koff = torch.from_numpy(p2.var_key_off.astype(np.int32)).cuda()
kall = torch.from_numpy(p2.key_allele.copy()).cuda()
for _ in range(3):
    _cuda.check(lib.ht_pack_i32(al.data_ptr(), 2 * ns, ns, rows.data_ptr(), koff.data_ptr(), kall.data_ptr(), p2.n_vars, 0, p2.n_keys,
                                pl.data_ptr(), 0, 0), "ht_pack_i32")
torch.cuda.synchronize()
print("done", p2.n_vars, "referenced variants of", nv)
