#!/usr/bin/env python
"""What page-locking a caller's (touched, pageable) array costs: cudaHostRegister / cudaHostUnregister of blocks of
several sizes, and the H2D rate out of a registered block -- could a pageable gts.data be registered block by
block ahead of its DMA instead of being staged through page-locked slots?"""
import ctypes, json, time, sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from haptools_b200 import _cuda
lib = _cuda.lib()
torch.cuda.init(); torch.zeros(1, device="cuda")
a = np.empty(4 << 30, dtype=np.uint8); a[:] = 1
dev = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
res = {}
for mb in (64, 256, 1024):
    nb = mb << 20
    t_reg, t_unreg, t_copy = [], [], []
    for rep in range(3):
        off = rep * nb
        p = a.ctypes.data + off
        t0 = time.perf_counter(); rc = lib.ht_host_register(ctypes.c_void_p(p), nb); t1 = time.perf_counter()
        assert rc == 0, lib.ht_last_error()
        src = torch.from_numpy(a[off:off + nb])
        torch.cuda.synchronize(); t2 = time.perf_counter(); dev[:nb].copy_(src, non_blocking=True); torch.cuda.synchronize(); t3 = time.perf_counter()
        t4 = time.perf_counter(); lib.ht_host_unregister(ctypes.c_void_p(p)); t5 = time.perf_counter()
        t_reg.append(t1 - t0); t_copy.append(t3 - t2); t_unreg.append(t5 - t4)
    res[f"{mb} MB"] = {"register_GBps": nb / min(t_reg) / 1e9, "unregister_GBps": nb / min(t_unreg) / 1e9, "h2d_GBps": nb / min(t_copy) / 1e9,
                       "register_ms": min(t_reg) * 1e3, "unregister_ms": min(t_unreg) * 1e3}
    print(mb, res[f"{mb} MB"], flush=True)
Path("gpurun_out/register.json").write_text(json.dumps(res, indent=1))
