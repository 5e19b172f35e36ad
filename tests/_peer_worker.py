"""
Worker of tests/test_gpu_peer.py::test_two_processes_share_one_gpu -- run as
    python -m torch.distributed.run --nproc-per-node 2 tests/_peer_worker.py
Two ranks (gloo process group: both may sit on the same GPU, which NCCL refuses) shard a cohort, every rank packs its
shard and PeerGather.run() stores its records into BOTH ranks' gathered buffers through CUDA IPC peer memory; every
rank then checks its whole gathered buffer against the oracle's result for the whole cohort.
"""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent))


def main():
    from haptools_b200 import dist as hdist
    from haptools_b200 import engine
    from oracle import transform_oracle as orc
    from test_gpu_parity import _random_arrays

    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)) % torch.cuda.device_count())
    n, p, H = 5000, 60, 77  # 5 tiles over 2 ranks: 3 + 2, the last one ragged
    rng = np.random.default_rng(123)  # the same cohort on every rank
    gt, _, plan, want = _random_arrays(rng, n, p, H, max_k=4, missing=0.02)
    lo, hi = hdist.shard_bounds(n, world, rank)
    dplan = engine.DevicePlan(plan)
    shard = torch.from_numpy(np.ascontiguousarray(gt[lo:hi])).cuda()
    planes = engine.alloc_planes(hi - lo, plan.n_keys, "cuda")
    engine.pack(dplan, shard.data_ptr(), shard.stride(), hi - lo, planes)
    pg = hdist.PeerGather(n, H)
    for _ in range(2):  # the buffer is written again in place
        bits = pg.run(dplan, planes, hi - lo)
        torch.cuda.synchronize()
        dist.barrier()
    got = hdist.unpack_result_bits(bits.cpu().numpy().view(np.uint32), n)
    ok = bool(np.array_equal(got, want))
    dist.barrier()
    pg.close()
    flags = [None] * world
    dist.all_gather_object(flags, ok)
    if rank == 0:
        print("PEER_GATHER_OK" if all(flags) else f"PEER_GATHER_MISMATCH {flags}", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
