"""Sample sharding + the optional all-gather of the packed result, world_size 2 on CPU (gloo)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from haptools_b200 import dist as hdist


@pytest.mark.parametrize("n,world", [(1, 1), (1024, 2), (2504, 2), (2504, 8), (500_000, 8), (5, 4)])
def test_shard_bounds_partition(n, world):
    bounds = [hdist.shard_bounds(n, world, r) for r in range(world)]
    assert bounds[0][0] == 0 and bounds[-1][1] == n
    for (a0, a1), (b0, b1) in zip(bounds, bounds[1:]):
        assert a1 == b0 and a0 <= a1
    for lo, hi in bounds[:-1]:
        assert (lo % 1024 == 0 or lo == n) and (hi % 1024 == 0 or hi == n)
    per = hdist.shard_tiles(n, world) * 1024
    assert all(hi - lo <= per for lo, hi in bounds)


def test_pack_unpack_roundtrip():
    rng = np.random.default_rng(0)
    out = rng.random((2504, 7, 2)) < 0.3
    bits = hdist.pack_result_bits(out)
    assert bits.shape == (3, 7, 32, 2) and bits.dtype == np.uint32
    assert np.array_equal(hdist.unpack_result_bits(bits, 2504), out)
    s, h, c = 1030, 3, 1
    assert bool((bits[s // 1024, h, (s % 1024) // 32, c] >> (s % 32)) & 1) == bool(out[s, h, c])


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, H, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(42)  # every rank can rebuild the cohort-wide result
        full = rng.random((n, H, 2)) < 0.4
        lo, hi = hdist.shard_bounds(n, world, rank)
        local = torch.from_numpy(hdist.pack_result_bits(full[lo:hi]).view(np.int32))
        gathered = hdist.allgather_bits(local, n)
        got = hdist.unpack_result_bits(gathered.numpy().view(np.uint32), n)
        ret[rank] = bool(np.array_equal(got, full))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [2504, 5000, 700])
def test_allgather_packed_result_gloo_world2(n):
    world, H = 2, 5
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), n, H, ret), nprocs=world, join=True)
        assert dict(ret) == {0: True, 1: True}
