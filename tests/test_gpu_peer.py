"""
ht_transform_bits_gather / haptools_b200.dist.PeerGather: the bit-packed transform that stores its records into every
rank's gathered buffer (compute + all-gather in one kernel over peer memory).  On one GPU: one destination = 
ht_transform_bits, several destinations and a tile offset through the raw ABI, and two PROCESSES sharing the GPU
through CUDA IPC handles (tests/_peer_worker.py) -- the same code path the ranks of an NVLink box take.
"""
import ctypes
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from haptools_b200 import _cuda, engine  # noqa: E402
from haptools_b200 import dist as hdist  # noqa: E402
from test_gpu_parity import _random_arrays  # noqa: E402


def _packed(rng, n, p, H, **kw):
    gt, _, plan, want = _random_arrays(rng, n, p, H, **kw)
    dplan = engine.DevicePlan(plan)
    g = torch.from_numpy(gt).cuda()
    planes = engine.alloc_planes(n, plan.n_keys, "cuda")
    engine.pack(dplan, g.data_ptr(), g.stride(), n, planes)
    return plan, dplan, planes, want


@pytest.mark.parametrize("n,H", [(1, 1), (1000, 33), (2504, 130), (4097, 9)])
def test_world_of_one_equals_transform_bits(n, H):
    rng = np.random.default_rng(n + H)
    plan, dplan, planes, want = _packed(rng, n, 50, H, max_k=5, missing=0.01)
    pg = hdist.PeerGather(n, H)
    try:
        bits = pg.run(dplan, planes, n)
        torch.cuda.synchronize()
        first = bits.clone()
        bits = pg.run(dplan, planes, n, ordered=True)  # haplotypes computed in key order: the same records
        torch.cuda.synchronize()
        assert torch.equal(bits, first)
        ref = torch.empty_like(bits)
        engine.transform_bits(dplan, planes, n, ref, method="direct")
        torch.cuda.synchronize()
        assert torch.equal(bits, ref)
        np.testing.assert_array_equal(hdist.unpack_result_bits(bits.cpu().numpy().view(np.uint32), n), want)
    finally:
        pg.close()


def test_several_destinations_and_tile_offset():
    lib = _cuda.lib()
    rng = np.random.default_rng(8)
    n, H = 2100, 40  # 3 tiles
    plan, dplan, planes, want = _packed(rng, n, 30, H, max_k=3)
    ref = torch.empty((3, H, 32, 2), dtype=torch.int32, device="cuda")
    engine.transform_bits(dplan, planes, n, ref, method="direct")
    bufs = [torch.full((6, H, 32, 2), 0x5A5A5A5A, dtype=torch.int32, device="cuda") for _ in range(3)]
    dsts = (ctypes.c_void_p * 3)(*[b.data_ptr() for b in bufs])
    order = hdist.PeerGather.key_order(dplan)
    assert sorted(order.cpu().tolist()) == list(range(H))
    for hap_order in (None, order.data_ptr()):  # caller order (one 2 KB bulk store per CTA) / key order (one per warp)
        for b in bufs:
            b.fill_(0x5A5A5A5A)
        _cuda.check(lib.ht_transform_bits_gather(planes.data_ptr(), n, plan.n_keys, dplan.hap_off.data_ptr(), dplan.hap_key.data_ptr(),
                                                 hap_order, H, dsts, 3, 2, 0), "ht_transform_bits_gather")
        torch.cuda.synchronize()
        for b in bufs:
            assert torch.equal(b[2:5], ref)
            assert bool((b[:2] == 0x5A5A5A5A).all()) and bool((b[5:] == 0x5A5A5A5A).all())
    # argument errors
    assert lib.ht_transform_bits_gather(planes.data_ptr(), n, plan.n_keys, dplan.hap_off.data_ptr(), dplan.hap_key.data_ptr(),
                                        None, H, dsts, 0, 0, 0) == _cuda.HT_ERR_BAD_ARG
    assert lib.ht_transform_bits_gather(planes.data_ptr(), n, plan.n_keys, dplan.hap_off.data_ptr(), dplan.hap_key.data_ptr(),
                                        None, H, dsts, 17, 0, 0) == _cuda.HT_ERR_BAD_ARG
    assert lib.ht_transform_bits_gather(planes.data_ptr(), n, plan.n_keys, dplan.hap_off.data_ptr(), dplan.hap_key.data_ptr(),
                                        None, H, dsts, 3, -1, 0) == _cuda.HT_ERR_BAD_ARG


def test_two_processes_share_one_gpu():
    """Two ranks, CUDA IPC handles exchanged through the process group, each storing its shard into both buffers."""
    import socket

    repo = Path(__file__).resolve().parent.parent
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", PYTHONPATH=str(repo))
    with socket.socket() as sock:  # a port nobody is using right now
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(repo / "tests" / "_peer_worker.py")]
    res = subprocess.run(cmd, env=env, cwd=str(repo), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert res.returncode == 0 and "PEER_GATHER_OK" in res.stdout, res.stdout[-3000:]
