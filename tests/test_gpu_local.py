"""
Parity of the two-phase ("local") transform -- ht_transform_bits_local, ht_unpack_bits_u8,
ht_transform_u8_local -- against the oracle and against the one-kernel form, through the C-ABI.
Bit-exact.
"""
import ctypes

import numpy as np
import pytest
import torch

from oracle import transform_oracle as orc

pytestmark = pytest.mark.gpu

from haptools_b200 import _cuda, engine, synth  # noqa: E402
from haptools_b200 import dist as hdist  # noqa: E402
from haptools_b200.plan import plan_from_refs  # noqa: E402
from test_gpu_parity import SHAPES, _random_arrays  # noqa: E402


def _planes(gt, plan, dplan):
    n = gt.shape[0]
    g = torch.from_numpy(np.ascontiguousarray(gt[:, :, :2])).cuda()
    planes = engine.alloc_planes(n, plan.n_keys, g.device)
    engine.pack(dplan, g.data_ptr(), g.stride(), n, planes)
    return planes


def _u8(dplan, planes, n, method, pitch=None, counts=None, workspace=None, offset=0, use_quads=True):
    H = dplan.n_haps
    pitch = pitch or engine.out_pitch(H)
    buf = torch.full((n * pitch + offset + 64,), 7, dtype=torch.uint8, device="cuda")
    engine.transform_u8(dplan, planes, n, buf.data_ptr() + offset, pitch, counts=counts, method=method, workspace=workspace, use_quads=use_quads)
    torch.cuda.synchronize()
    body = buf[offset : offset + n * pitch].view(n, pitch)
    assert bool((body[:, 2 * H :] == 7).all()) and bool((buf[offset + n * pitch :] == 7).all())  # nothing outside the rows
    return body[:, : 2 * H].cpu().numpy().reshape(n, H, 2).view(np.bool_)


@pytest.mark.parametrize("n,p,H", SHAPES)
def test_local_random_vs_oracle(n, p, H):
    rng = np.random.default_rng(n * 31 + p * 7 + H)
    gt, _, plan, want = _random_arrays(rng, n, p, H, multi=True, missing=0.03)
    for kw in ({}, {"local_dmax": 8, "local_max_block": 5}):
        dplan = engine.DevicePlan(plan, **kw)
        planes = _planes(gt, plan, dplan)
        counts = torch.empty(H, dtype=torch.int64, device="cuda")
        got = _u8(dplan, planes, n, "local", counts=counts)
        np.testing.assert_array_equal(got, want)
        np.testing.assert_array_equal(counts.cpu().numpy(), want.sum(axis=(0, 2)))
        np.testing.assert_array_equal(_u8(dplan, planes, n, "direct"), want)
        # the one-kernel form with and without the planner's quads (the kernel then pads the CSR itself)
        np.testing.assert_array_equal(_u8(dplan, planes, n, "direct", use_quads=False, counts=counts), want)
        np.testing.assert_array_equal(counts.cpu().numpy(), want.sum(axis=(0, 2)))


def test_local_structured_plan_blocks_groups_and_pitches(monkeypatch):
    """A plan with real locality (many haplotypes per variant), several dmax / block sizes,
    workspaces of 1 and 2 tiles (group loop), odd pitches (2-byte and 1-byte store paths)."""
    n, p, H = 3300, 700, 1500
    plan = synth.synth_plan(p, H, seed=11, max_alleles=12)
    gt = synth.synth_gt(n, p, seed=4, mode=1, missing_per_64k=500)
    want = orc.transform_from_csr(gt, plan.key_col().astype(np.int64), plan.key_allele,
                                  plan.hap_off.astype(np.int64), plan.hap_key.astype(np.int64))
    assert want.any()
    for kw in ({}, {"local_dmax": 64}, {"local_dmax": 40, "local_max_block": 33}, {"local_dmax": 700}):
        dplan = engine.DevicePlan(plan, **kw)
        lt, _ = dplan.local()
        planes = _planes(gt, plan, dplan)
        np.testing.assert_array_equal(_u8(dplan, planes, n, "local"), want, err_msg=str(kw))
        for tiles in (1, 2):
            ws = torch.empty(tiles * H * 64, dtype=torch.int32, device="cuda")
            np.testing.assert_array_equal(_u8(dplan, planes, n, "local", workspace=ws), want)
        np.testing.assert_array_equal(_u8(dplan, planes, n, "local", pitch=2 * H + 2), want)
        np.testing.assert_array_equal(_u8(dplan, planes, n, "local", pitch=2 * H + 3, offset=1), want)
    assert not engine.DevicePlan(plan).use_local()  # switched off by default (engine.LOCAL_AUTO)
    monkeypatch.setattr(engine, "LOCAL_AUTO", True)
    assert engine.DevicePlan(plan).use_local()
    np.testing.assert_array_equal(_u8(engine.DevicePlan(plan), planes, n, "auto"), want)
    with pytest.raises(ValueError):
        ws = torch.empty(H * 64 - 4, dtype=torch.int32, device="cuda")  # less than one tile
        _u8(dplan, planes, n, "local", workspace=ws)


def test_local_long_lists_wide_and_empty_haplotypes():
    """Blocks whose key lists exceed the staged 4096 keys (keys read from global memory),
    a haplotype wider than dmax (block not staged: records straight from L2), haplotypes
    without alleles (all True)."""
    rng = np.random.default_rng(8)
    n, p, H = 2100, 400, 300
    k = np.full(H, 40)
    k[::9] = 0
    k[7] = 330
    off = np.concatenate([[0], np.cumsum(k)])
    cols = []
    for h, x in enumerate(k):
        if x == 330:
            cols.append(np.sort(rng.choice(p, size=x, replace=False)))
        else:
            cols.append(np.sort(rng.choice(130, size=int(x), replace=False)) + 100)  # 130 columns -> <= 260 keys
    cols = np.concatenate(cols)
    alle = rng.integers(0, 2, size=len(cols))
    plan = plan_from_refs(cols, alle, off, p)
    # founders: make a good share of the long haplotypes present
    base = rng.integers(0, 2, size=(1, p, 2), dtype=np.uint8)
    gt = np.where(rng.random((n, p, 2)) < 0.01, 1 - base, base).astype(np.uint8)
    for h in (1, 2, 3, 7):
        gt[h * 50 : h * 50 + 40, cols[off[h] : off[h + 1]], :] = alle[off[h] : off[h + 1], None]
    want = orc.transform_from_csr(gt, plan.key_col().astype(np.int64), plan.key_allele,
                                  plan.hap_off.astype(np.int64), plan.hap_key.astype(np.int64))
    assert want[:, 1].any() and want[:, 7].any() and want[:, 0].all()
    dplan = engine.DevicePlan(plan, local_dmax=288)
    lt, _ = dplan.local()
    assert (lt.blk_key_cnt == 0).any()  # the wide haplotype
    nk = np.diff(lt.s_hap_off.astype(np.int64)[lt.blk_hap_off.astype(np.int64)])
    assert (nk[lt.blk_key_cnt > 0] > 4096).any()  # some staged block with a long list
    planes = _planes(gt, plan, dplan)
    counts = torch.empty(H, dtype=torch.int64, device="cuda")
    np.testing.assert_array_equal(_u8(dplan, planes, n, "local", counts=counts), want)
    np.testing.assert_array_equal(counts.cpu().numpy(), want.sum(axis=(0, 2)))


def test_bits_local_equals_bits_direct_and_unpack_roundtrip():
    n, p, H = 2600, 300, 700
    plan = synth.synth_plan(p, H, seed=2, max_alleles=8)
    gt = synth.synth_gt(n, p, seed=9, mode=1)
    dplan = engine.DevicePlan(plan)
    planes = _planes(gt, plan, dplan)
    tiles = (n + 1023) // 1024
    a = torch.empty((tiles, H, 32, 2), dtype=torch.int32, device="cuda")
    b = torch.empty_like(a)
    counts = torch.empty(H, dtype=torch.int64, device="cuda")
    engine.transform_bits(dplan, planes, n, a, method="direct")
    engine.transform_bits(dplan, planes, n, b, method="local", counts=counts)
    torch.cuda.synchronize()
    assert torch.equal(a, b)
    want = hdist.unpack_result_bits(a.cpu().numpy(), n)
    np.testing.assert_array_equal(counts.cpu().numpy(), want.sum(axis=(0, 2)))
    np.testing.assert_array_equal(engine.count_bits(b, n, H).cpu().numpy(), want.sum(axis=(0, 2)))
    for pitch in (engine.out_pitch(H), 2 * H + 6, 2 * H + 1):
        out = torch.zeros((n, pitch), dtype=torch.uint8, device="cuda")
        engine.unpack_bits_u8(b, n, H, out.data_ptr(), pitch)
        torch.cuda.synchronize()
        np.testing.assert_array_equal(out[:, : 2 * H].cpu().numpy().reshape(n, H, 2).view(np.bool_), want)
        assert not bool(out[:, 2 * H :].any())


def test_local_degenerate_sizes():
    lib = _cuda.lib()
    plan = plan_from_refs(np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(4, np.int64), 5)  # 3 empty haplotypes
    dplan = engine.DevicePlan(plan)
    got = _u8(dplan, None, 70, "local")
    assert got.shape == (70, 3, 2) and got.all()
    _, tables = dplan.local()
    assert lib.ht_transform_u8_local(None, 0, 0, ctypes.byref(tables), 3, None, 6, None, None, 0, None) == 0  # no samples
    assert lib.ht_transform_u8_local(None, -1, 0, ctypes.byref(tables), 3, None, 6, None, None, 0, None) == _cuda.HT_ERR_BAD_ARG
    assert lib.ht_transform_u8_local(None, 5, 0, None, 3, None, 6, None, None, 0, None) == _cuda.HT_ERR_BAD_ARG
    assert lib.ht_unpack_bits_u8(None, 5, 3, None, 6, None) == _cuda.HT_ERR_BAD_ARG
    assert lib.ht_local_workspace_bytes(5000, 10) == 5 * 10 * 256 and lib.ht_local_workspace_bytes(10**6, 10) == 64 * 10 * 256


def test_local_equals_direct_at_scale():
    """UKB-like density at a size the oracle cannot reach: the two forms must agree byte for
    byte, counts included (size-independent property)."""
    n, p, H = 40_000, 5_000, 10_000
    plan = synth.synth_plan(p, H, seed=2)
    dplan = engine.DevicePlan(plan)
    assert dplan.local()[0].worthwhile(plan.n_refs)
    gt = torch.empty((n, p, 2), dtype=torch.uint8, device="cuda")
    _cuda.check(_cuda.lib().ht_synth_gt(gt.data_ptr(), *gt.stride(), n, p, 0, 3, 1, 20, 0), "synth")
    planes = engine.alloc_planes(n, plan.n_keys, gt.device)
    engine.pack(dplan, gt.data_ptr(), gt.stride(), n, planes)
    pitch = engine.out_pitch(H)
    a = torch.empty((n, pitch), dtype=torch.uint8, device="cuda")
    b = torch.empty((n, pitch), dtype=torch.uint8, device="cuda")
    ca = torch.empty(H, dtype=torch.int64, device="cuda")
    cb = torch.empty(H, dtype=torch.int64, device="cuda")
    engine.transform_u8(dplan, planes, n, a.data_ptr(), pitch, counts=ca, method="direct")
    engine.transform_u8(dplan, planes, n, b.data_ptr(), pitch, counts=cb, method="local")
    torch.cuda.synchronize()
    assert torch.equal(a[:, : 2 * H], b[:, : 2 * H]) and torch.equal(ca, cb)
    assert int(ca.sum()) == int(a[:, : 2 * H].sum(dtype=torch.int64)) > 0
    # a slice against the oracle
    ns = 48
    sub = synth.synth_gt(ns, p, 3, mode=1, missing_per_64k=20, sample_offset=n - ns)
    want = orc.transform_from_csr(sub, plan.key_col().astype(np.int64), plan.key_allele,
                                  plan.hap_off.astype(np.int64), plan.hap_key.astype(np.int64))
    np.testing.assert_array_equal(b[n - ns :, : 2 * H].cpu().numpy().reshape(ns, H, 2).view(np.bool_), want)
