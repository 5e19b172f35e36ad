"""
ht_transform_u8_staged (csrc/ht_transform_half.cu, kStage): blocks of 128 haplotypes whose keys fall into a
short run stage that run in shared memory; other blocks of the same launch read from L2.  Bit-exact against
the oracle for position-sorted sets, mixed staged / unstaged blocks, ragged edges, haplotypes without alleles,
and against the unstaged kernel at size.
"""
import numpy as np
import pytest
import torch

from oracle import transform_oracle as orc

pytestmark = pytest.mark.gpu

from haptools_b200 import engine, synth  # noqa: E402
from haptools_b200.plan import plan_from_refs  # noqa: E402


def _local_case(rng, n, p, H, max_k, window, empty_every=0, wild_every=0):
    """Haplotypes in position order: haplotype h draws its variants from a window that moves with h; every
    `wild_every`-th one instead draws from everywhere (its block then spans too many keys to be staged)."""
    gt = rng.integers(0, 2, size=(n, p, 2), dtype=np.uint8)
    gt[rng.random((n, p, 2)) < 0.02] = 255
    cols, lens = [], []
    for h in range(H):
        k = int(rng.integers(1, max_k + 1))
        if empty_every and h % empty_every == empty_every - 1:
            k = 0
        if wild_every and h % wild_every == 0:
            c = np.sort(rng.choice(p, size=min(k, p), replace=False))
        else:
            start = int((p - window) * h / max(H - 1, 1)) if p > window else 0
            c = start + np.sort(rng.choice(min(window, p), size=min(k, window, p), replace=False))
        cols.append(c)
        lens.append(len(c))
    hap_off = np.concatenate([[0], np.cumsum(lens)])
    ref_col = np.concatenate(cols + [np.zeros(0, dtype=np.int64)]).astype(np.int64)
    ref_allele = rng.integers(0, 2, size=len(ref_col))
    for h in rng.choice(H, size=min(H, 8), replace=False):  # make some haplotypes present
        c = ref_col[hap_off[h]:hap_off[h + 1]]
        for s in rng.choice(n, size=min(n, 6), replace=False):
            gt[s, c, :] = ref_allele[hap_off[h]:hap_off[h + 1]][:, None]
    plan = plan_from_refs(ref_col, ref_allele, hap_off, p, None)
    idxs = [np.arange(hap_off[h], hap_off[h + 1]) for h in range(H)]
    want = orc.transform_core(gt[:, ref_col], ref_allele.astype(np.uint8), idxs, None, None)
    return gt, plan, want


def _run(gt, plan, stage_lines, method="staged", counts=False, extra_pitch=0):
    n, H = gt.shape[0], plan.n_haps
    dplan = engine.DevicePlan(plan)
    if stage_lines is not None:  # force a budget (the planner's choice otherwise)
        lo, cnt = plan.block_spans()
        dplan.stage_lines = stage_lines
        dplan.blk_key_lo = torch.from_numpy(lo.view(np.int32)).cuda()
        dplan.blk_key_cnt = torch.from_numpy(cnt.view(np.int32)).cuda()
    g = torch.from_numpy(gt).cuda()
    planes = engine.alloc_planes(n, plan.n_keys, g.device)
    engine.pack(dplan, g.data_ptr(), g.stride(), n, planes)
    pitch = engine.out_pitch(H) + extra_pitch
    out = torch.full((n, pitch), 9, dtype=torch.uint8, device="cuda")
    cts = torch.full((H,), -1, dtype=torch.int64, device="cuda") if counts else None
    engine.transform_u8(dplan, planes, n, out.data_ptr(), pitch, method=method, counts=cts)
    torch.cuda.synchronize()
    assert bool((out[:, 2 * H:] == 9).all())
    return out[:, : 2 * H].cpu().numpy().reshape(n, H, 2).view(np.bool_), (cts.cpu().numpy() if counts else None), dplan


SHAPES = [
    # n, p, H, max_k, window, empty_every, wild_every
    (1, 40, 1, 3, 16, 0, 0), (511, 300, 127, 6, 40, 0, 0), (512, 300, 128, 6, 40, 5, 0), (513, 600, 129, 8, 60, 0, 0),
    (1537, 2000, 700, 20, 80, 7, 0), (2504, 400, 1000, 20, 80, 0, 300), (3000, 900, 260, 12, 900, 0, 0),
    (700, 4000, 515, 20, 100, 3, 130), (1100, 300, 640, 9, 30, 16, 0), (600, 200, 384, 3, 20, 2, 0),
]


@pytest.mark.parametrize("hpc", [1, 2])
@pytest.mark.parametrize("stage_lines", [None, 32, 256, 1536])
@pytest.mark.parametrize("n,p,H,max_k,window,empty_every,wild_every", SHAPES)
def test_staged_vs_oracle(monkeypatch, n, p, H, max_k, window, empty_every, wild_every, stage_lines, hpc):
    monkeypatch.setenv("HT_HALF_PER_CTA", str(hpc))  # half tiles per CTA (csrc/ht_transform_half.cu)
    rng = np.random.default_rng(31 * n + H)
    gt, plan, want = _local_case(rng, n, p, H, max_k, window, empty_every, wild_every)
    got, counts, dplan = _run(gt, plan, stage_lines, counts=True, extra_pitch=32 if n % 2 else 0)
    np.testing.assert_array_equal(got, want)
    np.testing.assert_array_equal(counts, want.sum(axis=(0, 2)))
    if stage_lines is not None:
        _, cnt = plan.block_spans()
        staged = (cnt > 0) & (cnt <= stage_lines)
        if wild_every and stage_lines == 256:
            assert staged.any() and (~staged).any(), "this case is meant to mix staged and unstaged blocks"


def test_planner_picks_staging_for_sorted_sets_only():
    sorted_plan = synth.synth_plan(5000, 20_000, seed=3, sort=True)
    unsorted_plan = synth.synth_plan(5000, 20_000, seed=3, sort=False)
    assert 0 < sorted_plan.stage_lines() <= 608 and sorted_plan.stage_lines() % 32 == 0
    assert unsorted_plan.stage_lines() == 0
    lo, cnt = sorted_plan.block_spans()
    off, key = sorted_plan.hap_off.astype(np.int64), sorted_plan.hap_key.astype(np.int64)
    for b in (0, 7, len(cnt) - 1):
        ks = key[off[b * 128]:off[min((b + 1) * 128, sorted_plan.n_haps)]]
        assert lo[b] == ks.min() and cnt[b] == ks.max() - ks.min() + 1


def test_staged_agrees_with_direct_at_size():
    """65,536 + 300 samples x 20,008 position-sorted haplotypes of 2-20 alleles: staged == direct, bit for bit,
    on random plane contents (every bit pattern, not just what a pack produces)."""
    n, p, H = 65_536 + 300, 10_000, 20_008
    plan = synth.synth_plan(p, H, seed=6, sort=True)
    dplan = engine.DevicePlan(plan)
    assert dplan.stage_lines > 0
    planes = engine.alloc_planes(n, plan.n_keys, "cuda")
    planes.random_(-(2**31), 2**31 - 1)
    outs = []
    for method in ("direct", "staged"):
        out = torch.empty((n, engine.out_pitch(H)), dtype=torch.uint8, device="cuda")
        counts = torch.zeros(H, dtype=torch.int64, device="cuda")
        engine.transform_u8(dplan, planes, n, out.data_ptr(), out.stride(0), method=method, counts=counts)
        torch.cuda.synchronize()
        outs.append((out[:, : 2 * H], counts))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


def test_staged_rejects_bad_arguments():
    from haptools_b200 import _cuda

    lib = _cuda.lib()
    assert lib.ht_transform_u8_staged(None, 10, 0, None, None, None, None, None, None, 99999, 1, None, 16, None, None) == _cuda.HT_ERR_BAD_ARG
    assert lib.ht_transform_u8_staged(None, 10, 0, None, None, None, None, None, None, 64, 1, None, 16, None, None) == _cuda.HT_ERR_BAD_ARG
