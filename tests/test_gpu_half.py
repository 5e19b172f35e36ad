"""
The two forms of the one-kernel transform -- half-tile (csrc/ht_transform_half.cu, the default when the
planner's key quads are available) and tile (csrc/ht_transform.cu; HT_TRANSFORM_KERNEL=tile forces it)
-- against the oracle: ragged row counts and haplotype groups, haplotypes without alleles, a block
whose quads exceed the staging buffer (quads read from global memory), padded row pitch, fused counts.
Bit-exact: np.array_equal on the 0/1 bytes.
"""
import numpy as np
import pytest
import torch

from oracle import transform_oracle as orc

pytestmark = pytest.mark.gpu

from haptools_b200 import engine  # noqa: E402
from haptools_b200.plan import plan_from_refs  # noqa: E402


def _case(rng, n, p, H, max_k, long_hap=0):
    gt = rng.integers(0, 2, size=(n, p, 2), dtype=np.uint8)
    gt[rng.random((n, p, 2)) < 0.02] = 255
    k = rng.integers(0, min(max_k, p) + 1, size=H)
    if long_hap:
        k[H // 2] = min(long_hap, p)
    hap_off = np.concatenate([[0], np.cumsum(k)])
    ref_col = np.concatenate([np.sort(rng.choice(p, size=int(x), replace=False)) for x in k] + [np.zeros(0, dtype=np.int64)])
    ref_allele = rng.integers(0, 2, size=len(ref_col))
    # make some haplotypes present: copy their alleles into a few samples
    for h in rng.choice(H, size=min(H, 6), replace=False):
        cols = ref_col[hap_off[h] : hap_off[h + 1]].astype(np.int64)
        for s in rng.choice(n, size=min(n, 5), replace=False):
            gt[s, cols, :] = ref_allele[hap_off[h] : hap_off[h + 1]][:, None]
    plan = plan_from_refs(ref_col, ref_allele, hap_off, p, None)
    idxs = [np.arange(hap_off[h], hap_off[h + 1]) for h in range(H)]
    want = orc.transform_core(gt[:, ref_col.astype(np.int64)], ref_allele.astype(np.uint8), idxs, None, None)
    return gt, plan, want


SHAPES = [
    # n, p, H, max_k, long_hap
    (1, 8, 1, 3, 0), (31, 16, 7, 4, 0), (511, 24, 8, 5, 0), (512, 24, 9, 5, 0), (513, 24, 127, 5, 0),
    (1024, 40, 128, 6, 0), (1537, 64, 129, 20, 0), (2504, 200, 300, 20, 0), (3000, 300, 137, 60, 0),
    (2100, 4200, 40, 6, 4000), (700, 900, 260, 0, 0),
]


@pytest.mark.parametrize("kernel", ["half", "half2", "tile"])
@pytest.mark.parametrize("n,p,H,max_k,long_hap", SHAPES)
def test_transform_kernel_forms_vs_oracle(monkeypatch, kernel, n, p, H, max_k, long_hap):
    monkeypatch.setenv("HT_TRANSFORM_KERNEL", kernel[:4])
    monkeypatch.setenv("HT_HALF_PER_CTA", "2" if kernel == "half2" else "1")  # two half tiles per CTA
    rng = np.random.default_rng(1000 * n + H)
    gt, plan, want = _case(rng, n, p, H, max_k, long_hap)
    dplan = engine.DevicePlan(plan)
    gt_dev = torch.from_numpy(gt).cuda()
    planes = engine.alloc_planes(n, plan.n_keys, gt_dev.device)
    engine.pack(dplan, gt_dev.data_ptr(), gt_dev.stride(), n, planes)
    want_counts = want.reshape(n, H, 2).sum(axis=(0, 2)).astype(np.int64)
    for pitch in (engine.out_pitch(H), engine.out_pitch(H) + 48):
        out = torch.full((n, pitch), 9, dtype=torch.uint8, device="cuda")
        counts = torch.full((H,), -1, dtype=torch.int64, device="cuda")
        engine.transform_u8(dplan, planes, n, out.data_ptr(), pitch, method="direct", counts=counts)
        torch.cuda.synchronize()
        got = out[:, : 2 * H].cpu().numpy().reshape(n, H, 2).view(np.bool_)
        np.testing.assert_array_equal(got, want, err_msg=f"{kernel} pitch={pitch}")
        np.testing.assert_array_equal(counts.cpu().numpy(), want_counts, err_msg=f"{kernel} counts")
        # nothing is written beyond the 2H result bytes of a row
        assert bool((out[:, 2 * H :] == 9).all()), f"{kernel}: bytes beyond 2H were written"


def test_both_forms_agree_at_size():
    """131,072 samples x 4,096 haplotypes of 2-20 alleles over 20 k keys: the two kernels bit for bit."""
    import os

    from haptools_b200 import synth

    n, p, H = 131_072 + 777, 10_000, 4_096 + 8
    plan = synth.synth_plan(p, H, seed=5)
    dplan = engine.DevicePlan(plan)
    planes = engine.alloc_planes(n, plan.n_keys, "cuda")
    planes.random_(-(2**31), 2**31 - 1)
    outs = []
    old = os.environ.get("HT_TRANSFORM_KERNEL")
    try:
        for kernel in ("tile", "half"):
            os.environ["HT_TRANSFORM_KERNEL"] = kernel
            out = torch.empty((n, engine.out_pitch(H)), dtype=torch.uint8, device="cuda")
            counts = torch.zeros(H, dtype=torch.int64, device="cuda")
            engine.transform_u8(dplan, planes, n, out.data_ptr(), out.stride(0), method="direct", counts=counts)
            torch.cuda.synchronize()
            outs.append((out[:, : 2 * H], counts))
    finally:
        if old is None:
            os.environ.pop("HT_TRANSFORM_KERNEL", None)
        else:
            os.environ["HT_TRANSFORM_KERNEL"] = old
    assert torch.equal(outs[0][0], outs[1][0])
    assert torch.equal(outs[0][1], outs[1][1])
    assert int(outs[0][1].sum()) == int(outs[0][0].sum())


@pytest.mark.parametrize("form", ["tma", "line", "warp"])
@pytest.mark.parametrize("n,p,H", [(40, 8, 5), (1032, 40, 70), (2504, 64, 200), (3104, 200, 333), (1025, 1000, 400),
                                   (127, 72, 50), (129, 136, 90), (4000, 520, 700)])
def test_ancestry_pack_forms_vs_oracle(monkeypatch, form, n, p, H):
    """The three forms of the sample-major ancestry pack (csrc/ht_pack_anc_tma.cu: pack_sm2_anc_tma_kernel, the
    default; csrc/ht_pack.cu: pack_sm2_anc_line_kernel and pack_sm2_anc_kernel) against the oracle, with missing
    calls and labels >= 8 in the data; sample counts around the 128-row slabs and column counts around the
    64-column boxes of the TMA form."""
    monkeypatch.setenv("HT_PACK_ANC_KERNEL", form)
    from haptools_b200 import _cuda

    rng = np.random.default_rng(77 * n + H)
    gt = rng.integers(0, 2, size=(n, p, 2), dtype=np.uint8)
    gt[rng.random((n, p, 2)) < 0.01] = 255
    anc = np.ascontiguousarray(np.repeat(rng.integers(0, 5, size=(n, (p + 6) // 7, 2), dtype=np.uint8), 7, axis=1)[:, :p])
    anc[rng.random((n, p, 2)) < 0.002] = 9  # a label no haplotype asks for, outside the 3-bit planes
    k = rng.integers(0, min(6, p) + 1, size=H)
    hap_off = np.concatenate([[0], np.cumsum(k)])
    ref_col = np.concatenate([np.sort(rng.choice(p, size=int(x), replace=False)) for x in k] + [np.zeros(0, dtype=np.int64)])
    ref_allele = rng.integers(0, 2, size=len(ref_col))
    hap_label = rng.integers(0, 5, size=H)
    plan = plan_from_refs(ref_col, ref_allele, hap_off, p, np.repeat(hap_label, k))
    idxs = [np.arange(hap_off[h], hap_off[h + 1]) for h in range(H)]
    cols = ref_col.astype(np.int64)
    want = orc.transform_core(gt[:, cols], ref_allele.astype(np.uint8), idxs, anc[:, cols], hap_label.astype(np.uint8))
    dplan = engine.DevicePlan(plan)
    flags = torch.zeros(1, dtype=torch.int32, device="cuda")
    out = engine.transform_device(torch.from_numpy(gt).cuda(), dplan, ancestry=torch.from_numpy(anc).cuda(), flags=flags,
                                  pack_mode=_cuda.HT_PACK_SAMPLE_MAJOR_BIALLELIC if p % 8 == 0 else _cuda.HT_PACK_AUTO)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(out.cpu().numpy().view(np.bool_), want, err_msg=form)
    # the QC side channel: missing / multi-allelic calls in REFERENCED columns only
    sub = gt[:, plan.var_col.astype(np.int64)]
    want_flags = (1 if (sub >= 254).any() else 0) | (2 if ((sub >= 2) & (sub < 254)).any() else 0)
    assert int(flags.item()) == want_flags, form


@pytest.mark.parametrize("wide", ["0", "1"])
@pytest.mark.parametrize("group,chunk", [("0", "1"), ("1", "2"), ("3", "4"), ("64", "1")])
def test_ancestry_tma_pack_item_orders(monkeypatch, wide, group, chunk):
    """Item orders of the TMA-fed ancestry pack (HT_TMA_GROUP column blocks per group, HT_TMA_CHUNK consecutive
    items per CTA) and its 64-bit item arithmetic (HT_TMA_WIDE=1: the instance for launches of 2^32 items or
    more), multi-allelic calls in the data, several tiles and a ragged last column block: against the oracle."""
    from haptools_b200 import _cuda

    monkeypatch.setenv("HT_PACK_ANC_KERNEL", "tma")
    monkeypatch.setenv("HT_TMA_WIDE", wide)
    monkeypatch.setenv("HT_TMA_GROUP", group)
    monkeypatch.setenv("HT_TMA_CHUNK", chunk)
    rng = np.random.default_rng(5)
    n, p, H = 2300, 328, 500
    gt = rng.integers(0, 2, size=(n, p, 2), dtype=np.uint8)
    gt[rng.random((n, p, 2)) < 0.01] = 255
    gt[rng.random((n, p, 2)) < 0.002] = 3
    anc = np.ascontiguousarray(np.repeat(rng.integers(0, 5, size=(n, (p + 6) // 7, 2), dtype=np.uint8), 7, axis=1)[:, :p])
    k = rng.integers(0, 7, size=H)
    hap_off = np.concatenate([[0], np.cumsum(k)])
    ref_col = np.concatenate([np.sort(rng.choice(p, size=int(x), replace=False)) for x in k] + [np.zeros(0, dtype=np.int64)])
    ref_allele = rng.integers(0, 2, size=len(ref_col))
    hap_label = rng.integers(0, 5, size=H)
    plan = plan_from_refs(ref_col, ref_allele, hap_off, p, np.repeat(hap_label, k))
    idxs = [np.arange(hap_off[h], hap_off[h + 1]) for h in range(H)]
    cols = ref_col.astype(np.int64)
    want = orc.transform_core(gt[:, cols], ref_allele.astype(np.uint8), idxs, anc[:, cols], hap_label.astype(np.uint8))
    flags = torch.zeros(1, dtype=torch.int32, device="cuda")
    out = engine.transform_device(torch.from_numpy(gt).cuda(), engine.DevicePlan(plan), ancestry=torch.from_numpy(anc).cuda(),
                                  flags=flags, pack_mode=_cuda.HT_PACK_SAMPLE_MAJOR_BIALLELIC)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(out.cpu().numpy().view(np.bool_), want)
    assert int(flags.item()) == 3


@pytest.mark.parametrize("layout", ["C", "C3", "F"])
def test_pageable_upload_through_bounce_buffers(layout):
    """Host arrays in PAGEABLE memory of more than 4 MB take ht_copy2d's own staged upload (host threads
    + page-locked bounce buffers, csrc/ht_capi.cu): dense rows, rows with a phase column in between (pitch >
    width), and the variant-major view of the VCF reader; several chunks, several blocks per chunk."""
    n, p, H = 20_000, 600, 64
    rng = np.random.default_rng(3)
    gt, plan, want = _case(rng, n, p, H, 5)
    if layout == "C":
        arr = gt  # 24 MB, dense
    elif layout == "C3":
        c3 = np.ones((n, p, 3), dtype=np.uint8)
        c3[:, :, :2] = gt
        arr = c3[:, :, :2]
    else:
        arr = np.ascontiguousarray(gt.transpose(1, 0, 2)).transpose(1, 0, 2)
    got = engine.transform_host(arr, plan, chunk_samples=8192)
    np.testing.assert_array_equal(got, want)
    got = engine.transform_host(arr, plan)
    np.testing.assert_array_equal(got, want)
    # chunks dealt round-robin to every GPU of the box: the bounce buffers serve several devices in turn
    got = engine.transform_host(arr, plan, chunk_samples=4096, devices="all")
    np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("seed", range(12))
def test_random_shapes_both_forms(monkeypatch, seed):
    """Random shapes around the kernels' boundaries (512-sample half tiles, groups of 8 / 16 / 128
    haplotypes, staging capacity) through both transform kernels, against the oracle."""
    rng = np.random.default_rng(9000 + seed)
    n = int(rng.choice([1, 7, 511, 512, 513, 1023, 1024, 1025, 1535, 2047, 2049, 2560, 3071])) + int(rng.integers(0, 3))
    H = int(rng.choice([1, 7, 8, 9, 15, 16, 17, 127, 128, 129, 255, 257, 300]))
    p = int(rng.integers(4, 400))
    max_k = int(rng.choice([0, 1, 3, 4, 5, 8, 21, 60]))
    gt, plan, want = _case(rng, n, p, H, max_k)
    dplan = engine.DevicePlan(plan)
    gt_dev = torch.from_numpy(gt).cuda()
    planes = engine.alloc_planes(n, plan.n_keys, gt_dev.device)
    engine.pack(dplan, gt_dev.data_ptr(), gt_dev.stride(), n, planes)
    pitch = engine.out_pitch(H) + 16 * int(rng.integers(0, 3))
    for kernel in ("half", "tile"):
        monkeypatch.setenv("HT_TRANSFORM_KERNEL", kernel)
        out = torch.full((n, pitch), 5, dtype=torch.uint8, device="cuda")
        engine.transform_u8(dplan, planes, n, out.data_ptr(), pitch, method="direct")
        torch.cuda.synchronize()
        got = out[:, : 2 * H].cpu().numpy().reshape(n, H, 2).view(np.bool_)
        np.testing.assert_array_equal(got, want, err_msg=f"{kernel} n={n} H={H} p={p} max_k={max_k}")
        assert bool((out[:, 2 * H :] == 5).all())


@pytest.mark.parametrize("mapping", ["slab", "tile"])
@pytest.mark.parametrize("kind", ["plain", "ancestry"])
@pytest.mark.parametrize("n,p,H", [(40, 8, 5), (129, 520, 40), (1032, 40, 70), (2504, 600, 200), (3104, 1032, 333)])
def test_sample_major_pack_mappings_vs_oracle(monkeypatch, mapping, kind, n, p, H):
    """The two CTA mappings of the sample-major biallelic pack kernels (csrc/ht_pack.cu: slab = 128 rows x
    512 columns, the default of the plain kernel; tile = 1,024 rows x 64 columns, the default with ancestry)
    against the oracle: ragged rows and columns, missing calls, column counts around 512."""
    monkeypatch.setenv("HT_PACK_SM2_KERNEL", mapping)
    from haptools_b200 import _cuda

    rng = np.random.default_rng(31 * n + p + H)
    gt = rng.integers(0, 2, size=(n, p, 2), dtype=np.uint8)
    gt[rng.random((n, p, 2)) < 0.01] = 255
    anc = None
    k = rng.integers(0, min(6, p) + 1, size=H)
    hap_off = np.concatenate([[0], np.cumsum(k)])
    ref_col = np.concatenate([np.sort(rng.choice(p, size=int(x), replace=False)) for x in k] + [np.zeros(0, dtype=np.int64)])
    ref_allele = rng.integers(0, 2, size=len(ref_col))
    hap_label = ref_label = None
    if kind == "ancestry":
        anc = np.ascontiguousarray(np.repeat(rng.integers(0, 5, size=(n, (p + 6) // 7, 2), dtype=np.uint8), 7, axis=1)[:, :p])
        hap_label = rng.integers(0, 5, size=H)
        ref_label = np.repeat(hap_label, k)
    plan = plan_from_refs(ref_col, ref_allele, hap_off, p, ref_label)
    idxs = [np.arange(hap_off[h], hap_off[h + 1]) for h in range(H)]
    cols = ref_col.astype(np.int64)
    want = orc.transform_core(gt[:, cols], ref_allele.astype(np.uint8), idxs,
                              anc[:, cols] if anc is not None else None,
                              hap_label.astype(np.uint8) if hap_label is not None else None)
    dplan = engine.DevicePlan(plan)
    out = engine.transform_device(torch.from_numpy(gt).cuda(), dplan,
                                  ancestry=torch.from_numpy(anc).cuda() if anc is not None else None,
                                  pack_mode=_cuda.HT_PACK_SAMPLE_MAJOR_BIALLELIC if p % 8 == 0 else _cuda.HT_PACK_AUTO)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(out.cpu().numpy().view(np.bool_), want, err_msg=f"{mapping}/{kind}")


@pytest.mark.parametrize("order,waves", [("v", ""), ("t", ""), ("v", "1"), ("t", "1"), ("v", "32")])
@pytest.mark.parametrize("kind", ["plain", "ancestry"])
@pytest.mark.parametrize("n,p,H", [(40, 8, 5), (1024, 130, 40), (1032, 40, 70), (2504, 600, 200), (5000, 1032, 333)])
def test_variant_major_pack_item_orders_vs_oracle(monkeypatch, order, waves, kind, n, p, H):
    """pack_vm_kernel (csrc/ht_pack.cu) walks its (tile, variant) items variant-first (the default) or
    tile-first, over a grid of 1..32 waves of CTAs (one wave: many items per warp, so the two-items-ahead
    prefetch runs through every wrap of the walk): both orders against the oracle on ragged shapes with
    missing and multi-allelic calls."""
    monkeypatch.setenv("HT_VM_ORDER", order)
    if waves:
        monkeypatch.setenv("HT_VM_WAVES", waves)
    from haptools_b200 import _cuda

    rng = np.random.default_rng(17 * n + p + H)
    gt = rng.integers(0, 2, size=(n, p, 2), dtype=np.uint8)
    gt[rng.random((n, p, 2)) < 0.01] = 255
    gt[rng.random((n, p, 2)) < 0.01] = 2
    k = rng.integers(0, min(6, p) + 1, size=H)
    hap_off = np.concatenate([[0], np.cumsum(k)])
    ref_col = np.concatenate([np.sort(rng.choice(p, size=int(x), replace=False)) for x in k] + [np.zeros(0, dtype=np.int64)])
    ref_allele = rng.integers(0, 2, size=len(ref_col))
    anc = hap_label = ref_label = None
    if kind == "ancestry":
        anc = np.ascontiguousarray(np.repeat(rng.integers(0, 5, size=(n, (p + 6) // 7, 2), dtype=np.uint8), 7, axis=1)[:, :p])
        hap_label = rng.integers(0, 5, size=H)
        ref_label = np.repeat(hap_label, k)
    plan = plan_from_refs(ref_col, ref_allele, hap_off, p, ref_label)
    idxs = [np.arange(hap_off[h], hap_off[h + 1]) for h in range(H)]
    cols = ref_col.astype(np.int64)
    want = orc.transform_core(gt[:, cols], ref_allele.astype(np.uint8), idxs,
                              anc[:, cols] if anc is not None else None,
                              hap_label.astype(np.uint8) if hap_label is not None else None)

    def vm(a):  # variant-major on the device, rows padded to 16 bytes
        pitch = (2 * n + 15) // 16 * 16
        buf = torch.zeros((p, pitch), dtype=torch.uint8, device="cuda")
        buf[:, : 2 * n] = torch.from_numpy(np.ascontiguousarray(a.transpose(1, 0, 2)).reshape(p, 2 * n)).cuda()
        return buf[:, : 2 * n].unflatten(1, (n, 2)).permute(1, 0, 2)

    dplan = engine.DevicePlan(plan)
    out = engine.transform_device(vm(gt), dplan, ancestry=vm(anc) if anc is not None else None,
                                  pack_mode=_cuda.HT_PACK_VARIANT_MAJOR)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(out.cpu().numpy().view(np.bool_), want, err_msg=f"{order}/{waves}/{kind}")


@pytest.mark.parametrize("layout", ["F", "F3"])
@pytest.mark.parametrize("kind", ["plain", "ancestry"])
def test_variant_major_host_array_uploads_only_referenced_variants(layout, kind):
    """transform_host on a variant-major host array (VCF reader layout) whose plan refers to few of the
    variants: only those rows are uploaded (ht_copy_rows_h2d + DevicePlan.compact_tables), with and without
    the phase column between the samples' calls, with ancestry, pageable and page-locked sources."""
    n, p, H = 3000, 4000, 120
    rng = np.random.default_rng(5 + (kind == "ancestry"))
    gt = rng.integers(0, 2, size=(n, p, 2), dtype=np.uint8)
    gt[rng.random((n, p, 2)) < 0.01] = 255
    used = np.sort(rng.choice(p, size=500, replace=False))  # the haplotypes draw from 500 of the 4,000 variants
    k = rng.integers(1, 7, size=H)
    hap_off = np.concatenate([[0], np.cumsum(k)])
    ref_col = np.concatenate([np.sort(rng.choice(used, size=int(x), replace=False)) for x in k])
    ref_allele = rng.integers(0, 2, size=len(ref_col))
    anc = hap_label = ref_label = None
    if kind == "ancestry":
        anc = np.ascontiguousarray(np.repeat(rng.integers(0, 4, size=(n, (p + 6) // 7, 2), dtype=np.uint8), 7, axis=1)[:, :p])
        hap_label = rng.integers(0, 4, size=H)
        ref_label = np.repeat(hap_label, k)
    plan = plan_from_refs(ref_col, ref_allele, hap_off, p, ref_label)
    assert 2 * plan.n_vars <= p
    idxs = [np.arange(hap_off[h], hap_off[h + 1]) for h in range(H)]
    cols = ref_col.astype(np.int64)
    want = orc.transform_core(gt[:, cols], ref_allele.astype(np.uint8), idxs,
                              anc[:, cols] if anc is not None else None,
                              hap_label.astype(np.uint8) if hap_label is not None else None)

    def vm(a):  # the VCF reader's layout: (p, n, 2|3) in memory, seen as (n, p, 2)
        if layout == "F":
            return np.ascontiguousarray(a.transpose(1, 0, 2)).transpose(1, 0, 2)
        f3 = np.ones((p, n, 3), dtype=np.uint8)
        f3[:, :, :2] = a.transpose(1, 0, 2)
        return f3.transpose(1, 0, 2)[:, :, :2]

    g, an = vm(gt), (vm(anc) if anc is not None else None)
    for chunk in (None, 1024):
        got = engine.transform_host(g, plan, ancestry=an, chunk_samples=chunk)
        np.testing.assert_array_equal(got, want, err_msg=f"{layout}/{kind} chunk={chunk}")
    if layout == "F":  # page-locked source: same path (the gather always goes through the bounce buffers)
        pinned = engine.pinned_empty((p, n, 2))
        pinned[:] = gt.transpose(1, 0, 2)
        got = engine.transform_host(pinned.transpose(1, 0, 2), plan, ancestry=an)
        np.testing.assert_array_equal(got, want)
