"""
Parity of the CUDA path (through the C-ABI) against the oracle and the committed golden
vectors.  Bit-exact: every comparison is np.array_equal on the 0/1 bytes.
"""
import types

import numpy as np
import pytest
import torch

import _cases
from oracle import transform_oracle as orc

pytestmark = pytest.mark.gpu

from haptools_b200 import _cuda, data, engine, stream, synth  # noqa: E402
from haptools_b200 import transform as htransform  # noqa: E402
from haptools_b200.plan import plan_from_haplotypes, plan_from_refs  # noqa: E402

NS = types.SimpleNamespace(
    Variant=data.Variant, Haplotype=data.Haplotype, Haplotypes=data.Haplotypes,
    GenotypesVCF=data.GenotypesVCF, HaplotypeAncestry=htransform.HaplotypeAncestry,
    HaplotypesAncestry=htransform.HaplotypesAncestry, GenotypesAncestry=htransform.GenotypesAncestry,
)
GOLDEN = _cases.list_golden()
MODES = {"generic": _cuda.HT_PACK_GENERIC, "vm": _cuda.HT_PACK_VARIANT_MAJOR, "sm": _cuda.HT_PACK_SAMPLE_MAJOR,
         "sm2": _cuda.HT_PACK_SAMPLE_MAJOR_BIALLELIC, "auto": _cuda.HT_PACK_AUTO}


# ------------------------------------------------------------------ class-level API
@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: p.stem)
def test_golden_class_api(path):
    case = _cases.load_case(path)
    haps, gts = _cases.build_objects(case, NS)
    anc = case["kind"] == "ancestry"
    if anc and case["gt"].shape[2] != 2:
        with pytest.raises(ValueError):
            haps.transform(gts)
        return
    out = haps.transform(gts)
    assert isinstance(out, data.GenotypesVCF)
    assert out.data.dtype == np.bool_ and out.data.flags.c_contiguous
    assert out.samples == gts.samples
    assert list(out.variants["id"]) == [h["id"] for h in case["haps"]]
    assert all(a == ("A", "T") for a in out.variants["alleles"])
    np.testing.assert_array_equal(out.data, case["expected_plural"])
    for i, h in enumerate(haps.data.values()):
        if len(h.variants) == 0:
            with pytest.raises(ValueError):
                h.transform(gts)
            continue
        one = h.transform(gts)
        assert one.dtype == np.bool_ and one.shape == (len(gts.samples), 2)
        np.testing.assert_array_equal(one, case["expected_single"][i])


def test_hap_gts_is_filled_and_returned():
    case = _cases.load_case(_cases.GOLDEN_DIR / "g3_plain_refalt_tweaked.npz")
    haps, gts = _cases.build_objects(case, NS)
    before = gts.data.copy()
    target = data.GenotypesVCF(fname=None)
    assert haps.transform(gts, target) is target
    np.testing.assert_array_equal(target.data, case["expected_plural"])
    np.testing.assert_array_equal(gts.data, before)  # gts is not mutated


# ------------------------------------------------------------------ array-level, every kernel
def _layouts(gt):
    """The physical layouts of SURVEY.md 4c for the same logical (n, p, 2) array."""
    n, p = gt.shape[:2]
    g2 = np.ascontiguousarray(gt[:, :, :2])
    out = {"C": g2}
    out["F"] = np.ascontiguousarray(g2.transpose(1, 0, 2)).transpose(1, 0, 2)  # VCF reader
    c3 = np.zeros((n, p, 3), dtype=gt.dtype)
    c3[:, :, :2] = g2
    c3[:, :, 2] = 1
    out["C3"] = c3[:, :, :2]  # PGEN reader after check_phase
    f3 = np.ascontiguousarray(c3.transpose(1, 0, 2)).transpose(1, 0, 2)
    out["F3"] = f3[:, :, :2]  # VCF reader after check_phase
    return out


def _to_device_keep_strides(a):
    """Upload the memory span of a (possibly strided, positive strides) 1-byte numpy view and
    rebuild the same view on the device: the kernels see the reference's physical layout."""
    a = a.view(np.uint8)
    span = sum((n - 1) * st for n, st in zip(a.shape, a.strides)) + 1
    raw = np.lib.stride_tricks.as_strided(a, shape=(span,), strides=(1,))
    t = torch.from_numpy(np.array(raw)).cuda()
    return torch.as_strided(t, a.shape, a.strides)


def _run_device(gt_np, plan, anc_np=None, mode="auto", fmt="u8"):
    dplan = engine.DevicePlan(plan)
    gt = _to_device_keep_strides(gt_np)
    anc = _to_device_keep_strides(anc_np) if anc_np is not None else None
    flags = torch.zeros(1, dtype=torch.int32, device="cuda")
    out = engine.transform_device(gt, dplan, ancestry=anc, pack_mode=MODES[mode], fmt=fmt, flags=flags)
    torch.cuda.synchronize()
    return out, int(flags.item())


def _eligible(mode, lay, n, p):
    if mode == "vm":
        return lay == "F" and n % 8 == 0
    if mode == "sm":
        return lay == "C" and p % 2 == 0
    if mode == "sm2":
        return lay == "C" and p % 8 == 0
    return True


def _case_plan(case):
    haps, gts = _cases.build_objects(case, NS)
    anc = case["kind"] == "ancestry"
    return plan_from_haplotypes(list(haps.data.values()), gts, ancestry=anc)


@pytest.mark.parametrize("mode", ["generic", "vm", "sm", "sm2", "auto"])
@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: p.stem)
def test_golden_every_kernel(path, mode):
    case = _cases.load_case(path)
    if case["kind"] == "ancestry" and case["gt"].shape[2] != 2:
        pytest.skip("reference rejects a phase column in the ancestry path")
    plan = _case_plan(case)
    ran = 0
    for lay, gt in _layouts(case["gt"]).items():
        n, p = gt.shape[:2]
        if not _eligible(mode, lay, n, p):
            continue
        anc = None
        if case["kind"] == "ancestry":
            anc = _layouts(case["ancestry"])[lay]
        out, _ = _run_device(gt, plan, anc, mode)
        np.testing.assert_array_equal(out.cpu().numpy().view(np.bool_), case["expected_plural"], err_msg=f"{lay}/{mode}")
        ran += 1
    if not ran:
        pytest.skip("no layout of this case is eligible for this kernel")


def _random_arrays(rng, n, p, H, max_k=6, multi=False, missing=0.0, n_pops=0, long_hap=0):
    gt = rng.integers(0, 3 if multi else 2, size=(n, p, 2), dtype=np.uint8)
    if missing:
        m = rng.random((n, p, 2)) < missing
        gt[m] = np.where(rng.random(int(m.sum())) < 0.5, 255, 254)
    k = rng.integers(0, min(max_k, p) + 1, size=H)
    if long_hap:
        k[H // 2] = min(long_hap, p)
    hap_off = np.concatenate([[0], np.cumsum(k)])
    ref_col = np.concatenate([np.sort(rng.choice(p, size=int(x), replace=False)) for x in k] + [np.zeros(0, dtype=np.int64)])
    ref_allele = rng.integers(0, 3 if multi else 2, size=len(ref_col))
    anc = hap_label = ref_label = None
    if n_pops:
        anc = np.repeat(rng.integers(0, n_pops, size=(n, (p + 6) // 7, 2), dtype=np.uint8), 7, axis=1)[:, :p]
        anc = np.ascontiguousarray(anc)
        hap_label = rng.integers(0, n_pops, size=H)
        ref_label = np.repeat(hap_label, k)
    plan = plan_from_refs(ref_col, ref_allele, hap_off, p, ref_label)
    idxs = [np.arange(hap_off[h], hap_off[h + 1]) for h in range(H)]
    want = orc.transform_core(
        gt[:, ref_col.astype(np.int64)], ref_allele.astype(np.uint8), idxs,
        anc[:, ref_col.astype(np.int64)] if n_pops else None, hap_label.astype(np.uint8) if n_pops else None,
    )
    return gt, anc, plan, want


SHAPES = [
    (1, 4, 1), (31, 8, 7), (32, 12, 8), (33, 16, 9), (1023, 20, 127), (1024, 24, 128), (1025, 28, 129),
    (2504, 64, 300), (4096, 132, 16), (3000, 260, 257), (2100, 136, 64), (5000, 72, 40),
]


def test_biallelic_sample_major_clean_and_padded_rows():
    """The fast kernel on clean biallelic data (no invalid plane), on a row pitch larger than
    2p (device copies of host arrays are pitched), and with an odd column count."""
    rng = np.random.default_rng(9)
    for n, p, H in ((3000, 77, 50), (1100, 250, 30)):
        gt, _, plan, want = _random_arrays(rng, n, p, H, multi=False, missing=0.0)
        pitch = (2 * p + 15) // 16 * 16 + 32
        buf = torch.zeros((n, pitch), dtype=torch.uint8, device="cuda")
        buf[:, : 2 * p] = torch.from_numpy(gt.reshape(n, 2 * p)).cuda()
        view = torch.as_strided(buf, (n, p, 2), (pitch, 2, 1))
        dplan = engine.DevicePlan(plan)
        for mode in ("sm2", "sm", "auto"):
            out = engine.transform_device(view, dplan, pack_mode=MODES[mode])
            torch.cuda.synchronize()
            np.testing.assert_array_equal(out.cpu().numpy().view(np.bool_), want, err_msg=mode)


@pytest.mark.parametrize("mode", ["generic", "vm", "sm", "sm2"])
@pytest.mark.parametrize("n,p,H", SHAPES)
def test_random_vs_oracle(n, p, H, mode):
    rng = np.random.default_rng(n * 7919 + p * 31 + H)
    gt, _, plan, want = _random_arrays(rng, n, p, H, multi=True, missing=0.03)
    lay = {"generic": "C3", "vm": "F", "sm": "C", "sm2": "C"}[mode]
    if not _eligible(mode, lay, n, p):
        # pad the sample axis in memory so that variant rows are 16-byte aligned
        if mode == "vm":
            npad = (n + 7) // 8 * 8
            buf = np.zeros((p, npad, 2), dtype=np.uint8)
            buf[:, :n] = gt.transpose(1, 0, 2)
            arr = buf.transpose(1, 0, 2)[:n]
        else:
            pytest.skip("layout not eligible")
    else:
        arr = _layouts(gt)[lay]
    out, flags = _run_device(arr, plan, None, mode)
    np.testing.assert_array_equal(out.cpu().numpy().view(np.bool_), want)
    ref_cols = plan.var_col.astype(np.int64)
    sub = gt[:, ref_cols]
    want_flags = (1 if (sub >= 254).any() else 0) | (2 if ((sub >= 2) & (sub < 254)).any() else 0)
    assert flags == want_flags


@pytest.mark.parametrize("mode", ["generic", "vm", "sm", "sm2"])
@pytest.mark.parametrize("n,p,H", [(40, 8, 5), (1032, 40, 70), (2504, 64, 200), (3104, 200, 333)])
def test_random_ancestry_vs_oracle(n, p, H, mode):
    rng = np.random.default_rng(n + p + H)
    gt, anc, plan, want = _random_arrays(rng, n, p, H, n_pops=4, missing=0.01 if H != 200 else 0.0)
    lay = {"generic": "C3", "vm": "F", "sm": "C", "sm2": "C"}[mode]
    out, _ = _run_device(_layouts(gt)[lay], plan, _layouts(anc)[lay], mode)
    np.testing.assert_array_equal(out.cpu().numpy().view(np.bool_), want)


@pytest.mark.parametrize("case", ["labels_ge_8_in_keys", "labels_ge_8_in_data_only"])
def test_ancestry_label_ranges_sample_major(case):
    """The single-pass ancestry pack handles labels < 8; bigger labels in the DATA must never
    match (invalid plane), bigger labels in the KEYS fall back to the per-key pass."""
    rng = np.random.default_rng(77)
    n, p, H = 1504, 40, 60
    gt, anc, plan, want = _random_arrays(rng, n, p, H, n_pops=12, missing=0.02)
    if case == "labels_ge_8_in_data_only":
        # same data, haplotype populations restricted to 0..7
        k = np.diff(plan.hap_off.astype(np.int64))
        ref_col = plan.key_col().astype(np.int64)[plan.hap_key]
        ref_allele = plan.key_allele[plan.hap_key].astype(np.int64)
        hap_label = rng.integers(0, 8, size=H)
        plan = plan_from_refs(ref_col, ref_allele, plan.hap_off.astype(np.int64), p, np.repeat(hap_label, k))
        idxs = [np.arange(plan.hap_off[h], plan.hap_off[h + 1]) for h in range(H)]
        want = orc.transform_core(gt[:, ref_col], ref_allele.astype(np.uint8), idxs, anc[:, ref_col], hap_label.astype(np.uint8))
        assert plan.key_label.max() < 8 and anc.max() >= 8
    else:
        assert plan.key_label.max() >= 8
    for mode, lay in (("sm2", "C"), ("sm", "C"), ("vm", "F"), ("generic", "C3")):
        out, _ = _run_device(_layouts(gt)[lay], plan, _layouts(anc)[lay], mode)
        np.testing.assert_array_equal(out.cpu().numpy().view(np.bool_), want, err_msg=mode)


@pytest.mark.parametrize("seed", range(24))
def test_fuzz_host_api_vs_oracle(seed):
    """Randomised differential test through the host-array API: random shapes, physical
    layouts, dtype, key mix (multi-allelic, missing, ancestry with small/large labels, empty and
    long haplotypes), chunking -- always bit-exact against the oracle."""
    rng = np.random.default_rng(9000 + seed)
    n = int(rng.choice([1, 7, 33, 200, 1024, 1500, 2100, 4097]))
    p = int(rng.integers(1, 90))
    H = int(rng.choice([1, 3, 8, 17, 128, 130, 400]))
    anc_kind = int(rng.integers(0, 3))  # 0 none, 1 labels < 8, 2 labels up to 19
    gt, anc, plan, want = _random_arrays(
        rng, n, p, H, max_k=int(rng.integers(1, 8)), multi=bool(rng.integers(0, 2)),
        missing=float(rng.choice([0.0, 0.0, 0.01, 0.2])), n_pops=[0, 5, 20][anc_kind],
        long_hap=int(rng.choice([0, 0, 60])),
    )
    lay = str(rng.choice(["C", "F", "C3", "F3"]))
    arr = _layouts(gt)[lay]
    if anc is None and rng.integers(0, 3) == 0 and gt.max() <= 1:
        arr = arr.astype(np.bool_)  # check_biallelic output (haptools/data/genotypes.py:528)
    a = _layouts(anc)[lay] if anc is not None else None
    chunk = int(rng.choice([0, 1024, 2048])) or None
    got = engine.transform_host(arr, plan, ancestry=a, chunk_samples=chunk)
    assert got.dtype == np.bool_ and got.flags.c_contiguous
    np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("n", [8, 1024, 2504, 3000])
def test_variant_major_biallelic_single_pass(n):
    """Clean biallelic data in the VCF-reader layout: single-pass path of the variant-major
    kernel (keys (0), (1), (0, 1)); a few dirty variants fall back to the per-key compare."""
    rng = np.random.default_rng(n)
    p, H = 50, 64
    gt, _, plan, want = _random_arrays(rng, n, p, H, multi=False, missing=0.0)
    out, flags = _run_device(_layouts(gt)["F"], plan, None, "vm")
    np.testing.assert_array_equal(out.cpu().numpy().view(np.bool_), want)
    assert flags == 0
    gt[rng.integers(0, n), ::7, :] = 255  # missing calls in every 7th variant
    want = orc.transform_from_csr(gt, plan.key_col().astype(np.int64), plan.key_allele,
                                  plan.hap_off.astype(np.int64), plan.hap_key.astype(np.int64))
    out, _ = _run_device(_layouts(gt)["F"], plan, None, "vm")
    np.testing.assert_array_equal(out.cpu().numpy().view(np.bool_), want)


def test_long_haplotype_unstaged_keys():
    """A haplotype block whose key list exceeds the shared-memory stage (3072 keys)."""
    rng = np.random.default_rng(5)
    gt, _, plan, want = _random_arrays(rng, 2000, 4200, 40, long_hap=4000)
    out, _ = _run_device(gt, plan)
    np.testing.assert_array_equal(out.cpu().numpy().view(np.bool_), want)
    # make the 4000-allele haplotype present in a few samples
    off = plan.hap_off.astype(np.int64)
    keys = plan.hap_key[off[20] : off[21]].astype(np.int64)
    assert len(keys) == 4000
    cols = plan.key_col().astype(np.int64)[keys]
    for s in range(7):
        gt[s, cols, :] = plan.key_allele[keys][:, None]
    want = orc.transform_from_csr(gt, plan.key_col().astype(np.int64), plan.key_allele, off, plan.hap_key.astype(np.int64))
    assert want[:7, 20].all()
    out, _ = _run_device(gt, plan)
    np.testing.assert_array_equal(out.cpu().numpy().view(np.bool_), want)


def test_empty_inputs():
    plan = plan_from_refs(np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64), np.zeros(4, dtype=np.int64), 5)
    gt = np.zeros((10, 5, 2), dtype=np.uint8)
    out = engine.transform_host(gt, plan)
    assert out.shape == (10, 3, 2) and out.all()  # haplotypes without alleles are all-True
    out = engine.transform_host(gt[:0], plan)
    assert out.shape == (0, 3, 2)
    plan0 = plan_from_refs(np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64), np.zeros(1, dtype=np.int64), 5)
    assert engine.transform_host(gt, plan0).shape == (10, 0, 2)


def test_bool_dtype_input():
    case = _cases.load_case(_cases.GOLDEN_DIR / "rand_plain_bool.npz")
    assert case["gt"].dtype == np.bool_
    haps, gts = _cases.build_objects(case, NS)
    np.testing.assert_array_equal(haps.transform(gts).data, case["expected_plural"])


# ------------------------------------------------------------------ packed result, counts
def _unpack_bits(bits, n, H):
    b = bits.cpu().numpy().view(np.uint32)  # (tiles, H, 32, 2)
    t = b.shape[0]
    bytes_ = b.view(np.uint8).reshape(t, H, 32, 2, 4)
    u = np.unpackbits(bytes_, axis=-1, bitorder="little").reshape(t, H, 32, 2, 32)  # [t,h,w,c,bit]
    return u.transpose(0, 2, 4, 1, 3).reshape(t * 1024, H, 2)[:n].astype(np.bool_)


@pytest.mark.parametrize("n,p,H", [(33, 16, 9), (2504, 64, 300)])
def test_bits_and_counts(n, p, H):
    rng = np.random.default_rng(n)
    gt, _, plan, want = _random_arrays(rng, n, p, H, max_k=3)
    bits, _ = _run_device(gt, plan, fmt="bits")
    np.testing.assert_array_equal(_unpack_bits(bits, n, H), want)
    counts = engine.count_bits(bits, n, H)
    np.testing.assert_array_equal(counts.cpu().numpy(), want.sum(axis=(0, 2)))


def test_fused_allele_counts_and_maf():
    """check_maf's allele counts fall out of the transform kernel (N2)."""
    rng = np.random.default_rng(12)
    for n, p, H in ((2504, 64, 300), (1000, 16, 7)):
        gt, _, plan, want = _random_arrays(rng, n, p, H, max_k=3)
        dplan = engine.DevicePlan(plan)
        counts = torch.full((H,), -7, dtype=torch.int64, device="cuda")
        out = engine.transform_device(torch.from_numpy(gt).cuda(), dplan, counts=counts)
        torch.cuda.synchronize()
        np.testing.assert_array_equal(out.cpu().numpy().view(np.bool_), want)
        np.testing.assert_array_equal(counts.cpu().numpy(), want.sum(axis=(0, 2)))
        # the reference's formula (haptools/data/genotypes.py:600-602) on the host result
        ref_af = want.astype(np.bool_).sum(axis=(0, 2)) / (2 * n)
        np.testing.assert_array_equal(engine.maf_from_counts(counts, n), np.array([ref_af, 1 - ref_af]).min(axis=0))


# ------------------------------------------------------------------ pgenlib int32 buffers
@pytest.mark.parametrize("form", ["narrow", "generic"])
@pytest.mark.parametrize("n", [1500, 1501, 2048, 37])
def test_pack_i32_chunks(monkeypatch, form, n):
    """ht_pack_i32 in both forms (csrc/ht_pack.cu): int32 rows narrowed to bytes + the variant-major kernel (the
    default), and the one-kernel generic form; odd sample counts exercise the narrowing kernel's 2-value tail and
    the 16-byte alignment rule of its rows, flags are compared with the definition."""
    monkeypatch.setenv("HT_I32_KERNEL", form)
    rng = np.random.default_rng(11 + n)
    p, H = 96, 50
    gt, _, plan, want = _random_arrays(rng, n, p, H, multi=True, missing=0.05)
    # what PgenReader.read_alleles_list hands out: (variants, 2n) int32, -9 = missing
    alleles = gt.transpose(1, 0, 2).reshape(p, 2 * n).astype(np.int32)
    alleles[alleles >= 254] = -9
    dplan = engine.DevicePlan(plan)
    planes = engine.alloc_planes(n, plan.n_keys, "cuda")
    planes.zero_()
    lib = _cuda.lib()
    chunk = 40
    vk = plan.var_key_off.astype(np.int64)
    flags = torch.zeros(1, dtype=torch.int32, device="cuda")
    for c0 in range(0, p, chunk):
        c1 = min(p, c0 + chunk)
        sel = np.nonzero((plan.var_col >= c0) & (plan.var_col < c1))[0]
        if len(sel) == 0:
            continue
        v0, v1 = sel[0], sel[-1] + 1
        buf = torch.from_numpy(alleles[c0:c1].copy()).cuda()
        rows = torch.from_numpy((plan.var_col[v0:v1].astype(np.int64) - c0).astype(np.int32)).cuda()
        koff = torch.from_numpy((vk[v0 : v1 + 1] - vk[v0]).astype(np.int32)).cuda()
        kall = torch.from_numpy(plan.key_allele[vk[v0] : vk[v1]].copy()).cuda()
        _cuda.check(
            lib.ht_pack_i32(buf.data_ptr(), 2 * n, n, rows.data_ptr(), koff.data_ptr(), kall.data_ptr(),
                            v1 - v0, int(vk[v0]), plan.n_keys, planes.data_ptr(), flags.data_ptr(), 0),
            "ht_pack_i32",
        )
    torch.cuda.synchronize()
    out = torch.empty((n, engine.out_pitch(H)), dtype=torch.uint8, device="cuda")
    engine.transform_u8(dplan, planes, n, out.data_ptr(), out.stride(0))
    torch.cuda.synchronize()
    np.testing.assert_array_equal(out[:, : 2 * H].cpu().numpy().reshape(n, H, 2).view(np.bool_), want)
    sub = gt[:, plan.var_col.astype(np.int64)]
    want_flags = (1 if (sub >= 254).any() else 0) | (2 if ((sub >= 2) & (sub < 254)).any() else 0)
    assert int(flags.item()) == want_flags


@pytest.mark.parametrize("chunk", [7, 64, 1000])
def test_pgen_stream(chunk):
    """pgenlib-style int32 chunk buffers -> planes, never building gts.data (config 5 / N1)."""
    rng = np.random.default_rng(21)
    n, p_file, H = 1300, 400, 90
    read = np.sort(rng.choice(p_file, size=150, replace=False)).astype(np.uint32)  # variants the .hap needs
    gt, _, plan, want = _random_arrays(rng, n, len(read), H, multi=True, missing=0.04)
    file_alleles = rng.integers(0, 2, size=(p_file, 2 * n), dtype=np.int32)
    as_i32 = gt.transpose(1, 0, 2).reshape(len(read), 2 * n).astype(np.int32)
    as_i32[as_i32 == 255] = -9  # pgenlib's missing code; 254 stays 254, as in the reference's cast
    file_alleles[read] = as_i32
    reader = stream.ArrayPgenReader(file_alleles)
    stats = {}
    got = stream.transform_pgen_stream(reader, read, n, plan, chunk_variants=chunk, stats=stats)
    np.testing.assert_array_equal(got, want)
    assert stats["h2d_bytes"] == plan.n_vars * 2 * n * 4  # only referenced variants travel
    bits = stream.transform_pgen_stream(reader, read, n, plan, chunk_variants=chunk, fmt="bits")
    from haptools_b200 import dist as hdist
    np.testing.assert_array_equal(hdist.unpack_result_bits(bits, n), want)


def test_pgen_stream_many_chunks_ring():
    """A longer stream (24 chunks through 3 pinned host buffers and 2 device buffers) from a
    reader that fills the buffers like pgenlib does; checked against the oracle."""
    n, V, H = 6000, 1536, 400
    plan = synth.synth_plan(V, H, seed=5, max_alleles=6)
    gt = synth.synth_gt(n, V, seed=8, mode=1, missing_per_64k=300)
    rows = gt.transpose(1, 0, 2).reshape(V, 2 * n).astype(np.int32)
    rows[rows == 255] = -9
    calls = []

    def fill(idx, out):
        calls.append(len(idx))
        np.take(rows, idx, axis=0, out=out)

    reader = stream.ArrayPgenReader(fill=fill, n_samples=n)
    stats = {}
    got = stream.transform_pgen_stream(reader, np.arange(V, dtype=np.uint32), n, plan, chunk_variants=64, stats=stats)
    want = orc.transform_from_csr(gt, plan.key_col().astype(np.int64), plan.key_allele,
                                  plan.hap_off.astype(np.int64), plan.hap_key.astype(np.int64))
    np.testing.assert_array_equal(got, want)
    assert stats["chunks"] == len(calls) == -(-plan.n_vars // 64) and sum(calls) == plan.n_vars


@pytest.mark.parametrize("source", ["host_matrix", "device_matrix", "callable"])
@pytest.mark.parametrize("chunk", [9, 64, 1000])
def test_pgen_stream_with_ancestry(chunk, source):
    """`transform --ancestry` on PGEN input (haptools/transform.py:610-660): int32 chunks + the local-ancestry
    labels of each chunk's variants -> ancestry planes (ht_pack_i32_anc), against the oracle."""
    rng = np.random.default_rng(23)
    n, H = 1300, 90
    gt, anc, plan, want = _random_arrays(rng, n, 150, H, missing=0.03, n_pops=5)
    rows = gt.transpose(1, 0, 2).reshape(150, 2 * n).astype(np.int32)
    rows[rows == 255] = -9
    reader = stream.ArrayPgenReader(rows)
    src = {"host_matrix": anc, "device_matrix": torch.from_numpy(anc).cuda(),
           "callable": lambda cols: torch.from_numpy(np.ascontiguousarray(anc[:, cols])).cuda()}[source]
    got = stream.transform_pgen_stream(reader, np.arange(150, dtype=np.uint32), n, plan, chunk_variants=chunk, ancestry=src)
    np.testing.assert_array_equal(got, want)
    with pytest.raises(ValueError, match="given together"):
        stream.transform_pgen_stream(reader, np.arange(150, dtype=np.uint32), n, plan, chunk_variants=chunk)


def test_pgen_stream_with_breakpoints_on_the_device():
    """The whole of transform_haps' ancestry branch without its two big host arrays: labels come from the .bp blocks
    on the device, chunk by chunk (Breakpoints.population_provider -> ht_bp_ancestry), genotypes from int32 chunks.
    Expected: the oracle on the (n, p, 2) matrices the reference would have built (breakpoints oracle + transform
    oracle)."""
    from haptools_b200.data.breakpoints import HapBlock
    from oracle import breakpoints_oracle as borc

    rng = np.random.default_rng(29)
    n, p, H, pops = 700, 260, 120, ("YRI", "CEU", "PEL")
    bps = data.Breakpoints(fname=None)
    bps.data = {}
    for s in range(n):
        strands = []
        for c in range(2):
            k = int(rng.integers(1, 6))
            ends = np.sort(rng.choice(np.arange(10, 5000), size=k, replace=False))
            ends[-1] = 5000
            strands.append(np.array([(str(rng.choice(pops)), "1", int(e), e / 1e4) for e in ends], dtype=HapBlock))
        bps.data[f"S{s}"] = strands
    bps.encode(labels=pops)
    variants = np.zeros(p, dtype=[("chrom", "U10"), ("pos", np.uint32)])
    variants["chrom"], variants["pos"] = "1", np.sort(rng.choice(np.arange(1, 5001), size=p, replace=False))
    anc = borc.population_array(bps.data, variants)
    gt = rng.integers(0, 2, size=(n, p, 2), dtype=np.uint8)
    gt[rng.random((n, p, 2)) < 0.02] = 255
    k = rng.integers(1, 5, size=H)
    hap_off = np.concatenate([[0], np.cumsum(k)])
    ref_col = np.concatenate([np.sort(rng.choice(p, size=int(x), replace=False)) for x in k])
    ref_allele = rng.integers(0, 2, size=len(ref_col))
    hap_label = rng.integers(0, len(pops), size=H)
    plan = plan_from_refs(ref_col, ref_allele, hap_off, p, np.repeat(hap_label, k))
    idxs = [np.arange(hap_off[h], hap_off[h + 1]) for h in range(H)]
    want = orc.transform_core(gt[:, ref_col], ref_allele.astype(np.uint8), idxs, anc[:, ref_col], hap_label.astype(np.uint8))
    rows = gt.transpose(1, 0, 2).reshape(p, 2 * n).astype(np.int32)
    rows[rows == 255] = -9
    reader = stream.ArrayPgenReader(rows)
    for chunk in (17, 4096):
        provider = bps.population_provider(variants)
        got = stream.transform_pgen_stream(reader, np.arange(p, dtype=np.uint32), n, plan, chunk_variants=chunk, ancestry=provider)
        np.testing.assert_array_equal(got, want)
    assert want.any()
    # a variant beyond every block: the reference's error, raised after the stream
    far = variants.copy()
    far["pos"][int(plan.var_col[3])] = 6000
    with pytest.raises(ValueError, match="do not specify an ancestry"):
        stream.transform_pgen_stream(reader, np.arange(p, dtype=np.uint32), n, plan, chunk_variants=64,
                                     ancestry=bps.population_provider(far))


# ------------------------------------------------------------------ host pipeline
@pytest.mark.parametrize("lay", ["C", "F", "C3", "F3"])
def test_host_pipeline_chunked(lay):
    rng = np.random.default_rng(3)
    n, p, H = 5000, 48, 77
    gt, anc, plan, want = _random_arrays(rng, n, p, H, n_pops=3 if lay in ("C", "F") else 0, missing=0.02)
    arr = _layouts(gt)[lay]
    a = _layouts(anc)[lay] if anc is not None else None
    whole = engine.transform_host(arr, plan, ancestry=a)
    np.testing.assert_array_equal(whole, want)
    chunked = engine.transform_host(arr, plan, ancestry=a, chunk_samples=1024)
    np.testing.assert_array_equal(chunked, want)


def test_relayout_kernel_and_phase_column_layouts_take_the_fast_path(monkeypatch):
    """Arrays that interleave the phase column are re-laid-out on the device and then packed by
    the fast kernels (never by the generic one)."""
    rng = np.random.default_rng(6)
    n, p, H = 2100, 40, 33
    gt, _, plan, want = _random_arrays(rng, n, p, H, missing=0.02)
    modes = []
    real_pack = engine.pack

    def spy(dplan, gt_ptr, gt_strides, *a, **kw):
        modes.append(tuple(int(x) for x in gt_strides))
        return real_pack(dplan, gt_ptr, gt_strides, *a, **kw)

    monkeypatch.setattr(engine, "pack", spy)
    for lay in ("C3", "F3"):
        np.testing.assert_array_equal(engine.transform_host(_layouts(gt)[lay], plan), want)
    assert all(engine._fast_layout(st) for st in modes), modes
    # the kernel itself, both directions
    src = torch.from_numpy(np.ascontiguousarray(_layouts(gt)["C3"].base if _layouts(gt)["C3"].base is not None else gt)).cuda()
    c3 = _to_device_keep_strides(_layouts(gt)["C3"])
    dst = torch.zeros((n, 96), dtype=torch.uint8, device="cuda")
    _cuda.check(_cuda.lib().ht_relayout_u8(c3.data_ptr(), *c3.stride(), n, p, dst.data_ptr(), 96, 2, 0), "relayout")
    np.testing.assert_array_equal(dst[:, : 2 * p].cpu().numpy().reshape(n, p, 2), gt)
    dstv = torch.zeros((p, 2 * n + 8), dtype=torch.uint8, device="cuda")
    _cuda.check(_cuda.lib().ht_relayout_u8(c3.data_ptr(), *c3.stride(), n, p, dstv.data_ptr(), 2, 2 * n + 8, 0), "relayout")
    np.testing.assert_array_equal(dstv[:, : 2 * n].cpu().numpy().reshape(p, n, 2).transpose(1, 0, 2), gt)


def test_host_exotic_strides():
    rng = np.random.default_rng(4)
    gt, _, plan, want = _random_arrays(rng, 300, 20, 30)
    rev = gt[::-1]  # negative sample stride
    np.testing.assert_array_equal(engine.transform_host(rev, plan), want[::-1])
    strand_major = np.ascontiguousarray(gt.transpose(2, 0, 1)).transpose(1, 2, 0)
    np.testing.assert_array_equal(engine.transform_host(strand_major, plan), want)


# ------------------------------------------------------------------ synthetic generator
@pytest.mark.parametrize("mode,miss", [(0, 0), (1, 0), (0, 700), (1, 64)])
def test_synth_device_equals_numpy(mode, miss):
    n, p, off = 300, 257, 12345
    for shape, perm in (((n, p, 2), None), ((p, n, 2), (1, 0, 2))):
        t = torch.zeros(shape, dtype=torch.uint8, device="cuda")
        v = t if perm is None else t.permute(*perm)
        _cuda.check(_cuda.lib().ht_synth_gt(v.data_ptr(), *v.stride(), n, p, off, 99, mode, miss, 0), "synth")
        torch.cuda.synchronize()
        np.testing.assert_array_equal(v.cpu().numpy(), synth.synth_gt(n, p, 99, mode, miss, sample_offset=off))
    t = torch.zeros((n, p, 2), dtype=torch.uint8, device="cuda")
    _cuda.check(_cuda.lib().ht_synth_ancestry(t.data_ptr(), *t.stride(), n, p, off, 7, 5, 50, 0), "synth_anc")
    np.testing.assert_array_equal(t.cpu().numpy(), synth.synth_ancestry(n, p, 7, 5, 50, sample_offset=off))


# ------------------------------------------------------------------ BASELINE-size properties
def _device_cohort(n, p, seed, layout, mode=1):
    if layout == "sample_major":
        gt = torch.empty((n, p, 2), dtype=torch.uint8, device="cuda")
    else:
        gt = torch.empty((p, n, 2), dtype=torch.uint8, device="cuda").permute(1, 0, 2)
    _cuda.check(_cuda.lib().ht_synth_gt(gt.data_ptr(), *gt.stride(), n, p, 0, seed, mode, 0, 0), "synth")
    return gt


def test_config_1000g_scale_slice_and_checksums():
    """BASELINE configs[1]: 2,504 samples x 1M variants x 10k haplotypes (variant-major)."""
    n, p, H = 2504, 1_000_000, 10_000
    plan = synth.synth_plan(p, H, seed=1)
    dplan = engine.DevicePlan(plan)
    gt = _device_cohort(n, p, 0, "variant_major")
    out = engine.transform_device(gt, dplan)
    bits = engine.transform_device(gt, dplan, fmt="bits")
    counts = engine.count_bits(bits, n, H)
    torch.cuda.synchronize()
    got = out.cpu().numpy().view(np.bool_)
    # every output byte is 0 or 1 and the two result formats agree (checksum of checksums)
    assert out.max().item() <= 1
    np.testing.assert_array_equal(counts.cpu().numpy(), got.sum(axis=(0, 2)))
    assert got.any() and not got.all()
    # oracle on a sample slice, regenerated on the host from the same hash
    sl = slice(1000, 1000 + 300)
    cols = plan.var_col.astype(np.int64)
    sub = synth.synth_gt(300, p, 0, mode=1, sample_offset=1000, variants=cols)
    key_in_sub = np.searchsorted(cols, plan.key_col().astype(np.int64))
    want = orc.transform_from_csr(sub, key_in_sub, plan.key_allele, plan.hap_off.astype(np.int64), plan.hap_key.astype(np.int64))
    np.testing.assert_array_equal(got[sl], want)


def test_config_ukb_scale_shard_slice_and_checksums():
    """BASELINE configs[2], the shard one of 8 GPUs owns: 62,500 of 500k samples x 50k
    variants x 100k haplotypes (sample-major)."""
    n, p, H = 62_500, 50_000, 100_000
    plan = synth.synth_plan(p, H, seed=2)
    dplan = engine.DevicePlan(plan)
    gt = _device_cohort(n, p, 3, "sample_major")
    out = engine.transform_device(gt, dplan)
    bits = engine.transform_device(gt, dplan, fmt="bits")
    counts = engine.count_bits(bits, n, H).cpu().numpy()
    torch.cuda.synchronize()
    col_sums = out.sum(dim=(0, 2), dtype=torch.int64).cpu().numpy()
    np.testing.assert_array_equal(counts, col_sums)
    assert out.max().item() <= 1
    for s0 in (0, 31_000, n - 64):
        sub = synth.synth_gt(64, p, 3, mode=1, sample_offset=s0)
        want = orc.transform_from_csr(sub, plan.key_col().astype(np.int64), plan.key_allele,
                                      plan.hap_off.astype(np.int64), plan.hap_key.astype(np.int64))
        np.testing.assert_array_equal(out[s0 : s0 + 64].cpu().numpy().view(np.bool_), want)


def test_config_ancestry_scale_slices_and_checksums():
    """BASELINE configs[3] at its stated shape: 100k samples x 200k variants x 50k haplotypes,
    5 populations -- GT 40 GB + ANC 40 GB resident in HBM, generated on the device; the oracle (haptools/
    transform.py:137-148 restated) on three 64-sample slices regenerated on the host from the same hash, and the
    size-independent checks on everything: bytes are 0/1, fused counts == column sums == popcount of the packed
    result, and dropping the ancestry condition can only add presence."""
    n, p, H, n_pops = 100_000, 200_000, 50_000, 5
    free = torch.cuda.mem_get_info()[0]
    if free < 115 << 30:
        n = 25_000  # a quarter of the cohort on a smaller / busier device; same plan, same checks
    plan = synth.synth_plan(p, H, seed=4, n_pops=n_pops)
    dplan = engine.DevicePlan(plan)
    lib = _cuda.lib()
    gt = torch.empty((n, p, 2), dtype=torch.uint8, device="cuda")
    anc = torch.empty((n, p, 2), dtype=torch.uint8, device="cuda")
    _cuda.check(lib.ht_synth_gt(gt.data_ptr(), *gt.stride(), n, p, 0, 5, 1, 0, 0), "synth")
    _cuda.check(lib.ht_synth_ancestry(anc.data_ptr(), *anc.stride(), n, p, 0, 6, n_pops, 2000, 0), "synth_anc")
    counts = torch.empty(H, dtype=torch.int64, device="cuda")
    out = engine.transform_device(gt, dplan, ancestry=anc, counts=counts)
    torch.cuda.synchronize()
    assert out.max().item() <= 1
    sums = torch.zeros(H, dtype=torch.int64, device="cuda")
    for lo in range(0, n, 20_000):
        sums += out[lo: lo + 20_000].sum(dim=(0, 2), dtype=torch.int64)
    assert torch.equal(counts, sums) and int(sums.sum()) > 0
    bits = engine.transform_device(gt, dplan, ancestry=anc, fmt="bits")
    assert torch.equal(engine.count_bits(bits, n, H), sums)
    del bits
    # the oracle on slices: first, middle, last samples
    ref_col, ref_allele, hap_off, hap_label = synth.synth_haplotype_refs(p, H, 4, n_pops=n_pops)
    plain = plan_from_refs(ref_col, ref_allele, hap_off, p)
    cols = plain.var_col.astype(np.int64)
    key_in_sub = np.searchsorted(cols, plain.key_col().astype(np.int64))
    for s0 in (0, n // 2 + 37, n - 64):
        sub = synth.synth_gt(64, p, 5, mode=1, sample_offset=s0, variants=cols)
        sub_anc = synth.synth_ancestry(64, p, 6, n_pops, 2000, sample_offset=s0, variants=cols)
        want = orc.transform_from_csr(sub, key_in_sub, plain.key_allele, plain.hap_off.astype(np.int64),
                                      plain.hap_key.astype(np.int64), sub_anc, hap_label.astype(np.uint8))
        np.testing.assert_array_equal(out[s0: s0 + 64].cpu().numpy().view(np.bool_), want)
        assert want.any()
    # monotonicity: the plain transform of the same genotypes is present wherever the ancestry-aware one is
    m = min(n, 8192)
    plain_out = engine.transform_device(gt[:m], engine.DevicePlan(plain))
    assert bool(((out[:m] & ~plain_out) == 0).all()) and int(plain_out.sum()) >= int(out[:m].sum())


def test_host_pipeline_round_robin_over_devices():
    """transform_host(devices=...): sample chunks dealt round-robin to the GPUs of the box (on a
    one-GPU box the same code path runs with a single device)."""
    rng = np.random.default_rng(12)
    n, p, H = 9000, 64, 90
    gt, anc, plan, want = _random_arrays(rng, n, p, H, n_pops=3, missing=0.02)
    for devices in ([0], "all"):
        got = engine.transform_host(gt, plan, ancestry=anc, chunk_samples=1024, devices=devices)
        np.testing.assert_array_equal(got, want)
    if torch.cuda.device_count() >= 2:
        got = engine.transform_host(gt, plan, ancestry=anc, devices=[1, 0])  # chunk size derived from the device count
        np.testing.assert_array_equal(got, want)
        with torch.cuda.device(1):
            got = engine.transform_host(gt[:, :, :2], plan_from_refs(*_plain_refs(plan), p), devices="all")
        assert got.shape == (n, H, 2)


def _plain_refs(plan):
    """(ref_col, ref_allele, hap_off) of a plan with its ancestry labels dropped."""
    key_col = plan.key_col().astype(np.int64)
    return key_col[plan.hap_key.astype(np.int64)], plan.key_allele[plan.hap_key.astype(np.int64)].astype(np.int64), plan.hap_off.astype(np.int64)


def test_size_independent_properties_at_scale():
    """Properties that need no oracle, at a size the oracle cannot reach (120k samples x 20k
    variants x 30k haplotypes = 7.2e9 hap-calls):
      * union: a haplotype made of the alleles of two others is present exactly where both are;
      * sample independence: a permuted cohort gives the permuted result;
      * counts fused into the kernel == popcount of the packed result == column sums."""
    n, p, H = 120_000, 20_000, 30_000
    dev = torch.device("cuda")
    ref_col, ref_allele, hap_off, _ = synth.synth_haplotype_refs(p, H, 7, max_alleles=8)
    # the last H/3 haplotypes are unions of two earlier ones
    third = H // 3
    rng = np.random.default_rng(1)
    pairs = rng.integers(0, 2 * third, size=(third, 2))
    cols, alls, offs = [ref_col[: hap_off[2 * third]]], [ref_allele[: hap_off[2 * third]]], list(hap_off[: 2 * third + 1])
    for a, b in pairs:
        idx = np.concatenate([np.arange(hap_off[a], hap_off[a + 1]), np.arange(hap_off[b], hap_off[b + 1])])
        cols.append(ref_col[idx]); alls.append(ref_allele[idx]); offs.append(offs[-1] + len(idx))
    plan = plan_from_refs(np.concatenate(cols), np.concatenate(alls), np.asarray(offs), p)
    H = plan.n_haps
    dplan = engine.DevicePlan(plan, dev)
    gt = torch.empty((n, p, 2), dtype=torch.uint8, device=dev)
    _cuda.check(_cuda.lib().ht_synth_gt(gt.data_ptr(), *gt.stride(), n, p, 0, 5, 1, 30, 0), "synth")
    counts = torch.empty(H, dtype=torch.int64, device=dev)
    out = engine.transform_device(gt, dplan, counts=counts)
    torch.cuda.synchronize()
    assert int(counts.sum()) > 0
    # union property, checked on the device column block by column block
    pa, pb = torch.from_numpy(pairs[:, 0]).to(dev), torch.from_numpy(pairs[:, 1]).to(dev)
    for lo in range(0, third, 2000):
        hi = min(third, lo + 2000)
        both = out[:, pa[lo:hi]] & out[:, pb[lo:hi]]
        assert torch.equal(out[:, 2 * third + lo : 2 * third + hi], both)
    # counts: fused == column sums == popcount of the packed form
    sums = torch.zeros(H, dtype=torch.int64, device=dev)
    for lo in range(0, n, 20_000):
        sums += out[lo : lo + 20_000].sum(dim=(0, 2), dtype=torch.int64)
    assert torch.equal(counts, sums)
    bits = engine.transform_device(gt, dplan, fmt="bits")
    assert torch.equal(engine.count_bits(bits, n, H), sums)
    # sample independence on a permuted slice of the cohort
    perm = torch.randperm(4096, device=dev)
    sub = engine.transform_device(gt[:4096][perm].contiguous(), dplan)
    assert torch.equal(sub, out[:4096][perm])


@pytest.mark.parametrize("lay", ["C", "F", "C3", "F3"])
def test_host_to_resident_bits(lay):
    """engine.transform_host_bits: chunked upload, packed result resident on the device, counts."""
    from haptools_b200 import dist as hdist

    rng = np.random.default_rng(31)
    n, p, H = 4500, 56, 61
    gt, anc, plan, want = _random_arrays(rng, n, p, H, n_pops=3 if lay in ("C", "F") else 0, multi=lay not in ("C", "F"), missing=0.02)
    arr = _layouts(gt)[lay]
    a = _layouts(anc)[lay] if anc is not None else None
    for chunk in (None, 1024, 2048):
        bits, counts = engine.transform_host_bits(arr, plan, ancestry=a, chunk_samples=chunk, counts=True)
        torch.cuda.synchronize()
        np.testing.assert_array_equal(hdist.unpack_result_bits(bits.cpu().numpy().view(np.uint32), n), want)
        np.testing.assert_array_equal(counts.cpu().numpy(), want.sum(axis=(0, 2)))
    assert engine.transform_host_bits(arr, plan, ancestry=a).shape == ((n + 1023) // 1024, H, 32, 2)
