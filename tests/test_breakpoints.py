"""
Local ancestry from breakpoints: Breakpoints.encode / population_array
(haptools/data/breakpoints.py:162-200, 259-330).  Fixtures were produced by the reference's own
class (tests/golden/make_golden_bp.py).
"""
import numpy as np
import pytest

import _cases
from haptools_b200 import data
from haptools_b200.data.breakpoints import HapBlock
from oracle import breakpoints_oracle as borc

BP = _cases.list_golden_bp()


def test_fixtures_present():
    assert len(BP) == 3


@pytest.mark.parametrize("path", BP, ids=lambda p: p.stem)
def test_encode_and_oracle_match_reference(path):
    bps, variants, meta, expected = _cases.load_bp_case(path, data.Breakpoints, HapBlock)
    order = meta["label_order"]
    bps.encode(labels=tuple(order) if order else None)
    assert bps.labels == meta["labels"]  # same integers as the reference's encode()
    with pytest.raises(ValueError, match="already been encoded"):
        bps.encode()
    got = borc.population_array(bps.data, variants)
    np.testing.assert_array_equal(got, expected)
    names = tuple(bps.data)[::2]
    np.testing.assert_array_equal(borc.population_array(bps.data, variants, samples=names), expected[::2])


def test_oracle_errors():
    bps, variants, meta, _ = _cases.load_bp_case(BP[1], data.Breakpoints, HapBlock)
    bps.encode()
    far = variants.copy()
    far["pos"][3] = 6000
    with pytest.raises(ValueError, match="do not specify an ancestry"):
        borc.population_array(bps.data, far)
    other = variants.copy()
    other["chrom"][0] = "chr9"
    with pytest.raises(ValueError, match="is absent in the breakpoints"):
        borc.population_array(bps.data, other)


# ------------------------------------------------------------------ device
@pytest.mark.gpu
@pytest.mark.parametrize("path", BP, ids=lambda p: p.stem)
def test_device_population_array(path):
    bps, variants, meta, expected = _cases.load_bp_case(path, data.Breakpoints, HapBlock)
    order = meta["label_order"]
    # un-encoded: label strings come back, like the reference
    inv = {v: k for k, v in meta["labels"].items()}
    strings = bps.population_array(variants)
    assert strings.dtype.kind == "U" and strings.shape == expected.shape
    assert all(strings[idx] == inv[int(expected[idx])] for idx in [(0, 0, 0), (1, 5, 1), (expected.shape[0] - 1, 7, 0)])
    bps.encode(labels=tuple(order) if order else None)
    got = bps.population_array(variants)
    assert got.dtype == np.uint8
    np.testing.assert_array_equal(got, expected)
    names = tuple(bps.data)[1::2]
    np.testing.assert_array_equal(bps.population_array(variants, samples=names), expected[1::2])
    t = bps.population_tensor(variants)
    assert t.is_cuda and tuple(t.shape) == expected.shape


@pytest.mark.gpu
def test_device_errors_match_reference_messages():
    bps, variants, meta, _ = _cases.load_bp_case(BP[1], data.Breakpoints, HapBlock)
    bps.encode()
    far = variants.copy()
    far["pos"][3] = 6000
    with pytest.raises(ValueError, match=r"The breakpoints for chromosome .* in sample S0_1 do not specify an ancestry"):
        bps.population_array(far)
    other = variants.copy()
    other["chrom"][0] = "chr9"
    with pytest.raises(ValueError, match=r"Chromosome chr9 in the genotypes is absent in the breakpoints for sample S0_1"):
        bps.population_array(other)


@pytest.mark.gpu
def test_bp_to_ancestry_transform_on_device():
    """.bp blocks -> ancestry tensor -> ancestry transform without the label matrix ever
    visiting the host (haptools/transform.py:644-660 + 91-149)."""
    import torch

    from haptools_b200 import engine
    from haptools_b200.plan import plan_from_refs
    from oracle import transform_oracle as orc

    bps, variants, meta, expected = _cases.load_bp_case(BP[1], data.Breakpoints, HapBlock)
    bps.encode()
    n, p = expected.shape[:2]
    rng = np.random.default_rng(1)
    gt = rng.integers(0, 2, size=(n, p, 2), dtype=np.uint8)
    H = 50
    k = rng.integers(1, 4, size=H)
    hap_off = np.concatenate([[0], np.cumsum(k)])
    ref_col = np.concatenate([np.sort(rng.choice(p, size=int(x), replace=False)) for x in k])
    ref_allele = rng.integers(0, 2, size=len(ref_col))
    hap_label = rng.integers(0, len(bps.labels), size=H)
    plan = plan_from_refs(ref_col, ref_allele, hap_off, p, np.repeat(hap_label, k))
    anc = bps.population_tensor(variants)
    out = engine.transform_device(torch.from_numpy(gt).cuda(), engine.DevicePlan(plan), ancestry=anc)
    idxs = [np.arange(hap_off[h], hap_off[h + 1]) for h in range(H)]
    want = orc.transform_core(gt[:, ref_col], ref_allele.astype(np.uint8), idxs, expected[:, ref_col], hap_label.astype(np.uint8))
    np.testing.assert_array_equal(out.cpu().numpy().view(np.bool_), want)
