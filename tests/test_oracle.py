"""
Pins the oracle (oracle/transform_oracle.py): against the reference's own literals for this
path, against the committed golden fixtures generated from the
unmodified reference (tests/golden/make_golden.py), and -- only where /root/reference is
mounted -- against the live reference on fresh random inputs.
"""
import types

import numpy as np
import pytest

import _cases
from oracle import ref_harness
from oracle import transform_oracle as orc

GOLDEN = _cases.list_golden()


class _NS:
    """Plain stand-ins for the reference classes: the oracle is duck-typed."""

    @staticmethod
    def _obj(**kw):
        return types.SimpleNamespace(**kw)

    class Variant:
        def __init__(self, start, end, id, allele):
            self.start, self.end, self.id, self.allele = start, end, id, allele

    class Haplotype:
        def __init__(self, chrom, start, end, id, ancestry=None):
            self.chrom, self.start, self.end, self.id, self.ancestry = chrom, start, end, id, ancestry
            self.variants = ()

    HaplotypeAncestry = Haplotype

    class Haplotypes:
        def __init__(self, fname=None):
            self.data = None

    HaplotypesAncestry = Haplotypes

    class GenotypesVCF:
        def __init__(self, fname=None):
            self.data = self.samples = self.variants = None
            self.ancestry = None
            self.ancestry_labels = {}

    GenotypesAncestry = GenotypesVCF


def run_oracle(case):
    haps, gts = _cases.build_objects(case, _NS)
    hap_list = list(haps.data.values())
    anc = case["kind"] == "ancestry"
    plural = (orc.haplotypes_ancestry_transform if anc else orc.haplotypes_transform)(hap_list, gts)
    one = orc.haplotype_ancestry_transform if anc else orc.haplotype_transform
    single = []
    for i, h in enumerate(hap_list):
        if len(h.variants) == 0:
            # the reference's single form raises on an empty haplotype (haplotypes.py:499-511);
            # the fixtures store the plural column for it
            with pytest.raises(ValueError):
                one(h, gts)
            single.append(plural[:, i])
        else:
            single.append(one(h, gts))
    return plural, np.array(single).reshape((len(hap_list), gts.data.shape[0], 2))


def test_golden_present():
    assert len(GOLDEN) >= 20


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: p.stem)
def test_oracle_matches_golden(path):
    case = _cases.load_case(path)
    plural, single = run_oracle(case)
    assert plural.dtype == np.bool_ and plural.shape == case["expected_plural"].shape
    np.testing.assert_array_equal(plural, case["expected_plural"])
    if case["kind"] == "ancestry" and case["gt"].shape[2] != 2:
        return
    np.testing.assert_array_equal(single, case["expected_single"])


def test_reference_literals():
    """The literals written in the reference's tests (tests/test_data.py:1616-1695,
    tests/test_transform.py:167-212)."""
    g1 = _cases.load_case(_cases.GOLDEN_DIR / "g1_plain_refalt.npz")
    np.testing.assert_array_equal(
        g1["expected_single"][0], np.array([[0, 1], [0, 1], [1, 1], [1, 1], [0, 0]], dtype=bool)
    )
    g2 = _cases.load_case(_cases.GOLDEN_DIR / "g2_plain_multiallelic.npz")
    np.testing.assert_array_equal(
        g2["expected_single"][0], np.array([[0, 0], [0, 0], [1, 0], [0, 0], [0, 1]], dtype=bool)
    )
    g3 = _cases.load_case(_cases.GOLDEN_DIR / "g3_plain_refalt_tweaked.npz")
    want = np.array(
        [
            [[0, 1], [0, 0], [0, 0]],
            [[0, 1], [0, 0], [1, 0]],
            [[1, 0], [0, 1], [0, 0]],
            [[1, 1], [0, 0], [0, 0]],
            [[0, 0], [0, 1], [1, 0]],
        ],
        dtype=bool,
    )
    np.testing.assert_array_equal(g3["expected_plural"], want)
    g5 = _cases.load_case(_cases.GOLDEN_DIR / "g5_ancestry.npz")
    np.testing.assert_array_equal(
        g5["expected_single"][0], np.array([[0, 0], [0, 0], [1, 1], [0, 0], [0, 0]], dtype=bool)
    )


def test_csr_form_equals_object_form():
    from haptools_b200.plan import plan_from_haplotypes

    for path in GOLDEN:
        case = _cases.load_case(path)
        if case["kind"] == "ancestry":
            continue
        haps, gts = _cases.build_objects(case, _NS)
        gts.index = lambda samples=True, variants=True, g=gts: setattr(
            g, "_var_idx", {v: i for i, v in enumerate(g.variants["id"])}
        )
        plan = plan_from_haplotypes(list(haps.data.values()), gts)
        out = orc.transform_from_csr(
            case["gt"], plan.key_col().astype(np.int64), plan.key_allele, plan.hap_off, plan.hap_key
        )
        np.testing.assert_array_equal(out, case["expected_plural"])


@pytest.mark.skipif(not ref_harness.available(), reason="reference checkout not mounted here")
@pytest.mark.parametrize("seed", range(6))
def test_oracle_matches_live_reference(seed):
    ns = ref_harness.load()
    rng = np.random.default_rng(1000 + seed)
    kind = "ancestry" if seed % 2 else "plain"
    case = _cases.random_case(
        rng, n=int(rng.integers(1, 200)), p=int(rng.integers(2, 40)), n_haps=int(rng.integers(1, 40)),
        kind=kind, max_alleles=2 if kind == "ancestry" else 4, missing_rate=0.05,
        empty_hap=bool(seed & 2), layout="F" if seed & 4 else "C",
    )
    haps, gts = _cases.build_objects(case, ns)
    want = haps.transform(gts).data
    case["expected_plural"] = want
    plural, single = run_oracle(case)
    np.testing.assert_array_equal(plural, want)
    for i, h in enumerate(haps.data.values()):
        if len(h.variants) == 0:
            with pytest.raises(ValueError):
                h.transform(gts)
        else:
            np.testing.assert_array_equal(single[i], h.transform(gts))
