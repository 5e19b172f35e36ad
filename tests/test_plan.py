"""Host planner: mirrors the reference's key dict / index arrays and its exceptions
(haptools/data/haplotypes.py:1375-1401, haptools/transform.py:103-135)."""
import types

import numpy as np
import pytest

import _cases
from haptools_b200 import data, synth
from haptools_b200 import transform as htransform
from haptools_b200.plan import plan_from_haplotypes, plan_from_refs
from oracle import transform_oracle as orc

NS = types.SimpleNamespace(
    Variant=data.Variant, Haplotype=data.Haplotype, Haplotypes=data.Haplotypes,
    GenotypesVCF=data.GenotypesVCF, HaplotypeAncestry=htransform.HaplotypeAncestry,
    HaplotypesAncestry=htransform.HaplotypesAncestry, GenotypesAncestry=htransform.GenotypesAncestry,
)


def _objects(name):
    case = _cases.load_case(_cases.GOLDEN_DIR / name)
    return case, *_cases.build_objects(case, NS)


def _check_plan(plan):
    assert np.all(np.diff(plan.var_col.astype(np.int64)) > 0)
    assert plan.var_key_off[0] == 0 and plan.var_key_off[-1] == plan.n_keys
    assert np.all(np.diff(plan.var_key_off.astype(np.int64)) >= 1)
    assert plan.hap_off[0] == 0 and plan.hap_off[-1] == plan.n_refs
    assert plan.n_refs == 0 or plan.hap_key.max() < plan.n_keys
    cv = plan.col_var()
    assert np.array_equal(np.nonzero(cv >= 0)[0], plan.var_col)


@pytest.mark.parametrize("path", _cases.list_golden(), ids=lambda p: p.stem)
def test_plan_matches_reference_key_dict(path):
    case, haps, gts = _objects(path.name)
    hap_list = list(haps.data.values())
    anc = case["kind"] == "ancestry"
    plan = plan_from_haplotypes(hap_list, gts, ancestry=anc)
    _check_plan(plan)
    alleles, idxs = orc.build_keys(hap_list)
    assert plan.n_haps == len(hap_list)
    assert [int(x) for x in np.diff(plan.hap_off.astype(np.int64))] == [len(i) for i in idxs]
    # same (column, allele) per haplotype-variant reference, whatever the key numbering
    var_idx = {v: i for i, v in enumerate(gts.variants["id"])}
    key_col = plan.key_col()
    flat = plan.hap_key
    pos = 0
    for hap in hap_list:
        for v in hap.variants:
            k = flat[pos]
            assert key_col[k] == var_idx[v.id]
            want = int(v.allele != gts.variants["alleles"][var_idx[v.id]][0]) if anc else \
                list(gts.variants["alleles"][var_idx[v.id]]).index(v.allele)
            if gts.data.dtype == np.bool_:
                want = min(want, 1)
            assert plan.key_allele[k] == want
            if anc:
                assert plan.key_label[k] == gts.ancestry_labels[hap.ancestry]
            pos += 1
    if not anc:
        assert plan.n_keys == len(alleles)  # one key per distinct (variant, allele)


def test_missing_allele_raises_like_the_reference():
    case, haps, gts = _objects("g1_plain_refalt.npz")
    hap = list(haps.data.values())[0]
    hap.variants = (data.Variant(start=10114, end=10115, id="1:10114:T:C", allele="G"),)
    with pytest.raises(ValueError, match="Some alleles were not present in the genotypes"):
        plan_from_haplotypes(list(haps.data.values()), gts)


def test_absent_variant_raises():
    case, haps, gts = _objects("g1_plain_refalt.npz")
    hap = list(haps.data.values())[0]
    hap.variants = hap.variants + (data.Variant(start=5, end=6, id="1:5:A:C", allele="A"),)
    with pytest.raises(ValueError, match="absent in the provided genotypes"):
        plan_from_haplotypes([hap], gts, who=f"haplotype '{hap.id}'")


def test_duplicate_variant_ids_raise():
    case, haps, gts = _objects("g1_plain_refalt.npz")
    gts.variants["id"][1] = gts.variants["id"][0]
    gts._var_idx = None
    with pytest.raises(ValueError, match="Found duplicate variant IDs"):
        plan_from_haplotypes(list(haps.data.values()), gts)


def test_unknown_population():
    case, haps, gts = _objects("g5_ancestry.npz")
    hap_list = list(haps.data.values())
    hap_list[0].ancestry = "NOPE"
    # plural form: uint8 <- -1 (haptools/transform.py:115) overflows on numpy >= 2
    with pytest.raises(OverflowError):
        plan_from_haplotypes(hap_list, gts, ancestry=True)
    plan, never = plan_from_haplotypes(hap_list[:1], gts, ancestry=True, unknown_label="never")
    assert never.tolist() == [True]


def test_type_ids_and_repeats_are_skipped():
    case, haps, gts = _objects("g1_plain_refalt.npz")
    haps.data["STR_1"] = data.Repeat(chrom="1", start=1, end=5, id="STR_1")
    haps.index(force=True)
    assert haps.type_ids["R"] == ["STR_1"] and len(haps.type_ids["H"]) == 3
    haps.data["H9"] = data.Haplotype(chrom="1", start=1, end=2, id="H9")
    haps.index()  # a no-op once built (haptools/data/haplotypes.py:983-985)
    assert "H9" not in haps.type_ids["H"]


def test_biallelic_tables():
    # column 2: alleles 0 and 1; column 5: allele 1 only; column 6: allele 2 (-> var_other)
    plan = plan_from_refs(np.array([2, 2, 5, 6, 6]), np.array([0, 1, 1, 2, 0]), np.array([0, 2, 5]), 8)
    ck, other = plan.biallelic_tables()
    assert ck.shape == (8, 2) and ck.dtype == np.int32
    assert ck[2].tolist() == [0, 1] and ck[5].tolist() == [-1, 2] and ck[6].tolist() == [-1, -1]
    assert (ck[[0, 1, 3, 4, 7]] == -1).all()
    assert other.tolist() == [2] and plan.var_col[other[0]] == 6
    anc = plan_from_refs(np.array([2, 2, 5, 6]), np.array([1, 0, 1, 1]), np.array([0, 1, 2, 3, 4]), 8, np.array([3, 1, 0, 9]))
    k01, rng, other = anc.fast_tables()
    assert k01 is None and rng.shape == (8, 2)
    assert rng[2].tolist() == [0, 2] and rng[5].tolist() == [2, 1] and rng[6].tolist() == [0, 0]  # label 9 -> other
    assert other.tolist() == [2]
    ck, other = anc.biallelic_tables()
    assert (ck == -1).all() and len(other) == anc.n_vars


def test_plan_from_refs_dedups_and_validates():
    plan = plan_from_refs(np.array([5, 2, 5, 2]), np.array([1, 0, 1, 1]), np.array([0, 2, 4]), 8)
    _check_plan(plan)
    assert plan.n_keys == 3 and plan.n_vars == 2
    assert plan.var_col.tolist() == [2, 5]
    assert plan.hap_key.tolist() == [2, 0, 2, 1]
    with pytest.raises(ValueError):
        plan_from_refs(np.array([9]), np.array([0]), np.array([0, 1]), 8)
    with pytest.raises(ValueError):
        plan_from_refs(np.array([1]), np.array([0]), np.array([0, 2]), 8)


def test_synth_numpy_properties():
    a = synth.synth_gt(50, 200, seed=3, mode=1)
    b = synth.synth_gt(20, 200, seed=3, mode=1, sample_offset=30)
    assert np.array_equal(a[30:], b)  # a pure function of the global sample index
    c = synth.synth_gt(50, 200, seed=3, mode=1, variants=np.arange(100, 200))
    assert np.array_equal(a[:, 100:], c)
    assert set(np.unique(a)) <= {0, 1}
    m = synth.synth_gt(300, 300, seed=3, mode=0, missing_per_64k=2000)
    assert 0.01 < (m == 255).mean() < 0.06
    anc = synth.synth_ancestry(10, 5000, seed=1, n_pops=5, tract_len=500)
    assert anc.max() < 5 and (np.diff(anc[:, :, 0].astype(int), axis=1) != 0).mean() < 0.01
    plan = synth.synth_plan(5000, 300, seed=1)
    _check_plan(plan)
    k = np.diff(plan.hap_off.astype(np.int64))
    assert k.min() >= 2 and k.max() <= 20


def test_quad_tables_restate_the_lists():
    """plan.quad_tables: padded, flagged quads of record byte offsets, one offset per 16 haplotypes."""
    import numpy as np
    from haptools_b200.plan import plan_from_refs

    rng = np.random.default_rng(4)
    H, p = 70, 40
    k = rng.integers(0, 11, size=H)
    k[3] = 0
    off = np.concatenate([[0], np.cumsum(k)])
    cols = np.concatenate([np.sort(rng.choice(p, size=int(x), replace=False)) for x in k] + [np.zeros(0, np.int64)])
    plan = plan_from_refs(cols, rng.integers(0, 2, size=len(cols)), off, p)
    q16, quads = plan.quad_tables()
    assert q16.dtype == np.uint32 and quads.dtype == np.uint32 and quads.shape[1] == 4
    assert len(q16) == (H + 15) // 16 + 1 and q16[0] == 0 and q16[-1] == len(quads)
    qi = 0
    for h in range(H):
        keys = plan.hap_key[plan.hap_off[h]:plan.hap_off[h + 1]].astype(np.int64)
        nq = max(1, (len(keys) + 3) // 4)
        if h % 16 == 0:
            assert q16[h // 16] == qi
        flat = quads[qi:qi + nq].astype(np.int64)
        flags = flat[:, 0] & 0xFF
        assert np.all(flat[:, 1:] & 0xFF == 0)
        assert list(flags & 1) == [0] * (nq - 1) + [1]
        assert np.all((flags & 2) == (2 if len(keys) == 0 else 0))
        got = (flat >> 8).reshape(-1)
        if len(keys):
            np.testing.assert_array_equal(got[:len(keys)], keys)
            assert np.all(got[len(keys):] == keys[-1])
        else:
            assert np.all(got == 0)
        qi += nq
    assert qi == len(quads)
    empty = plan_from_refs(np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(1, np.int64), p)
    assert empty.quad_tables() == (None, None)


# ---------------------------------------------------------------- vectorised planner (N4)
def _same_plan(a, b):
    for f in ("n_cols", "var_col", "var_key_off", "key_allele", "key_label", "hap_off", "hap_key"):
        x, y = getattr(a, f), getattr(b, f)
        assert (x is None) == (y is None), f
        if x is not None:
            np.testing.assert_array_equal(x, y, err_msg=f)


@pytest.mark.parametrize("path", _cases.list_golden(), ids=lambda p: p.stem)
def test_vectorised_planner_equals_reference_loop(path):
    from haptools_b200.plan import plan_from_haplotypes_loop

    case, haps, gts = _objects(path.name)
    hap_list = list(haps.data.values())
    anc = case["kind"] == "ancestry"
    _same_plan(plan_from_haplotypes(hap_list, gts, ancestry=anc), plan_from_haplotypes_loop(hap_list, gts, ancestry=anc))
    if anc:
        (pa, na), (pb, nb) = (f(hap_list, gts, ancestry=True, unknown_label="never") for f in (plan_from_haplotypes, plan_from_haplotypes_loop))
        _same_plan(pa, pb)
        np.testing.assert_array_equal(na, nb)


def _random_objects(rng, p, H, multi, anc_labels=None, bool_gt=False):
    gts = (htransform.GenotypesAncestry if anc_labels else data.GenotypesVCF)(None)
    alleles = [tuple(rng.permutation(["A", "C", "G", "T"])[: (rng.integers(2, 5) if multi else 2)]) for _ in range(p)]
    gts.variants = np.array([(f"v{i}", "1", i + 1, alleles[i]) for i in range(p)], dtype=gts.variants.dtype)
    gts.samples = ("s0", "s1")
    gts.data = np.zeros((2, p, 2), dtype=np.bool_ if bool_gt else np.uint8)
    if anc_labels:
        gts.ancestry_labels = anc_labels
    haps = []
    for h in range(H):
        k = int(rng.integers(0, 7))
        cols = np.sort(rng.choice(p, size=k, replace=False))
        hp = htransform.HaplotypeAncestry("1", 1, 2, f"H{h}", rng.choice(list(anc_labels))) if anc_labels else data.Haplotype("1", 1, 2, f"H{h}")
        hp.variants = tuple(data.Variant(int(c) + 1, int(c) + 2, f"v{c}", str(rng.choice(alleles[c]))) for c in cols)
        haps.append(hp)
    return haps, gts


@pytest.mark.parametrize("multi,anc,bool_gt", [(False, False, False), (True, False, False), (True, False, True), (True, True, False)])
def test_vectorised_planner_random(multi, anc, bool_gt):
    from haptools_b200.plan import plan_from_haplotypes_loop

    rng = np.random.default_rng(17)
    labels = {"YRI": 0, "CEU": 1, "PEL": 2} if anc else None
    haps, gts = _random_objects(rng, 60, 200, multi, labels, bool_gt)
    _same_plan(plan_from_haplotypes(haps, gts, ancestry=anc), plan_from_haplotypes_loop(haps, gts, ancestry=anc))


def test_vectorised_planner_errors_match_the_loop():
    from haptools_b200.plan import plan_from_haplotypes_loop

    rng = np.random.default_rng(3)
    for fn in (plan_from_haplotypes, plan_from_haplotypes_loop):
        haps, gts = _random_objects(rng, 30, 40, True)
        bad = data.Haplotype("1", 1, 2, "bad")
        bad.variants = (data.Variant(1, 2, "v3", "A"), data.Variant(1, 2, "nope", "A"), data.Variant(1, 2, "nada", "C"))
        with pytest.raises(ValueError, match="absent in the provided genotypes") as e:
            fn(haps + [bad], gts)
        assert "nope" in str(e.value) and "nada" in str(e.value)
        wrong = data.Haplotype("1", 1, 2, "wrong")
        wrong.variants = (data.Variant(1, 2, "v3", "N"),)
        with pytest.raises(ValueError, match="Some alleles were not present in the genotypes"):
            fn(haps + [wrong], gts)
        gts2 = data.GenotypesVCF(None)
        gts2.variants = np.array([("d", "1", 1, ("A", "C")), ("d", "1", 2, ("A", "C"))], dtype=gts2.variants.dtype)
        gts2.data = np.zeros((1, 2, 2), dtype=np.uint8)
        with pytest.raises(ValueError, match="duplicate variant IDs"):
            fn(haps, gts2)
        haps_a, gts_a = _random_objects(rng, 30, 20, False, {"YRI": 0, "CEU": 1})
        haps_a[4].ancestry = "XXX"
        with pytest.raises(OverflowError):
            fn(haps_a, gts_a, ancestry=True)
        plan, never = fn(haps_a, gts_a, ancestry=True, unknown_label="never")
        assert never[4] == (len(haps_a[4].variants) > 0) and never.sum() <= 1


def test_transform_result_variants_layout():
    """hap_gts.variants built column-wise equals the reference's tuple-per-haplotype array
    (haptools/data/haplotypes.py:1369-1372)."""
    case, haps, gts = _objects("file_example_examplehap.npz")
    hap_list = list(haps.data.values())
    want = np.array([(h.id, h.chrom, h.start, ("A", "T")) for h in hap_list], dtype=data.GenotypesVCF(None).variants.dtype)
    got = np.empty(len(hap_list), dtype=want.dtype)
    got["id"] = [h.id for h in hap_list]
    got["chrom"] = [h.chrom for h in hap_list]
    got["pos"] = [h.start for h in hap_list]
    got["alleles"].fill(("A", "T"))
    assert (got == want).all() and all(type(a) is tuple for a in got["alleles"])
