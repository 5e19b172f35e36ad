#!/usr/bin/env python
"""
Generates the golden vectors in this directory by running the UNMODIFIED reference
(CAST-genomics/haptools at /root/reference, imported through oracle/ref_harness.py).

Run it in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Outputs (all small, committed):
  g1_g4_plain_*.npz, g5_g6_ancestry_*.npz -- the reference's own unit-test fixtures for
        this path (tests/test_data.py:1616-1695, tests/test_transform.py:167-212); the
        expected arrays stored are what the reference RETURNED, and the script asserts
        they equal the literals written in those tests.
  file_*.npz  -- the reference's tests/data fixtures (simple.vcf/.hap, simple-ancestry.vcf,
        example.vcf.gz x example.hap.gz / basic.hap, apoe.vcf.gz x apoe4.hap) parsed with
        the small text readers below (cyvcf2/pysam are not installed) and transformed by
        the reference.
  rand_*.npz  -- seeded random cases: multi-allelic, missing (254/255), bool dtype,
        phase column, empty haplotype, shared keys, variant-major layout, ancestry.
"""
from __future__ import annotations

import gzip
import importlib.util
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REPO = HERE.parent.parent
sys.path.insert(0, str(REPO))

from oracle import ref_harness  # noqa: E402

ns = ref_harness.load()  # puts /root/reference first on sys.path

spec = importlib.util.spec_from_file_location("_cases", REPO / "tests" / "_cases.py")
_cases = importlib.util.module_from_spec(spec)
spec.loader.exec_module(_cases)

from tests.test_data import TestHaplotypes, TestGenotypesVCF  # noqa: E402 (reference's)
from tests.test_transform import (  # noqa: E402 (reference's)
    TestHaplotypesAncestry,
    TestGenotypesAncestry,
)

DATADIR = ref_harness.REFERENCE_ROOT / "tests" / "data"


# ------------------------------------------------------------------ helpers
def case_from_objects(kind, haps, gts, source):
    """Turn reference objects into a portable case and run the reference on them."""
    haps.index(force=True)
    hap_list = [haps.data[h] for h in haps.type_ids["H"]]
    case = dict(
        kind=kind,
        gt=np.array(gts.data),  # preserves strides? no: made C-contiguous on save anyway
        samples=list(gts.samples),
        var_id=[str(v) for v in gts.variants["id"]],
        var_chrom=[str(v) for v in gts.variants["chrom"]],
        var_pos=[int(v) for v in gts.variants["pos"]],
        var_alleles=[list(a) for a in gts.variants["alleles"]],
        haps=[
            dict(
                id=h.id,
                chrom=h.chrom,
                start=int(h.start),
                end=int(h.end),
                **({"ancestry": h.ancestry} if kind == "ancestry" else {}),
                variants=[
                    dict(start=int(v.start), end=int(v.end), id=v.id, allele=v.allele)
                    for v in h.variants
                ],
            )
            for h in hap_list
        ],
        source=source,
    )
    if kind == "ancestry":
        case["ancestry"] = np.array(gts.ancestry)
        case["ancestry_labels"] = dict(gts.ancestry_labels)
    out = haps.transform(gts)
    case["expected_plural"] = np.array(out.data)
    assert case["expected_plural"].dtype == np.bool_
    case["expected_single"] = np.array([h.transform(gts) for h in hap_list]).reshape(
        (len(hap_list), len(gts.samples), 2)
    )
    # the two reference entry points must agree with each other
    np.testing.assert_array_equal(
        case["expected_plural"], case["expected_single"].transpose((1, 0, 2))
    )
    return case


def fill_expected(case):
    """Run the reference on a portable case (built by _cases.random_case)."""
    haps, gts = _cases.build_objects(case, ns)
    haps.index(force=True)
    hap_list = [haps.data[h] for h in haps.type_ids["H"]]
    case["expected_plural"] = np.array(haps.transform(gts).data)
    singles = []
    for i, h in enumerate(hap_list):
        if len(h.variants) == 0:
            # the reference's single-haplotype form cannot broadcast a (1, 0) allele
            # array (haplotypes.py:499-511) and raises ValueError; the plural form
            # yields all-True (np.all over an empty axis, haplotypes.py:1412).  Store the
            # plural column; tests assert that the single form raises.
            try:
                h.transform(gts)
                raise AssertionError("expected the reference to raise on an empty haplotype")
            except ValueError:
                pass
            singles.append(case["expected_plural"][:, i])
        else:
            singles.append(h.transform(gts))
    case["expected_single"] = np.array(singles).reshape(
        (len(hap_list), case["gt"].shape[0], 2)
    )
    return case


def read_vcf_text(path, want_pop=False):
    """A minimal phased-GT VCF text reader (test-only; no cyvcf2 here)."""
    opener = gzip.open if str(path).endswith(".gz") else open
    samples, rows, pops = None, [], []
    with opener(path, "rt") as f:
        for line in f:
            if line.startswith("##"):
                continue
            fields = line.rstrip("\n").split("\t")
            if line.startswith("#CHROM"):
                samples = fields[9:]
                continue
            chrom, pos, vid, ref, alt = fields[0], int(fields[1]), fields[2], fields[3], fields[4]
            fmt = fields[8].split(":")
            gt_i = fmt.index("GT")
            gts, pop = [], []
            for call in fields[9:]:
                parts = call.split(":")
                a, b = parts[gt_i].split("|")
                gts.append((255 if a == "." else int(a), 255 if b == "." else int(b)))
                if want_pop:
                    pop.append(tuple(parts[fmt.index("POP")].split(",")))
            rows.append((chrom, pos, vid, (ref,) + tuple(alt.split(",")), gts))
            pops.append(pop)
    # variant-major physical layout, exposed as the transposed view -- exactly what the
    # reference's reader builds (haptools/data/genotypes.py:183-186, 210)
    data = np.array([r[4] for r in rows], dtype=np.uint8).transpose((1, 0, 2))
    return samples, rows, data, pops


def read_hap_text(path, ancestry=False):
    opener = gzip.open if str(path).endswith(".gz") else open
    haps, order = {}, []
    with opener(path, "rt") as f:
        for line in f:
            fields = line.rstrip("\n").split("\t")
            if fields[0] == "H":
                kwargs = dict(chrom=fields[1], start=int(fields[2]), end=int(fields[3]), id=fields[4])
                if ancestry:
                    haps[fields[4]] = ns.HaplotypeAncestry(ancestry=fields[5], **kwargs)
                else:
                    haps[fields[4]] = ns.Haplotype(**kwargs)
                order.append(fields[4])
                haps[fields[4]].variants = tuple()
            elif fields[0] == "V":
                haps[fields[1]].variants += (
                    ns.Variant(start=int(fields[2]), end=int(fields[3]), id=fields[4], allele=fields[5]),
                )
    cls = ns.HaplotypesAncestry if ancestry else ns.Haplotypes
    hp = cls(fname=None)
    hp.data = {k: haps[k] for k in order}
    return hp


def gts_from_vcf(path, ancestry_labels=None):
    samples, rows, data, pops = read_vcf_text(path, want_pop=ancestry_labels is not None)
    gts = (ns.GenotypesAncestry if ancestry_labels is not None else ns.GenotypesVCF)(fname=None)
    gts.samples = tuple(samples)
    gts.data = data
    gts.variants = np.array(
        [(r[2], r[0], r[1], r[3]) for r in rows],
        dtype=[("id", "U50"), ("chrom", "U10"), ("pos", np.uint32), ("alleles", object)],
    )
    if ancestry_labels is not None:
        gts.ancestry_labels = dict(ancestry_labels)
        gts.ancestry = np.array(
            [[[ancestry_labels[x] for x in call] for call in pop] for pop in pops],
            dtype=np.uint8,
        ).transpose((1, 0, 2))
    return gts


# ------------------------------------------------------------------ main
def main():
    for old in HERE.glob("*.npz"):
        old.unlink()
    n_written = 0

    def emit(name, case):
        nonlocal n_written
        _cases.save_case(HERE / f"{name}.npz", case)
        n_written += 1
        print(f"{name}: gt{case['gt'].shape} haps={len(case['haps'])} ones={int(case['expected_plural'].sum())}")

    # ---- G1/G3: the reference's dummy haplotypes on its refalt fixture
    th, tg = TestHaplotypes(), TestGenotypesVCF()
    haps, gens = th._get_dummy_haps(), tg._get_fake_genotypes_refalt()
    case = case_from_objects("plain", haps, gens, "tests/test_data.py:1616-1631 (G1)")
    np.testing.assert_array_equal(
        case["expected_single"][0], np.array([[0, 1], [0, 1], [1, 1], [1, 1], [0, 0]])
    )
    emit("g1_plain_refalt", case)

    gens.data[[2, 4], 0, 1] = 1
    gens.data[[1, 4], 2, 0] = 1
    case = case_from_objects("plain", haps, gens, "tests/test_data.py:1650-1668 (G3)")
    np.testing.assert_array_equal(
        case["expected_plural"],
        np.array(
            [
                [[0, 1], [0, 0], [0, 0]],
                [[0, 1], [0, 0], [1, 0]],
                [[1, 0], [0, 1], [0, 0]],
                [[1, 1], [0, 0], [0, 0]],
                [[0, 0], [0, 1], [1, 0]],
            ]
        ),
    )
    emit("g3_plain_refalt_tweaked", case)

    # ---- G2/G4: multi-allelic
    haps, gens = th._get_dummy_haps_multiallelic(), tg._get_fake_genotypes_multiallelic()
    # H3's second variant (1:10122:A:G, allele C) is fine for the plural form
    hap1 = list(haps.data.values())[0]
    np.testing.assert_array_equal(
        hap1.transform(gens), np.array([[0, 0], [0, 0], [1, 0], [0, 0], [0, 1]])
    )
    case = case_from_objects("plain", haps, gens, "tests/test_data.py:1633-1648 (G2)")
    emit("g2_plain_multiallelic", case)
    gens.data[[2, 4], 0, 1] = 1
    gens.data[[1, 4], 2, 0] = 1
    gens.data[[1, 4], 3, 0] = 2
    case = case_from_objects("plain", haps, gens, "tests/test_data.py:1673-1692 (G4)")
    np.testing.assert_array_equal(
        case["expected_plural"],
        np.array(
            [
                [[0, 0], [0, 0], [0, 0]],
                [[0, 0], [0, 0], [1, 0]],
                [[1, 0], [0, 1], [0, 0]],
                [[0, 0], [0, 0], [0, 0]],
                [[0, 0], [0, 1], [1, 0]],
            ]
        ),
    )
    emit("g4_plain_multiallelic_tweaked", case)

    # ---- with the phase column still attached (plain path slices [:, :, :2])
    gens = tg._get_fake_genotypes_refalt(with_phase=True)
    case = case_from_objects("plain", th._get_dummy_haps(), gens, "phase column kept (haplotypes.py:1404)")
    emit("g1_plain_refalt_phasecol", case)

    # ---- G5/G6: ancestry
    tha, tga = TestHaplotypesAncestry(), TestGenotypesAncestry()
    haps, gens = tha._get_dummy_haps(), tga._get_fake_genotypes()
    gens.check_phase()
    case = case_from_objects("ancestry", haps, gens, "tests/test_transform.py:167-199 (G5, G6 stage 1)")
    np.testing.assert_array_equal(
        case["expected_single"][0], np.array([[0, 0], [0, 0], [1, 1], [0, 0], [0, 0]])
    )
    emit("g5_ancestry", case)
    gens.data[[2, 4], 0, 1] = 1
    gens.data[[1, 4], 2, 0] = 1
    case = case_from_objects("ancestry", haps, gens, "tests/test_transform.py:201-212 (G6 stage 2)")
    exp = np.zeros((5, 3, 2), dtype=bool)
    exp[2, 0, 0] = 1
    exp[[2, 4], 1, 1] = 1
    np.testing.assert_array_equal(case["expected_plural"], exp)
    emit("g6_ancestry_tweaked", case)

    # ---- file fixtures (G7-G9 at array level)
    gts = gts_from_vcf(DATADIR / "simple.vcf")
    case = case_from_objects("plain", read_hap_text(DATADIR / "simple.hap"), gts, "tests/data/simple.vcf x simple.hap (G7)")
    np.testing.assert_array_equal(case["expected_plural"][:, 0], [[0, 1], [0, 1], [1, 1], [1, 1], [0, 0]])
    assert not case["expected_plural"][:, 1:].any()
    emit("file_simple", case)

    labels = {"YRI": 0, "CEU": 1, "ASW": 2}
    gts = gts_from_vcf(DATADIR / "simple-ancestry.vcf", labels)
    case = case_from_objects(
        "ancestry", read_hap_text(DATADIR / "simple.hap", ancestry=True), gts,
        "tests/data/simple-ancestry.vcf x simple.hap --ancestry (G7)",
    )
    np.testing.assert_array_equal(case["expected_plural"][:, 0], [[0, 0], [0, 0], [1, 1], [0, 0], [0, 0]])
    emit("file_simple_ancestry", case)

    gts = gts_from_vcf(DATADIR / "example.vcf.gz")
    emit("file_example_examplehap", case_from_objects(
        "plain", read_hap_text(DATADIR / "example.hap.gz"), gts, "tests/data/example.vcf.gz x example.hap.gz"))
    emit("file_example_basichap", case_from_objects(
        "plain", read_hap_text(DATADIR / "basic.hap"), gts, "tests/data/example.vcf.gz x basic.hap (G9 inputs)"))

    gts = gts_from_vcf(DATADIR / "apoe.vcf.gz")
    case = case_from_objects("plain", read_hap_text(DATADIR / "apoe4.hap"), gts, "tests/data/apoe.vcf.gz x apoe4.hap (G8)")
    s = [case["samples"].index("HG00097"), case["samples"].index("NA12878")]
    np.testing.assert_array_equal(case["expected_plural"][s, 0], [[0, 1], [0, 0]])
    emit("file_apoe", case)

    # ---- seeded random cases
    rng = np.random.default_rng(20261019)
    specs = [
        ("rand_plain_small", dict(n=37, p=11, n_haps=9)),
        ("rand_plain_multiallelic_missing", dict(n=130, p=40, n_haps=33, max_alleles=4, missing_rate=0.05)),
        ("rand_plain_bool", dict(n=65, p=23, n_haps=17, as_bool=True)),
        ("rand_plain_phasecol_F", dict(n=257, p=31, n_haps=70, phase_column=True, layout="F", missing_rate=0.02)),
        ("rand_plain_emptyhap", dict(n=33, p=12, n_haps=7, empty_hap=True)),
        ("rand_plain_wide", dict(n=1100, p=64, n_haps=150, max_hap_len=20, max_alleles=3, layout="F")),
        ("rand_plain_manyhaps", dict(n=96, p=50, n_haps=300, max_hap_len=3)),
        ("rand_ancestry_small", dict(n=41, p=13, n_haps=10, kind="ancestry")),
        ("rand_ancestry_F", dict(n=300, p=37, n_haps=65, kind="ancestry", layout="F", n_pops=5, missing_rate=0.03)),
        ("rand_ancestry_emptyhap", dict(n=34, p=9, n_haps=5, kind="ancestry", empty_hap=True)),
    ]
    for name, kw in specs:
        case = _cases.random_case(rng, **kw)
        case["source"] = f"random_case({kw}) seed 20261019"
        emit(name, fill_expected(case))
    print(f"wrote {n_written} golden files")


if __name__ == "__main__":
    main()
