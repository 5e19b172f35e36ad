"""VCF output of the transform (N4): the text oracle against the reference's own expected CLI
output (CPU) and the device-formatted writer against both (GPU)."""
import io
import json
import types

import numpy as np
import pytest

import _cases
from haptools_b200 import data
from haptools_b200 import transform as htransform
from oracle import vcf_text_oracle as vto

NS = types.SimpleNamespace(
    Variant=data.Variant, Haplotype=data.Haplotype, Haplotypes=data.Haplotypes,
    GenotypesVCF=data.GenotypesVCF, HaplotypeAncestry=htransform.HaplotypeAncestry,
    HaplotypesAncestry=htransform.HaplotypesAncestry, GenotypesAncestry=htransform.GenotypesAncestry,
)
GOLDEN_TEXT = json.loads((_cases.GOLDEN_DIR / "vcftext.json").read_text())


def _case_for(name):
    g = GOLDEN_TEXT[name]
    case = _cases.load_case(_cases.GOLDEN_DIR / g["fixture"])
    keep = list(range(len(case["haps"])))
    if name == "test_basic_subset":
        # transform_haps drops the haplotypes whose variants are not all in the genotypes
        # (haptools/transform.py:630-641); the test removed the last two variants
        present = set(case["var_id"][:-2])
        keep = [i for i, h in enumerate(case["haps"]) if {v["id"] for v in h["variants"]} <= present]
    return g, case, keep


@pytest.mark.parametrize("name", sorted(GOLDEN_TEXT))
def test_text_oracle_matches_reference_cli_output(name):
    g, case, keep = _case_for(name)
    mat = case["expected_plural"][:, keep]
    if g["maf"] is not None:  # check_maf(threshold, discard_also=True), genotypes.py:598-613
        af = mat.astype(bool).sum(axis=(0, 2)) / (2 * mat.shape[0])
        sel = np.nonzero(~(np.minimum(af, 1 - af) < g["maf"]))[0]
        mat, keep = mat[:, sel], [keep[i] for i in sel]
    variants = [(case["haps"][i]["id"], case["haps"][i]["chrom"], case["haps"][i]["start"], ("A", "T")) for i in keep]
    assert vto.vcf_text(mat, case["samples"], variants) == g["text"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(GOLDEN_TEXT))
def test_writer_reproduces_reference_cli_output(name):
    from haptools_b200 import writers

    g, case, keep = _case_for(name)
    haps, gts = _cases.build_objects(case, NS)
    ids = [case["haps"][i]["id"] for i in keep]
    haps.data = {k: v for k, v in haps.data.items() if k in ids}
    buf = io.BytesIO()
    n_rec = writers.transform_to_vcf(haps, gts, buf, maf=g["maf"], ancestry=case["kind"] == "ancestry")
    assert buf.getvalue().decode() == g["text"]
    assert n_rec == len(g["text"].splitlines()) - 5


@pytest.mark.gpu
@pytest.mark.parametrize("n,H", [(1, 3), (5, 1), (127, 20), (128, 9), (1024, 12), (1500, 70), (4100, 33)])
def test_writer_random_vs_text_oracle(n, H):
    import torch

    from haptools_b200 import dist as hdist
    from haptools_b200 import writers

    rng = np.random.default_rng(n * 13 + H)
    mat = rng.random((n, H, 2)) < 0.4
    bits = torch.from_numpy(hdist.pack_result_bits(mat).view(np.int32)).cuda()
    samples = [f"S{i}" for i in range(n)]
    variants = np.empty(H, dtype=data.GenotypesVCF(None).variants.dtype)
    variants["id"] = [f"H{i}" for i in range(H)]
    variants["chrom"] = [str(1 + (i * 3) // max(H, 1)) for i in range(H)]
    variants["pos"] = np.arange(H) * 7 + 100
    variants["alleles"].fill(("A", "T"))
    as_tuples = lambda idx: [(variants["id"][i], variants["chrom"][i], variants["pos"][i], ("A", "T")) for i in idx]
    for keep, chunk in ((None, None), (None, 4), (np.array(sorted(rng.choice(H, size=max(1, H // 2), replace=False))), 3)):
        buf = io.BytesIO()
        writers.write_vcf_from_bits(buf, variants, samples, bits, n, keep=keep, chunk_haps=chunk)
        idx = list(range(H)) if keep is None else list(keep)
        assert buf.getvalue().decode() == vto.vcf_text(mat[:, idx], samples, as_tuples(idx))
    buf = io.BytesIO()
    assert writers.write_vcf_from_bits(buf, variants, samples, bits, n, keep=np.zeros(0, dtype=np.int64)) == 0
    assert buf.getvalue().decode().count("\n") == 4  # header without contig lines
