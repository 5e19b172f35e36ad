"""
TEST INFRASTRUCTURE ONLY -- never imported by the product path (haptools_b200/).

Plain-Python restatement of the VCF text `GenotypesVCF.write` produces through pysam/htslib
(haptools/data/genotypes.py:752-813) for a phased matrix without a phase column: header, then
one record per variant with "a|b" (or "." for a missing allele) per sample.  pysam is not
installed here, so the restatement is pinned by the literal expected strings of the reference's
own CLI tests (tests/test_transform.py:215-392; tests/golden/vcftext.json).
"""
from __future__ import annotations

import numpy as np


def vcf_text(data: np.ndarray, samples, variants) -> str:
    """data (n, p, 2) uint8/bool; variants: iterable of (id, chrom, pos, alleles)."""
    chroms = list(dict.fromkeys(str(v[1]) for v in variants))
    out = ["##fileformat=VCFv4.2", '##FILTER=<ID=PASS,Description="All filters passed">']
    out += [f"##contig=<ID={c}>" for c in chroms]
    out.append('##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">')
    out.append("\t".join(["#CHROM", "POS", "ID", "REF", "ALT", "QUAL", "FILTER", "INFO", "FORMAT"] + list(samples)))
    for j, (vid, chrom, pos, alleles) in enumerate(variants):
        alt = ",".join(alleles[1:]) if len(alleles) > 1 else "."
        gts = []
        for s in range(data.shape[0]):
            a, b = (int(x) for x in data[s, j, :2])
            gts.append("|".join("." if x == 255 else str(x) for x in (a, b)))
        out.append("\t".join([str(chrom), str(int(pos)), str(vid), alleles[0], alt, ".", ".", ".", "GT"] + gts))
    return "\n".join(out) + "\n"
