"""
Output of `haptools transform` without the host matrix: the tail of
`transform_haps` (haptools/transform.py:662-676) --

    hp.transform(gt, hp_gt);  hp_gt.check_maf(threshold=maf, discard_also=True);  hp_gt.write()

-- where `GenotypesVCF.write` (haptools/data/genotypes.py:752-813) is a Python loop over every
(haplotype, sample) pair through pysam.  Here the result stays bit-packed on the device
(`engine.transform_device(fmt="bits")`), the MAF filter uses the allele counts that fall out of
the transform (`ht_count_bits`; `check_maf`, genotypes.py:598-602) and the GT columns of every
VCF line are formatted by a kernel (`ht_format_vcf_gt`: "a|b\\t" per sample); the host only
writes the header and the nine fixed columns of each record.  Text output only (`.vcf`, a file
object or stdout): bgzip/BCF encoding is htslib's and out of scope.

The text is what htslib prints for the records `GenotypesVCF.write` builds; pinned by the
expected strings of the reference's own CLI tests (tests/test_transform.py:215-392) in
tests/golden/vcftext.json.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _cuda, engine
from ._cuda import check


def vcf_header(samples, chroms) -> str:
    """Header as htslib writes it for the one GenotypesVCF.write builds (genotypes.py:756-778):
    fileformat, the PASS filter every htslib header carries, one contig line per chromosome (in
    order of first appearance), the GT FORMAT line, the column line."""
    lines = ["##fileformat=VCFv4.2", '##FILTER=<ID=PASS,Description="All filters passed">']
    lines += [f"##contig=<ID={c}>" for c in dict.fromkeys(str(c) for c in chroms)]
    lines.append('##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">')
    cols = ["#CHROM", "POS", "ID", "REF", "ALT", "QUAL", "FILTER", "INFO", "FORMAT"] + [str(s) for s in samples]
    lines.append("\t".join(cols))
    return "\n".join(lines) + "\n"


def vcf_fixed_columns(variants: np.ndarray) -> list:
    """The nine fixed columns (with the trailing tab) of every record: CHROM POS ID REF ALT and
    '.' for QUAL / FILTER / INFO, FORMAT = GT (genotypes.py:782-794)."""
    out = []
    for v in variants:
        alleles = v["alleles"]
        alt = ",".join(alleles[1:]) if len(alleles) > 1 else "."
        out.append(f"{v['chrom']}\t{v['pos']}\t{v['id']}\t{alleles[0]}\t{alt}\t.\t.\t.\tGT\t".encode())
    return out


def maf_keep(counts, n_samples: int, threshold: float) -> np.ndarray:
    """Indices of the haplotypes `check_maf(threshold, discard_also=True)` keeps
    (genotypes.py:598-613), from their allele counts."""
    maf = engine.maf_from_counts(counts, n_samples)
    return np.nonzero(~(maf < threshold))[0]


def write_vcf_from_bits(fileobj, variants: np.ndarray, samples, bits: torch.Tensor, n_samples: int,
                        keep: np.ndarray = None, chunk_haps: int = None) -> int:
    """
    Write the VCF of a bit-packed transform result `(tiles, H, 32, 2)` to a binary file object.
    `variants` is the `(H,)` structured array of the result (`hap_gts.variants`), `keep` an
    optional ascending index array (the haplotypes that survived `check_maf`).  Returns the number
    of records written.  GT text of chunk k + 1 is formatted and copied while chunk k is written.
    """
    engine.require_cuda()
    H = int(bits.shape[1]) if bits is not None and bits.dim() == 4 else len(variants)
    if len(variants) != H or len(samples) != n_samples:
        raise ValueError("variants / samples do not match the packed result")
    rows = np.arange(H) if keep is None else np.asarray(keep, dtype=np.int64)
    fileobj.write(vcf_header(samples, variants["chrom"][rows]).encode())
    if len(rows) == 0:
        return 0
    fixed = vcf_fixed_columns(variants[rows])
    if n_samples == 0:
        for f in fixed:
            fileobj.write(f[:-1] + b"\n")
        return len(rows)
    # runs of consecutive haplotypes -> one kernel call each, at most chunk_haps rows
    chunk_haps = chunk_haps or max(1, min(4096, (256 << 20) // (4 * n_samples)))
    runs, a = [], 0
    while a < len(rows):
        b = a + 1
        while b < len(rows) and b - a < chunk_haps and rows[b] == rows[b - 1] + 1:
            b += 1
        runs.append((a, b))
        a = b
    dev = bits.device
    pitch = 4 * n_samples
    with torch.cuda.device(dev):
        s_cmp, cur = torch.cuda.Stream(dev), torch.cuda.current_stream(dev)
        s_cmp.wait_stream(cur)
        cap = max(b - a for a, b in runs)
        dbuf = [torch.empty((cap, pitch), dtype=torch.uint8, device=dev) for _ in range(2)]
        hbuf = [torch.empty((cap, pitch), dtype=torch.uint8, pin_memory=True) for _ in range(2)]
        events = [None, None]

        def launch(k):
            a, b = runs[k]
            with torch.cuda.stream(s_cmp):
                check(_cuda.lib().ht_format_vcf_gt(bits.data_ptr(), n_samples, H, int(rows[a]), b - a,
                                                  dbuf[k % 2].data_ptr(), pitch, s_cmp.cuda_stream), "ht_format_vcf_gt")
                hbuf[k % 2][: b - a].copy_(dbuf[k % 2][: b - a], non_blocking=True)
                events[k % 2] = s_cmp.record_event()

        launch(0)
        for k, (a, b) in enumerate(runs):
            if k + 1 < len(runs):
                launch(k + 1)
            events[k % 2].synchronize()
            body = hbuf[k % 2].numpy()
            for i in range(b - a):
                fileobj.write(fixed[a + i])
                fileobj.write(body[i].data)  # a contiguous row: written without a copy
        cur.wait_stream(s_cmp)
    return len(rows)


def transform_to_vcf(hp, gts, fileobj, maf: float = None, ancestry: bool = False, device=None) -> int:
    """
    `hp.transform(gts, hp_gt)` + optional `check_maf(maf, discard_also=True)` + `hp_gt.write()`
    (haptools/transform.py:669-676) with the result kept bit-packed on the device.  `hp` is a
    `Haplotypes` / `HaplotypesAncestry`, `gts` its genotypes; returns the number of records.
    """
    from .data.genotypes import GenotypesVCF
    from .plan import plan_from_haplotypes

    engine.require_cuda()
    hp.index()
    haps = [hp.data[h] for h in hp.type_ids["H"]]
    variants = np.empty(len(haps), dtype=GenotypesVCF(None).variants.dtype)
    variants["id"] = [h.id for h in haps]
    variants["chrom"] = [h.chrom for h in haps]
    variants["pos"] = [h.start for h in haps]
    variants["alleles"].fill(("A", "T"))
    plan = plan_from_haplotypes(haps, gts, ancestry=ancestry)
    if ancestry and gts.data.shape[2] != 2:
        raise ValueError(
            f"could not broadcast input array from shape {(gts.data.shape[0], gts.data.shape[2])} "
            f"into shape {(gts.data.shape[0], 2)}"
        )
    n = gts.data.shape[0]
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    keep = None
    bits = None
    if len(haps) and n:
        with torch.cuda.device(dev):
            # sample chunks stream through the device; only the bit-packed result stays resident
            bits, counts = engine.transform_host_bits(gts.data, plan, ancestry=gts.ancestry if ancestry else None,
                                                      device=dev, counts=True)
            if maf is not None:
                keep = maf_keep(counts, n, maf)
    elif maf is not None and len(haps):
        keep = np.arange(len(haps))
    return write_vcf_from_bits(fileobj, variants, gts.samples, bits, n, keep=keep)
