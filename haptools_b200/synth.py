"""
Synthetic inputs for the benchmark configs of BASELINE.json.

Genotypes / local-ancestry labels are a pure function of (seed, global sample index, variant,
strand) through splitmix64, so the device generator (csrc/ht_synth.cu) and the numpy
restatement below agree bit for bit: any sample slice of a device-generated, HBM-resident
cohort can be regenerated on the host and pushed through the oracle.

Haplotype sets are generated on the host directly as flat reference arrays (no Python
objects): haplotype h has k_h ~ U{2..20} alleles on distinct variants drawn from a local
window, allele REF/ALT with probability 1/2.
"""
from __future__ import annotations

import numpy as np

from .plan import TransformPlan, plan_from_refs

_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x: np.ndarray) -> np.ndarray:
    """numpy restatement of ht::splitmix64 (csrc/ht_common.cuh); uint64 in, uint64 out."""
    with np.errstate(over="ignore"):
        x = (np.asarray(x, dtype=np.uint64) + np.uint64(0x9E3779B97F4A7C15)) & _M
        x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M
        x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M
        return x ^ (x >> np.uint64(31))


def synth_gt(n: int, p: int, seed: int, mode: int = 0, missing_per_64k: int = 0,
             sample_offset: int = 0, variants: np.ndarray = None) -> np.ndarray:
    """(n, p, 2) uint8, identical to ht_synth_gt.  ``variants`` restricts to some columns."""
    s = (np.arange(n, dtype=np.uint64) + np.uint64(sample_offset))[:, None]
    v = (np.arange(p, dtype=np.uint64) if variants is None else np.asarray(variants, dtype=np.uint64))[None, :]
    with np.errstate(over="ignore"):
        hs = splitmix64(np.uint64(seed) + s)
    r = splitmix64(hs ^ v)
    out = np.empty((n, v.shape[1], 2), dtype=np.uint8)
    if mode == 0:
        out[:, :, 0] = (r & np.uint64(1)).astype(np.uint8)
        out[:, :, 1] = ((r >> np.uint64(1)) & np.uint64(1)).astype(np.uint8)
    elif mode == 1:
        rb = splitmix64(hs ^ (np.uint64(1 << 40) | (v >> np.uint64(6))))
        for c in range(2):
            f = (rb >> np.uint64(4 * c)) & np.uint64(15)
            g = splitmix64(splitmix64(np.uint64(seed) ^ np.uint64(0xF00D) ^ (f << np.uint64(48))) ^ v)
            out[:, :, c] = (g & np.uint64(1)).astype(np.uint8)
    else:
        raise ValueError("mode must be 0 or 1")
    if missing_per_64k:
        for c in range(2):
            miss = ((r >> np.uint64(16 + 16 * c)) & np.uint64(0xFFFF)) < np.uint64(missing_per_64k)
            out[:, :, c][miss] = 255
    return out


def synth_ancestry(n: int, p: int, seed: int, n_pops: int = 5, tract_len: int = 2000,
                   sample_offset: int = 0, variants: np.ndarray = None) -> np.ndarray:
    """(n, p, 2) uint8 labels, identical to ht_synth_ancestry."""
    s = (np.arange(n, dtype=np.uint64) + np.uint64(sample_offset))[:, None]
    v = (np.arange(p, dtype=np.uint64) if variants is None else np.asarray(variants, dtype=np.uint64))[None, :]
    with np.errstate(over="ignore"):
        hs = splitmix64(np.uint64(seed) + s)
    out = np.empty((n, v.shape[1], 2), dtype=np.uint8)
    for c in range(2):
        off = splitmix64(hs ^ np.uint64((2 << 40) | c))
        tract = (v + (off & np.uint64(0xFFFF))) // np.uint64(tract_len)
        with np.errstate(over="ignore"):
            lab = splitmix64((off >> np.uint64(16)) + tract) % np.uint64(n_pops)
        out[:, :, c] = lab.astype(np.uint8)
    return out


def synth_haplotype_refs(p: int, n_haps: int, seed: int, min_alleles: int = 2, max_alleles: int = 20,
                         n_pops: int = 0, sort: bool = False):
    """
    Flat haplotype definitions: returns (ref_col, ref_allele, hap_off, hap_label|None).
    Haplotype h: k_h ~ U{min..max} distinct variants at strictly increasing columns with gaps
    ~ U{1..4} from a uniformly drawn window start (so the window is <= 4*k_h wide), allele
    index ~ Bernoulli(1/2); with n_pops > 0 also a population label ~ U{0..n_pops-1}.
    `sort=True` orders the haplotypes by the position of their first variant, the order .hap
    files have after `haptools index` (tabix needs it); the default keeps generation order.
    """
    rng = np.random.default_rng(seed)
    k = rng.integers(min_alleles, max_alleles + 1, size=n_haps, dtype=np.int64)
    k = np.minimum(k, max(1, p // 4))
    hap_off = np.concatenate([[0], np.cumsum(k)])
    K = int(hap_off[-1])
    hap_of_ref = np.repeat(np.arange(n_haps), k)
    gaps = rng.integers(1, 5, size=K, dtype=np.int64)
    nz = k > 0  # haplotypes without alleles (min_alleles=0) own no entries
    first = hap_off[:-1]
    gaps[first[nz]] = 0
    csum = np.cumsum(gaps)
    within = csum - csum[np.minimum(first, max(K - 1, 0))][hap_of_ref] if K else csum
    span = np.zeros(n_haps, dtype=np.int64)  # offset of the last variant of each haplotype
    span[nz] = within[hap_off[1:][nz] - 1]
    start = (rng.random(n_haps) * (p - span)).astype(np.int64)
    ref_col = start[hap_of_ref] + within
    ref_allele = rng.integers(0, 2, size=K, dtype=np.int64)
    hap_label = rng.integers(0, n_pops, size=n_haps, dtype=np.int64) if n_pops else None
    if sort:
        order = np.argsort(start, kind="stable")
        idx = np.concatenate([np.arange(hap_off[h], hap_off[h + 1]) for h in order] + [np.zeros(0, dtype=np.int64)])
        ref_col, ref_allele = ref_col[idx], ref_allele[idx]
        hap_off = np.concatenate([[0], np.cumsum(k[order])])
        if hap_label is not None:
            hap_label = hap_label[order]
    return ref_col, ref_allele, hap_off, hap_label


def synth_plan(p: int, n_haps: int, seed: int, n_pops: int = 0, **kw) -> TransformPlan:
    ref_col, ref_allele, hap_off, hap_label = synth_haplotype_refs(p, n_haps, seed, n_pops=n_pops, **kw)
    ref_label = np.repeat(hap_label, np.diff(hap_off)) if hap_label is not None else None
    return plan_from_refs(ref_col, ref_allele, hap_off, p, ref_label)


def plan_to_csr(plan: TransformPlan):
    """(key_col, key_allele, hap_off, hap_key[, key_label]) for oracle.transform_from_csr."""
    return plan.key_col().astype(np.int64), plan.key_allele, plan.hap_off.astype(np.int64), plan.hap_key.astype(np.int64)


def synth_objects(p: int, n_haps: int, seed: int, n_pops: int = 0, sort: bool = False, gts=None):
    """
    The same synthetic haplotype set as synth_plan, as the OBJECTS the class API takes (variant
    IDs ``v{i}``, chrom "1", pos i + 1, alleles ("A", "C")): returns ``(haplotypes, gts)`` -- a
    ``Haplotypes`` / ``HaplotypesAncestry`` holding ``n_haps`` ``Haplotype`` objects (population ``P{k}`` with
    ``n_pops``), and a ``GenotypesVCF`` / ``GenotypesAncestry`` whose ``variants`` are filled in (``data``,
    ``samples`` and ``ancestry`` are the caller's to set).  ``plan_from_haplotypes(haps, gts)`` on them gives the
    tables of ``synth_plan(p, n_haps, seed, n_pops, sort=sort)``.
    """
    from . import data as hdata
    from . import transform as htransform

    ref_col, ref_allele, hap_off, hap_label = synth_haplotype_refs(p, n_haps, seed, n_pops=n_pops, sort=sort)
    ids = np.char.add("v", np.arange(p).astype("U"))
    if gts is None:
        gts = (htransform.GenotypesAncestry if n_pops else hdata.GenotypesVCF)(fname=None)
        variants = np.empty(p, dtype=gts.variants.dtype)
        variants["id"], variants["chrom"], variants["pos"] = ids, "1", np.arange(1, p + 1)
        variants["alleles"].fill(("A", "C"))
        gts.variants = variants
        if n_pops:
            gts.ancestry_labels = {f"P{k}": k for k in range(n_pops)}
    id_list = ids.tolist()
    cols, alleles, off = ref_col.tolist(), ref_allele.tolist(), hap_off.tolist()
    data = {}
    for h in range(n_haps):
        a, b = off[h], off[h + 1]
        start, end = (cols[a] + 1, cols[b - 1] + 2) if b > a else (1, 2)
        if n_pops:
            hap = htransform.HaplotypeAncestry("1", start, end, f"H{h}", f"P{int(hap_label[h])}")
        else:
            hap = hdata.Haplotype("1", start, end, f"H{h}")
        hap.variants = tuple(hdata.Variant(cols[j] + 1, cols[j] + 2, id_list[cols[j]], "C" if alleles[j] else "A") for j in range(a, b))
        data[hap.id] = hap
    haps = (htransform.HaplotypesAncestry if n_pops else hdata.Haplotypes)(fname=None)
    haps.data = data
    return haps, gts
