"""
ORACLE -- TEST INFRASTRUCTURE ONLY.

A CPU (numpy) restatement of the reference's haplotype-transform algorithm.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module; the product path
(``haptools_b200/``) never does and has no CPU fallback.

Parity status: PINNED.  ``tests/test_oracle.py`` checks every function below against
(a) the reference's own golden vectors for this path (i.e.
tests/test_data.py:1616-1695 and tests/test_transform.py:167-212 of the reference),
(b) golden fixtures under ``tests/golden/`` that were produced by importing the
unmodified reference in the build container (``tests/golden/make_golden.py``), and
(c) -- only where /root/reference is mounted -- the live reference on randomised
inputs (``oracle/ref_harness.py``).

Each function cites the reference lines it follows (paths relative to the reference
checkout, CAST-genomics/haptools v0.6.2).  The arithmetic is the reference's own:
a fancy-index column gather, ``np.equal`` against the per-key allele index, and a
per-haplotype ``np.all`` over the haplotype's key columns.
"""
from __future__ import annotations

from collections import Counter
from typing import Sequence

import numpy as np


# --------------------------------------------------------------------------------------
# host-side planning, as the reference does it
# --------------------------------------------------------------------------------------
def build_keys(haps: Sequence) -> tuple[dict, list]:
    """
    The (variant ID, allele) -> dense key index dict in first-seen order, and the
    per-haplotype arrays of key indices.

    Follows haptools/data/haplotypes.py:1375-1386 (and haptools/transform.py:103-122).
    """
    alleles: dict = {}
    idxs: list = [None] * len(haps)
    count = 0
    for i, hap in enumerate(haps):
        idxs[i] = np.empty(len(hap.variants), dtype=np.uintc)
        for j, variant in enumerate(hap.variants):
            key = (variant.id, variant.allele)
            if key not in alleles:
                alleles[key] = count
                count += 1
            idxs[i][j] = alleles[key]
    return alleles, idxs


def variant_index(variant_ids: Sequence[str]) -> dict:
    """
    ``id -> column`` with the reference's duplicate check.

    Follows haptools/data/genotypes.py:372-379 (``Genotypes.index``).
    """
    var_idx = dict(zip(variant_ids, range(len(variant_ids))))
    if len(var_idx) < len(variant_ids):
        duplicates = [v for v, c in Counter(variant_ids).items() if c > 1]
        a_few = 5 if len(duplicates) > 5 else len(duplicates)
        raise ValueError(f"Found duplicate variant IDs: {duplicates[:a_few]}")
    return var_idx


# --------------------------------------------------------------------------------------
# array-level cores
# --------------------------------------------------------------------------------------
def transform_core(
    gt_sub: np.ndarray,
    allele_arr: np.ndarray,
    idxs: Sequence[np.ndarray],
    anc_sub: np.ndarray = None,
    ancestries: np.ndarray = None,
) -> np.ndarray:
    """
    ``gt_sub`` is the key-gathered genotype array ``(n, A, 2)`` (what ``gts.subset``
    returns in the reference), ``allele_arr`` the ``(A,)`` allele index per key.

    Follows haptools/data/haplotypes.py:1404-1412; with ``anc_sub``/``ancestries`` it
    follows haptools/transform.py:137-148.
    """
    equality_arr = np.equal(allele_arr[np.newaxis, :, np.newaxis], gt_sub)
    n = gt_sub.shape[0]
    out = np.empty((n, len(idxs), 2), dtype=np.bool_)
    if anc_sub is None:
        for i in range(len(idxs)):
            out[:, i] = np.all(equality_arr[:, idxs[i]], axis=1)
    else:
        for i in range(len(idxs)):
            out[:, i] = np.logical_and(
                np.all(anc_sub[:, idxs[i]] == ancestries[i], axis=1),
                np.all(equality_arr[:, idxs[i]], axis=1),
            )
    return out


def transform_from_csr(
    gt_data: np.ndarray,
    key_col: np.ndarray,
    key_allele: np.ndarray,
    hap_off: np.ndarray,
    hap_key: np.ndarray,
    ancestry: np.ndarray = None,
    hap_label: np.ndarray = None,
) -> np.ndarray:
    """
    The same algorithm driven by flat arrays (used for the synthetic configs, where
    building a million Python ``Variant`` objects would only time the interpreter):
    ``key_col[a]`` is the column of key ``a`` in ``gt_data``, ``key_allele[a]`` its
    allele index, haplotype ``h`` owns keys ``hap_key[hap_off[h]:hap_off[h+1]]``.

    Gather = ``Genotypes.subset`` (haptools/data/genotypes.py:440), then
    haptools/data/haplotypes.py:1404-1412 / haptools/transform.py:137-148.
    """
    gt_sub = gt_data[:, key_col][:, :, :2]
    allele_arr = np.asarray(key_allele).astype(gt_data.dtype)
    idxs = [
        np.asarray(hap_key[hap_off[h] : hap_off[h + 1]], dtype=np.intp)
        for h in range(len(hap_off) - 1)
    ]
    anc_sub = None
    if ancestry is not None:
        anc_sub = ancestry[:, key_col]
    return transform_core(gt_sub, allele_arr, idxs, anc_sub, hap_label)


# --------------------------------------------------------------------------------------
# object-level restatements (duck-typed: work on the reference's objects and on the
# haptools_b200 host mirror alike)
# --------------------------------------------------------------------------------------
def haplotypes_transform(haps: Sequence, gts) -> np.ndarray:
    """
    ``Haplotypes.transform`` -> the ``hap_gts.data`` array, ``(n, H, 2)`` ``np.bool_``.

    Follows haptools/data/haplotypes.py:1375-1412.  ``haps`` is the list of Haplotype
    objects (``[self.data[h] for h in self.type_ids["H"]]``, :1364); ``gts`` needs
    ``.data`` and a structured ``.variants`` with ``id`` and ``alleles``.
    """
    alleles, idxs = build_keys(haps)
    var_idx = variant_index(list(gts.variants["id"]))
    cols = [var_idx[k[0]] for k in alleles if k[0] in var_idx]
    sub_variants = gts.variants[cols]
    gt_sub = gts.data[:, cols]
    try:
        allele_arr = np.array(
            [
                sub_variants[i]["alleles"].index(allele)
                for i, (vID, allele) in enumerate(alleles)
            ],
            dtype=gts.data.dtype,
        )
    except ValueError:
        raise ValueError("Some alleles were not present in the genotypes")
    return transform_core(gt_sub[:, :, :2], allele_arr, idxs)


def haplotype_transform(hap, gts) -> np.ndarray:
    """
    ``Haplotype.transform`` -> ``(n, 2)`` ``np.bool_``.

    Follows haptools/data/haplotypes.py:483-511.  Like the reference, a haplotype without
    variants raises ValueError here (a (1, 0) allele array does not broadcast against
    (n, 0, 2)); only the plural form yields all-True for it.
    """
    var_ids = tuple(var.id for var in hap.variants)
    var_idx = variant_index(list(gts.variants["id"]))
    cols = [var_idx[v] for v in var_ids if v in var_idx]
    if len(cols) < len(var_ids):
        missing_ids = set(var_ids) - set(gts.variants["id"][cols])
        raise ValueError(
            f"Variants {missing_ids} are present in haplotype '{hap.id}' but "
            "absent in the provided genotypes"
        )
    sub_variants = gts.variants[cols]
    try:
        allele_arr = np.array(
            [
                [
                    [sub_variants[i]["alleles"].index(var.allele)]
                    for i, var in enumerate(hap.variants)
                ]
            ]
        )
    except ValueError:
        raise ValueError("Some alleles were not present in the genotypes")
    return np.all(allele_arr == gts.data[:, cols][:, :, :2], axis=1)


def haplotypes_ancestry_transform(haps: Sequence, gts) -> np.ndarray:
    """
    ``HaplotypesAncestry.transform`` -> ``(n, H, 2)`` ``np.bool_``.

    Follows haptools/transform.py:103-149: allele index is ``int(allele != REF)``
    (:128-134), the whole last axis of ``gts.data`` is compared (:137), and each
    haplotype is additionally ANDed with ``ancestry == label`` over its variants
    (:144-148).  ``ancestries`` is a uint8 array (:112), so an unknown population
    (``-1``) raises OverflowError on numpy >= 2 exactly as the reference does.
    """
    alleles, idxs = build_keys(haps)
    ancestries = np.empty(len(haps), dtype=np.uint8)
    for i, hap in enumerate(haps):
        ancestries[i] = gts.ancestry_labels.get(hap.ancestry, -1)
    var_idx = variant_index(list(gts.variants["id"]))
    cols = [var_idx[k[0]] for k in alleles if k[0] in var_idx]
    sub_variants = gts.variants[cols]
    allele_arr = np.array(
        [
            int(allele != sub_variants[i]["alleles"][0])
            for i, (vID, allele) in enumerate(alleles)
        ],
        dtype=gts.data.dtype,
    )
    return transform_core(
        gts.data[:, cols], allele_arr, idxs, gts.ancestry[:, cols], ancestries
    )


def haplotype_ancestry_transform(hap, gts) -> np.ndarray:
    """
    ``HaplotypeAncestry.transform`` -> ``(n, 2)`` ``np.bool_``.

    Follows haptools/transform.py:38-68.
    """
    var_ids = tuple(var.id for var in hap.variants)
    var_idx = variant_index(list(gts.variants["id"]))
    cols = [var_idx[v] for v in var_ids if v in var_idx]
    if len(cols) < len(var_ids):
        missing_ids = set(var_ids) - set(gts.variants["id"][cols])
        raise ValueError(
            f"Variants {missing_ids} are present in haplotype '{hap.id}' but "
            "absent in the provided genotypes"
        )
    sub_variants = gts.variants[cols]
    allele_arr = np.array(
        [
            [
                [int(var.allele != sub_variants[i]["alleles"][0])]
                for i, var in enumerate(hap.variants)
            ]
        ]
    )
    hap_gts = np.all(allele_arr == gts.data[:, cols], axis=1)
    ancestry_label = gts.ancestry_labels.get(hap.ancestry, -1)
    ancestry_arr = np.all(gts.ancestry[:, cols] == ancestry_label, axis=1)
    return np.logical_and(hap_gts, ancestry_arr)
