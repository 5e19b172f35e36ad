"""
Host planner of the haplotype transform: turns haplotype objects (or flat arrays) into the
small index tables the CUDA kernels consume (include/haptools_cuda.h).

It mirrors, step by step, what the reference does on the host before its numpy arithmetic
(haptools/data/haplotypes.py:1363-1401, haptools/transform.py:96-135) -- including the
exceptions it raises -- but resolves ``key -> column of gts.data`` instead of copying the
columns (``gts.subset``, haptools/data/genotypes.py:440).

Vocabulary
  key    one distinct (variant column, allele index[, ancestry label]) a haplotype asks for;
         the reference's ``alleles`` dict entry, with the haplotype's ancestry label folded
         in for the ancestry-aware classes
  CSR    ``hap_key[hap_off[h]:hap_off[h+1]]`` = keys of haplotype ``h`` (the reference's
         ``idxs[h]``)
Keys are numbered by (column, allele, label) so that the keys of one variant are consecutive
(``var_key_off``): the pack kernel then reads every genotype column once.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Sequence

import numpy as np

_U32_MAX = 0xFFFFFFFF


@dataclass
class TransformPlan:
    """Index tables for one (haplotype set, genotype column layout) pair; all numpy."""

    n_cols: int  # number of variant columns of the genotype array the plan refers to
    var_col: np.ndarray  # (V,) uint32, ascending: referenced columns
    var_key_off: np.ndarray  # (V+1,) uint32
    key_allele: np.ndarray  # (A,) uint8
    key_label: np.ndarray | None  # (A,) uint8, ancestry plans only
    hap_off: np.ndarray  # (H+1,) uint32
    hap_key: np.ndarray  # (K,) uint32

    @property
    def n_haps(self) -> int:
        return len(self.hap_off) - 1

    @property
    def n_vars(self) -> int:
        return len(self.var_col)

    @property
    def n_keys(self) -> int:
        return len(self.key_allele)

    @property
    def n_refs(self) -> int:
        return len(self.hap_key)

    @property
    def has_ancestry(self) -> bool:
        return self.key_label is not None

    def col_var(self) -> np.ndarray:
        """Dense inverse of ``var_col``: (n_cols,) int32, -1 for unreferenced columns."""
        out = np.full(self.n_cols, -1, dtype=np.int32)
        out[self.var_col] = np.arange(self.n_vars, dtype=np.int32)
        return out

    def fast_tables(self):
        """
        Dense per-column tables for the single-pass sample-major pack kernels; returns
        (col_key01, col_key_range, var_other).

        plain plans     col_key01 (n_cols, 2) int32 = {key of allele 0, key of allele 1} (-1 =
                        none) for the variants whose keys are all plain alleles 0/1
        ancestry plans  col_key_range (n_cols, 2) int32 = {first key, key count} for the variants
                        whose keys all have allele <= 1 and label < 8
        var_other lists the remaining variants (packed by the generic kernel).
        """
        off = self.var_key_off.astype(np.int64)
        key_var = np.repeat(np.arange(self.n_vars), np.diff(off))
        bad_key = self.key_allele > 1
        if self.has_ancestry:
            bad_key = bad_key | (self.key_label > 7)
        simple_var = np.ones(self.n_vars, dtype=bool)
        simple_var[key_var[bad_key]] = False
        var_other = np.nonzero(~simple_var)[0].astype(np.uint32)
        cols = self.var_col.astype(np.int64)
        if self.has_ancestry:
            rng = np.zeros((self.n_cols, 2), dtype=np.int32)
            rng[cols[simple_var], 0] = off[:-1][simple_var]
            rng[cols[simple_var], 1] = np.diff(off)[simple_var]
            return None, rng, var_other
        k01 = np.full((self.n_cols, 2), -1, dtype=np.int32)
        ok = simple_var[key_var]
        k01[cols[key_var[ok]], self.key_allele[ok].astype(np.int64)] = np.nonzero(ok)[0].astype(np.int32)
        return k01, None, var_other

    def biallelic_tables(self):
        """(col_key01, var_other) of a plain plan; see fast_tables."""
        k01, _, other = self.fast_tables()
        if k01 is None:
            return np.full((self.n_cols, 2), -1, dtype=np.int32), np.arange(self.n_vars, dtype=np.uint32)
        return k01, other

    def quad_tables(self):
        """
        The key lists in the form the transform kernel walks them (csrc/ht_transform.cu): every
        list padded to whole quads of 4 by repeating its last key (AND is idempotent), entries
        are BYTE offsets of the (tile, key) record inside a tile (key * 256), and the low bits of
        the first entry of a quad carry flags: 1 = last quad of its haplotype, 2 = haplotype
        without alleles (one all-zero quad).  Returns (quad_off16, quads):
          quad_off16  (ceil(H/16) + 1,) uint32: first quad of every group of 16 haplotypes (the
                      unit one warp processes)
          quads       (Q, 4) uint32
        or (None, None) when a record offset does not fit 32 bits (>= 2^24 keys).
        """
        if self.n_keys >= (1 << 24) or self.n_haps == 0:
            return None, None
        H = self.n_haps
        off = self.hap_off.astype(np.int64)
        lens = np.diff(off)
        padded = np.maximum(4, (lens + 3) & ~3)
        qoff = np.concatenate([[0], np.cumsum(padded)])
        hap_of = np.repeat(np.arange(H), padded)
        j = np.arange(qoff[-1]) - qoff[hap_of]
        ln = lens[hap_of]
        src = off[hap_of] + np.minimum(j, np.maximum(ln - 1, 0))
        key = self.hap_key.astype(np.int64)
        val = np.where(ln > 0, key[np.minimum(src, max(len(key) - 1, 0))] << 8, 0) if len(key) else np.zeros(len(j), np.int64)
        first = (j & 3) == 0
        val = val | np.where(first & (j + 4 >= padded[hap_of]), 1, 0) | np.where(first & (ln == 0), 2, 0)
        quad_off16 = (qoff[np.minimum(np.arange(0, H + 16, 16), H)] // 4)[: (H + 15) // 16 + 1]
        return quad_off16.astype(np.uint32), val.astype(np.uint32).reshape(-1, 4)

    def block_spans(self, block: int = 128):
        """
        Key run of every block of `block` consecutive haplotypes (caller's order; the unit one CTA of the
        transform kernel handles): returns (blk_key_lo, blk_key_cnt), uint32 arrays of ceil(H / block) entries
        with cnt = max key - min key + 1 over the block's lists (0 for a block without keys).  Keys are numbered
        by variant column, so position-sorted haplotype sets (`haptools index`, haptools/index.py:31-94) have
        short runs; ht_transform_u8_staged stages them in shared memory.
        """
        H = self.n_haps
        nb = (H + block - 1) // block
        lo = np.zeros(nb, dtype=np.int64)
        cnt = np.zeros(nb, dtype=np.int64)
        if self.n_refs:
            off = self.hap_off.astype(np.int64)
            key = self.hap_key.astype(np.int64)
            first = off[np.minimum(np.arange(nb) * block, H)]
            last = off[np.minimum((np.arange(nb) + 1) * block, H)]
            has = last > first
            starts = first[has]  # consecutive non-empty starts delimit exactly the blocks' reference ranges
            mn = np.minimum.reduceat(key, starts)
            mx = np.maximum.reduceat(key, starts)
            lo[has] = mn
            cnt[has] = mx - mn + 1
        return lo.astype(np.uint32), cnt.astype(np.uint32)

    def stage_lines(self, max_lines: int = 608, block: int = 128) -> int:
        """
        Shared-memory budget (in 128-byte lines) worth giving ht_transform_u8_staged for this plan, or 0 when
        staging does not pay: the smallest multiple of 32 that lets the blocks holding >= 95 % of the haplotype
        alleles stage their runs, provided those runs are at most half as many lines as the L2 reads they replace.
        """
        if self.n_refs == 0 or self.n_haps == 0:
            return 0
        _, cnt = self.block_spans(block)
        cnt = cnt.astype(np.int64)
        off = self.hap_off.astype(np.int64)
        nb = len(cnt)
        refs = off[np.minimum((np.arange(nb) + 1) * block, self.n_haps)] - off[np.minimum(np.arange(nb) * block, self.n_haps)]
        order = np.argsort(cnt, kind="stable")
        covered = np.cumsum(refs[order])
        k = int(np.searchsorted(covered, 0.95 * self.n_refs, side="left"))
        need = int(cnt[order[min(k, nb - 1)]])
        lines = (need + 31) // 32 * 32
        if lines == 0 or lines > max_lines:
            return 0
        ok = cnt <= lines
        return lines if 2 * int(cnt[ok].sum()) <= int(refs[ok].sum()) else 0

    def key_col(self) -> np.ndarray:
        """Column of each key, (A,) uint32."""
        return np.repeat(self.var_col, np.diff(self.var_key_off.astype(np.int64)))

    def algorithmic_bytes(self, n_samples: int) -> dict:
        """compulsory bytes of the pack and transform kernels."""
        words = (n_samples + 31) // 32
        plane = self.n_keys * 2 * words * 4
        gt = n_samples * self.n_vars * 2
        out = n_samples * self.n_haps * 2
        pack = gt * (2 if self.has_ancestry else 1) + plane
        return {"pack": pack, "transform": plane + out, "total": pack + plane + out, "planes": plane, "gt": gt, "out": out}


def plan_from_refs(
    ref_col: np.ndarray,
    ref_allele: np.ndarray,
    hap_off: np.ndarray,
    n_cols: int,
    ref_label: np.ndarray | None = None,
) -> TransformPlan:
    """
    Build a plan from one (column, allele[, label]) triple per haplotype-variant reference:
    reference ``j`` belongs to the haplotype ``h`` with ``hap_off[h] <= j < hap_off[h+1]``.
    De-duplicates the triples into keys (haptools/data/haplotypes.py:1382-1386).
    """
    ref_col = np.asarray(ref_col, dtype=np.int64)
    ref_allele = np.asarray(ref_allele, dtype=np.int64)
    hap_off = np.asarray(hap_off, dtype=np.int64)
    n_refs = len(ref_col)
    if len(hap_off) == 0 or hap_off[0] != 0 or hap_off[-1] != n_refs or np.any(np.diff(hap_off) < 0):
        raise ValueError("hap_off is not a valid CSR offset array")
    if n_refs > _U32_MAX or n_cols > _U32_MAX:
        raise ValueError("too many haplotype alleles / variants for 32-bit indices")
    if n_refs and (ref_col.min() < 0 or ref_col.max() >= n_cols):
        raise ValueError("a haplotype references a variant column outside the genotypes")
    if n_refs and (ref_allele.min() < 0 or ref_allele.max() > 255):
        raise ValueError("allele indices must fit in a uint8")
    packed = (ref_col << 16) | (ref_allele << 8)
    if ref_label is not None:
        ref_label = np.asarray(ref_label, dtype=np.int64)
        if n_refs and (ref_label.min() < 0 or ref_label.max() > 255):
            raise ValueError("ancestry labels must fit in a uint8")
        packed = packed | ref_label
    keys, inverse = np.unique(packed, return_inverse=True)
    key_col = (keys >> 16).astype(np.uint32)
    var_col, first = np.unique(key_col, return_index=True)
    var_key_off = np.concatenate([first, [len(keys)]]).astype(np.uint32)
    return TransformPlan(
        n_cols=int(n_cols),
        var_col=var_col.astype(np.uint32),
        var_key_off=var_key_off,
        key_allele=((keys >> 8) & 0xFF).astype(np.uint8),
        key_label=(keys & 0xFF).astype(np.uint8) if ref_label is not None else None,
        hap_off=hap_off.astype(np.uint32),
        hap_key=inverse.reshape(-1).astype(np.uint32),
    )


def _variant_columns(gts, var_ids: Sequence[str], who: str) -> list:
    """``id -> column`` through ``Genotypes.index`` (duplicate IDs raise there,
    haptools/data/genotypes.py:372-379); absent IDs raise instead of the reference's
    misaligned gather."""
    gts.index(samples=False, variants=True)
    var_idx = gts._var_idx
    missing = [v for v in var_ids if v not in var_idx]
    if missing:
        raise ValueError(
            f"Variants {set(missing)} are present in {who} but absent in the provided genotypes"
        )
    return [var_idx[v] for v in var_ids]


def plan_from_haplotypes_loop(haps: Sequence, gts, ancestry: bool = False, who: str = None,
                         unknown_label: str = "raise", clip_bool: bool = True):
    """
    The literal, per-allele Python loop of the reference (kept as the cross-check of the
    vectorised plan_from_haplotypes below; tests/test_plan.py compares the two).

    ``haps``: Haplotype objects (``.variants`` of objects with ``.id``/``.allele``; with
    ``ancestry=True`` also ``.ancestry``); ``gts``: a GenotypesVCF-like object with
    ``.data``, a structured ``.variants`` (``id``, ``alleles``) and, with ``ancestry=True``,
    ``.ancestry_labels``.

    Follows haptools/data/haplotypes.py:1375-1401 (plain) and haptools/transform.py:103-135
    (ancestry).  ``unknown_label="raise"`` is the plural form's behaviour for a population
    that is not in ``ancestry_labels`` (OverflowError from ``uint8 <- -1`` on numpy >= 2,
    haptools/transform.py:115); ``"never"`` is the single-haplotype form's (-1 never equals a
    label, haptools/transform.py:62-67): returns ``(plan, never)`` with ``never[h]`` True for
    haplotypes that cannot be present.
    """
    # (variant ID, allele) -> index, in first-seen order, as the reference builds it
    alleles: dict = {}
    idxs = [None] * len(haps)
    ancestries = np.empty(len(haps), dtype=np.uint8) if ancestry else None
    never = np.zeros(len(haps), dtype=np.bool_)
    count = 0
    for i, hap in enumerate(haps):
        if ancestry:
            label = gts.ancestry_labels.get(hap.ancestry, -1)
            if label == -1 and unknown_label == "never":
                never[i] = len(hap.variants) > 0
                label = 0
            # uint8 <- -1 raises OverflowError on numpy >= 2, exactly like transform.py:115
            ancestries[i] = label
        idxs[i] = np.empty(len(hap.variants), dtype=np.uintc)
        for j, variant in enumerate(hap.variants):
            key = (variant.id, variant.allele)
            if key not in alleles:
                alleles[key] = count
                count += 1
            idxs[i][j] = alleles[key]
    cols = _variant_columns(gts, [k[0] for k in alleles], who or "the haplotypes")
    var_alleles = gts.variants["alleles"] if len(alleles) else ()
    as_bool = gts.data.dtype == np.bool_
    if ancestry:
        # biallelic semantics: any non-REF string is 1 (haptools/transform.py:128-134)
        allele_arr = [int(allele != var_alleles[c][0]) for c, (_, allele) in zip(cols, alleles)]
    else:
        try:
            allele_arr = [var_alleles[c].index(allele) for c, (_, allele) in zip(cols, alleles)]
        except ValueError:
            raise ValueError("Some alleles were not present in the genotypes")
    allele_arr = np.asarray(allele_arr, dtype=np.int64).reshape(-1)
    if as_bool and clip_bool:
        # the reference builds allele_arr with dtype=bool (haplotypes.py:1398)
        allele_arr = np.minimum(allele_arr, 1)
    lens = np.fromiter((len(x) for x in idxs), dtype=np.int64, count=len(idxs))
    hap_off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    flat = np.concatenate(idxs).astype(np.int64) if len(idxs) and hap_off[-1] else np.zeros(0, dtype=np.int64)
    cols = np.asarray(cols, dtype=np.int64).reshape(-1)
    ref_label = np.repeat(ancestries.astype(np.int64), lens) if ancestry else None
    plan = plan_from_refs(cols[flat], allele_arr[flat], hap_off, gts.data.shape[1], ref_label)
    return (plan, never) if unknown_label == "never" else plan


def _allele_indices(var_alleles, cols: np.ndarray, allele_strs: list, ancestry: bool) -> np.ndarray:
    """
    ``alleles.index(allele)`` (haptools/data/haplotypes.py:1395) -- or ``int(allele != REF)`` for
    the ancestry classes (haptools/transform.py:130) -- for every haplotype allele at once:
    REF/ALT1 are resolved with two vectorised string comparisons, only the alleles that match
    neither (multi-allelic sites) go through ``tuple.index``.
    """
    n = len(cols)
    if n == 0:
        return np.zeros(0, dtype=np.int64)
    want = np.asarray(allele_strs, dtype=object)
    ucols, inv = np.unique(cols, return_inverse=True)
    tuples = [var_alleles[c] for c in ucols]
    ref = np.asarray([t[0] if len(t) > 0 else None for t in tuples], dtype=object)[inv]
    is_ref = want == ref
    if ancestry:
        return (~is_ref).astype(np.int64)
    alt = np.asarray([t[1] if len(t) > 1 else None for t in tuples], dtype=object)[inv]
    out = np.where(is_ref, 0, np.where(want == alt, 1, -1)).astype(np.int64)
    rest = np.nonzero(out < 0)[0]
    try:
        for j in rest:  # ValueError from tuple.index if the allele is not there
            out[j] = tuples[inv[j]].index(allele_strs[j])
    except ValueError:
        raise ValueError("Some alleles were not present in the genotypes")
    return out


def plan_from_haplotypes(haps: Sequence, gts, ancestry: bool = False, who: str = None,
                         unknown_label: str = "raise", clip_bool: bool = True, log=None, n_distinct_out: list = None):
    """
    Same contract as plan_from_haplotypes_loop (the reference's host steps,
    haptools/data/haplotypes.py:1375-1401 / haptools/transform.py:103-135, exceptions
    included) without the per-allele Python loop: the haplotype objects are flattened with list
    comprehensions, IDs resolve to columns through ``Genotypes.index``'s dict in one pass,
    allele strings to indices with vectorised comparisons, and plan_from_refs de-duplicates
    (column, allele[, label]) triples with one np.unique.  (De-duplicating by (column, allele
    index) instead of the reference's (ID, allele string) yields the same matrix: equal keys
    give equal planes.)

    clip_bool  the PLURAL forms build ``allele_arr`` with ``dtype=gts.data.dtype`` (haplotypes.py:1398), so
               with bool genotypes an allele index >= 2 becomes True and matches like 1; the SINGLE forms
               (haplotypes.py:499-511) keep an integer array, and ``2 == True`` is False -- pass
               ``clip_bool=False`` there (a key with allele >= 2 then never matches 0/1 bytes).
    log        a logger that receives, at the step each belongs to, the two debug lines the reference emits
               before its arithmetic (haplotypes.py:1387, 1389 / transform.py:122, 124), verbatim:
               ``len(alleles)`` there counts distinct (variant ID, allele STRING) pairs.
    """
    H = len(haps)
    lens = np.fromiter((len(h.variants) for h in haps), dtype=np.int64, count=H)
    hap_off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    never = np.zeros(H, dtype=np.bool_)
    ref_label = None
    if ancestry:
        labels = gts.ancestry_labels
        lab = np.fromiter((labels.get(h.ancestry, -1) for h in haps), dtype=np.int64, count=H)
        unknown = lab == -1
        if unknown.any():
            if unknown_label == "never":
                never = unknown & (lens > 0)
                lab = np.where(unknown, 0, lab)
            else:
                # uint8 <- -1 raises OverflowError on numpy >= 2, exactly like transform.py:115
                np.empty(1, dtype=np.uint8)[0] = -1
        lab.astype(np.uint8, casting="unsafe")  # labels are uint8 in the reference (transform.py:104)
        ref_label = np.repeat(lab & 0xFF, lens)
    ids = [v.id for h in haps for v in h.variants]
    allele_strs = [v.allele for h in haps for v in h.variants]
    want_count = log is not None or n_distinct_out is not None
    logged = [0]

    def log_lines(upto: int, n_distinct=None):
        """The reference's two debug lines (haplotypes.py:1387, 1389), each once and in order.  `len(alleles)` there
        counts distinct (variant ID, allele STRING) pairs: a set of 1.1 M pairs costs 0.1 s, so on the normal path
        the count comes from the resolved (column, allele index) pairs instead (IDs are unique and `tuple.index`
        is injective on a variant's alleles) -- except for the ancestry classes, whose index int(allele != REF)
        is not."""
        if logged[0] < 1 <= upto and want_count:
            if n_distinct is None:
                n_distinct = len(set(zip(ids, allele_strs)))
            if n_distinct_out is not None:
                n_distinct_out.append(n_distinct)
            if log is not None:
                log.debug(f"Copying genotypes for {n_distinct} distinct alleles")
        if logged[0] < 2 <= upto and log is not None:
            log.debug("Creating array denoting alt allele status")
        logged[0] = max(logged[0], upto)

    gts.index(samples=False, variants=True)
    var_idx = gts._var_idx
    try:
        cols = np.fromiter(map(var_idx.__getitem__, ids), dtype=np.int64, count=len(ids))
    except KeyError:
        log_lines(1)  # the reference logs its first line before it looks the variants up
        missing = [v for v in dict.fromkeys(ids) if v not in var_idx]
        raise ValueError(
            f"Variants {set(missing)} are present in {who or 'the haplotypes'} but absent in the provided genotypes"
        )
    var_alleles = gts.variants["alleles"] if len(ids) else ()
    try:
        allele_idx = _allele_indices(var_alleles, cols, allele_strs, ancestry)
    except ValueError:
        log_lines(2)  # both lines precede the reference's allele look-up
        raise
    n_distinct = None
    if want_count and not ancestry and len(cols):
        n_alleles = int(allele_idx.max()) + 1
        n_distinct = len(np.unique(cols * n_alleles + allele_idx))
    elif want_count and not len(cols):
        n_distinct = 0
    log_lines(2, n_distinct)
    if clip_bool and gts.data.dtype == np.bool_:
        # the reference builds allele_arr with dtype=bool (haplotypes.py:1398)
        allele_idx = np.minimum(allele_idx, 1)
    plan = plan_from_refs(cols, allele_idx, hap_off, gts.data.shape[1], ref_label)
    return (plan, never) if unknown_label == "never" else plan


# ----------------------------------------------------------------------------------------
# locality tables of the two-phase ("local") transform
# ----------------------------------------------------------------------------------------
@dataclass
class LocalTables:
    """
    Haplotypes re-ordered by the position of their first key and cut into blocks whose keys
    fall into a short contiguous key range (keys are numbered by variant column, so a block of
    neighbouring haplotypes touches a contiguous run of (tile, key) records).  The local
    transform stages that run in shared memory once per (tile, block) instead of reading one
    256-byte record from L2 per haplotype allele, writes the bit-packed result in this order
    and a second kernel unpacks it into the caller's haplotype order (ht_transform_u8_local).

    order[j]        haplotype (caller's index) at sorted position j
    pos[h]          sorted position of haplotype h (inverse of order)
    blk_hap_off     (NB+1,) sorted positions: block b holds positions blk_hap_off[b]..[b+1]-1
    blk_key_lo      (NB,)   first key of the block's range
    blk_key_cnt     (NB,)   number of keys in the range; 0 = not staged (the block's haplotype
                    lists then hold global key numbers)
    s_hap_off       (H+1,)  CSR offsets in sorted order
    s_hap_key       (K,)    keys in sorted order, RELATIVE to blk_key_lo for staged blocks
    """

    dmax: int
    order: np.ndarray
    pos: np.ndarray
    blk_hap_off: np.ndarray
    blk_key_lo: np.ndarray
    blk_key_cnt: np.ndarray
    s_hap_off: np.ndarray
    s_hap_key: np.ndarray
    staged_reads: int  # (tile, key) records read per tile by phase 1

    @property
    def n_blocks(self) -> int:
        return len(self.blk_key_lo)

    def worthwhile(self, n_refs: int) -> bool:
        """Fewer L2 accesses than the one-kernel transform?  Per tile, in 256-byte records: the
        direct kernel reads one record per haplotype allele (n_refs); the local path reads
        `staged_reads` and moves the packed result four times (written, evicted, re-read)."""
        n_haps = len(self.order)
        return self.staged_reads + 4 * n_haps < 0.9 * n_refs


LOCAL_DMAX = 224  # records staged per block: 56 KB of shared memory (+ 8 KB of key lists), 3 CTAs per SM
LOCAL_MAX_BLOCK = 128  # haplotypes per block at most


def build_local_tables(plan: TransformPlan, dmax: int = LOCAL_DMAX, max_block: int = LOCAL_MAX_BLOCK) -> LocalTables:
    H, K = plan.n_haps, plan.n_refs
    off = plan.hap_off.astype(np.int64)
    key = plan.hap_key.astype(np.int64)
    lens = np.diff(off)
    nonempty = lens > 0
    mn = np.full(H, -1, dtype=np.int64)
    mx = np.full(H, -1, dtype=np.int64)
    if K:
        starts = off[:-1][nonempty]  # consecutive non-empty starts delimit exactly the lists
        mn[nonempty] = np.minimum.reduceat(key, starts)
        mx[nonempty] = np.maximum.reduceat(key, starts)
    order = np.argsort(mn, kind="stable")
    pos = np.empty(H, dtype=np.int64)
    pos[order] = np.arange(H)
    s_lens = lens[order]
    s_off = np.concatenate([[0], np.cumsum(s_lens)]).astype(np.int64)
    # gather the key lists in sorted order
    src = np.repeat(off[:-1][order] - s_off[:-1], s_lens) + np.arange(K)
    s_key = key[src]
    smn, smx = mn[order], mx[order]
    blk_off, blk_lo, blk_cnt = [0], [], []
    staged_reads = 0
    j = 0
    while j < H:
        e = min(H, j + max_block)
        cm = np.maximum.accumulate(smx[j:e])
        if cm[-1] < 0:  # only haplotypes without alleles
            take, lo, cnt = e - j, 0, 0
        else:
            first = int(np.argmax(smn[j:e] >= 0))
            lo = int(smn[j + first])
            take = int(np.searchsorted(cm, lo + dmax, side="left"))  # cm - lo + 1 <= dmax
            if take <= first:  # the first haplotype with alleles alone is wider than dmax
                take, cnt = first + 1, 0
                staged_reads += int(s_lens[j + first])
            else:
                cnt = int(cm[take - 1]) - lo + 1
                staged_reads += cnt
        blk_lo.append(lo)
        blk_cnt.append(cnt)
        j += take
        blk_off.append(j)
    blk_off = np.asarray(blk_off, dtype=np.int64)
    blk_lo = np.asarray(blk_lo, dtype=np.int64)
    blk_cnt = np.asarray(blk_cnt, dtype=np.int64)
    # relative keys for staged blocks
    blk_of_hap = np.repeat(np.arange(len(blk_lo)), np.diff(blk_off))
    sub = np.where(blk_cnt > 0, blk_lo, 0)[blk_of_hap]
    s_key = s_key - np.repeat(sub, s_lens)
    return LocalTables(
        dmax=int(dmax), order=order.astype(np.uint32), pos=pos.astype(np.uint32),
        blk_hap_off=blk_off.astype(np.uint32), blk_key_lo=blk_lo.astype(np.uint32),
        blk_key_cnt=blk_cnt.astype(np.uint32), s_hap_off=s_off.astype(np.uint32),
        s_hap_key=s_key.astype(np.uint32), staged_reads=int(staged_reads),
    )
