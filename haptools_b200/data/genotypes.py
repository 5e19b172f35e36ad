"""
In-memory genotype containers with the reference's attribute contract
(haptools/data/genotypes.py:22-73 ``Genotypes``, :712-736 ``GenotypesVCF``): ``data`` is an
``(n, p, 2|3)`` uint8/bool array, ``samples`` a tuple of IDs, ``variants`` a structured array.
File readers/writers (cyvcf2, pysam, pgenlib) are out of scope -- see DESIGN.md.
"""
from __future__ import annotations

import logging
from collections import Counter
from pathlib import Path

import numpy as np

_BASE_FIELDS = [("id", "U50"), ("chrom", "U10"), ("pos", np.uint32)]


def _duplicates(ids) -> list:
    dups = [i for i, c in Counter(ids).items() if c > 1]
    return dups[:5]


class Genotypes:
    """Container mirroring ``haptools.data.Genotypes`` (haptools/data/genotypes.py:22-73)."""

    def __init__(self, fname: Path | str = None, log: logging.Logger = None):
        self.fname = Path(fname) if isinstance(fname, str) else fname
        self.data = None
        self.log = log or logging.getLogger(f"haptools_b200.{type(self).__name__}")
        self.samples = tuple()
        self.variants = np.array([], dtype=_BASE_FIELDS)
        self._prephased = False
        self._samp_idx = None
        self._var_idx = None

    def __repr__(self):
        return str(self.fname)

    def unset(self) -> bool:
        return self.data is None

    def index(self, samples: bool = True, variants: bool = True):
        """ID -> position look-up tables; duplicate IDs raise the reference's ValueError
        (haptools/data/genotypes.py:346-379)."""
        if samples and self._samp_idx is None:
            self._samp_idx = {s: i for i, s in enumerate(self.samples)}
            if len(self._samp_idx) < len(self.samples):
                raise ValueError(f"Found duplicate sample IDs: {_duplicates(self.samples)}")
        if variants and self._var_idx is None:
            ids = self.variants["id"]
            self._var_idx = {v: i for i, v in enumerate(ids)}
            if len(self._var_idx) < len(ids):
                raise ValueError(f"Found duplicate variant IDs: {_duplicates(ids)}")

    def subset(self, samples: tuple = None, variants: tuple = None, inplace: bool = False):
        """Keep the given samples / variant IDs, in the given order
        (haptools/data/genotypes.py:381-442).  The transform itself never calls this -- the
        planner resolves columns in place -- it exists for callers that relied on it."""
        gts = self if inplace else self._blank_like()
        self._copy_fields_to(gts)
        self.index(samples=samples is not None, variants=variants is not None)
        if samples is not None:
            kept = tuple(s for s in samples if s in self._samp_idx)
            if len(kept) < len(samples):
                self.log.warning(
                    f"Saw {len(samples) - len(kept)} fewer samples than requested. "
                    f"Proceeding with {len(kept)} samples."
                )
            rows = tuple(self._samp_idx[s] for s in kept)
            gts.samples = kept
            if inplace:
                self._samp_idx = None
            gts._take(rows, axis=0)
        if variants is not None:
            cols = [self._var_idx[v] for v in variants if v in self._var_idx]
            if len(cols) < len(variants):
                self.log.warning(
                    f"Saw {len(variants) - len(cols)} fewer variants than requested. "
                    f"Proceeding with {len(cols)} variants."
                )
            gts.variants = self.variants[cols]
            if inplace:
                self._var_idx = None
            gts._take(cols, axis=1)
        if not inplace:
            return gts

    # helpers for subclasses that carry extra per-genotype arrays
    def _blank_like(self):
        return self.__class__(self.fname, self.log)

    def _copy_fields_to(self, other):
        other.samples, other.variants, other.data = self.samples, self.variants, self.data

    def _take(self, idx, axis):
        self.data = self.data[idx, :] if axis == 0 else self.data[:, idx]

    # ---- QC: the reference's numpy passes over the whole matrix run as ONE device pass
    # (engine.qc_host -> ht_qc_u8); messages, log lines and side effects are the reference's ----------------
    _MISSING_MIN = 254  # Genotypes.check_missing: >= max(uint8) - 1 (genotypes.py:463)

    def qc(self):
        """One device pass over ``self.data``: an engine.QCResult with everything the four checks need
        (a caller that runs several of them on unchanged data can pass it on as ``qc=``)."""
        from .. import engine

        return engine.qc_host(self.data, missing_min=self._MISSING_MIN)

    def _variant_fields(self, v: int):
        return tuple(self.variants[v])[:3]

    def check_missing(self, discard_also=False, qc=None):
        """haptools/data/genotypes.py:444-485."""
        qc = qc or self.qc()
        if qc.first_missing is not None:
            if discard_also:
                self._discard_samples(np.nonzero(qc.samp_missing)[0])
            else:
                s, v = qc.first_missing
                raise ValueError(
                    "Genotype with ID {} at POS {}:{} is missing for sample {}".format(
                        *self._variant_fields(v), self.samples[s]
                    )
                )
        if discard_also and not self.data.shape[0]:
            self.log.warning(
                "All samples were discarded! Check that that none of your variants are"
                " missing genotypes (GT: '.|.')."
            )

    def _discard_samples(self, samp_idx, level: str = "warning"):
        original_num_samples = len(self.samples)
        self.data = np.delete(self.data, samp_idx, axis=0)
        self.samples = tuple(np.delete(self.samples, samp_idx))
        getattr(self.log, level)(
            "Ignoring missing genotypes from "
            f"{original_num_samples - len(self.samples)} samples"
        )
        self._samp_idx = None

    def check_biallelic(self, discard_also=False, qc=None):
        """haptools/data/genotypes.py:487-528 (turns ``data`` into bool)."""
        if self.data.dtype == np.bool_:
            self.log.warning("All genotypes are already biallelic")
            return
        qc = qc or self.qc()
        if qc.first_multiallelic is not None:
            if discard_also:
                variant_idx = np.nonzero(qc.var_multiallelic)[0]
                # the reference counts (sample, variant) PAIRS here (len of np.nonzero's second array)
                self.log.info(f"Ignoring {qc.n_multiallelic} multiallelic variants")
                self.data = np.delete(self.data, variant_idx, axis=1)
                self.variants = np.delete(self.variants, variant_idx)
                self._var_idx = None
            else:
                s, v = qc.first_multiallelic
                raise ValueError(
                    "Variant with ID {} at POS {}:{} is multiallelic for sample {}".format(
                        *self._variant_fields(v), self.samples[s]
                    )
                )
        if discard_also and not self.data.shape[1]:
            self.log.warning(
                "All variants were discarded! Check that there are biallelic variants "
                "in your dataset."
            )
        self.data = self.data.astype(np.bool_)

    def check_phase(self, qc=None):
        """Drop the phase column (haptools/data/genotypes.py:530-557): raises if an unphased heterozygote
        -- ``(bool(g0) ^ bool(g1)) & ~bool(phase)`` -- is present, else keeps a strided view of the two
        strands."""
        if self._prephased or self.data.shape[2] < 3:
            self.log.warning("Phase information has already been removed from the data")
            return
        qc = qc or self.qc()
        if qc.first_unphased is not None:
            s, v = qc.first_unphased
            raise ValueError(
                "Variant with ID {} at POS {}:{} is unphased for sample {}".format(
                    *self._variant_fields(v), self.samples[s]
                )
            )
        self.data = self.data[:, :, :2]

    def check_maf(self, threshold: float = None, discard_also: bool = False, warn_only: bool = False, qc=None):
        """haptools/data/genotypes.py:559-625; returns the minor allele frequency of every variant."""
        qc = qc or self.qc()
        maf = qc.maf()
        if threshold is None:
            return maf
        rare_variants = maf < threshold
        if np.any(rare_variants):
            idx = np.nonzero(rare_variants)[0]
            if discard_also:
                original_num_variants = len(self.variants)
                self.data = np.delete(self.data, idx, axis=1)
                self.variants = np.delete(self.variants, idx)
                maf = np.delete(maf, idx)
                self.log.info(
                    f"Ignoring {original_num_variants - len(self.variants)} variants "
                    f"with MAF < {threshold}"
                )
                self._var_idx = None
            else:
                vals = self._variant_fields(idx[0]) + (maf[idx[0]], threshold)
                msg = "Variant with ID {} at POS {}:{} has MAF {} < {}".format(*vals)
                if warn_only:
                    self.log.warning(msg)
                else:
                    raise ValueError(msg)
        return maf


class GenotypesVCF(Genotypes):
    """``Genotypes`` whose ``variants`` also carry the allele strings
    (haptools/data/genotypes.py:712-736)."""

    def __init__(self, fname: Path | str = None, log: logging.Logger = None):
        super().__init__(fname, log)
        self.variants = np.array([], dtype=_BASE_FIELDS + [("alleles", object)])
