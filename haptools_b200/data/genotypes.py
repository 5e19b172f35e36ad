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

    def check_phase(self):
        """Drop the phase column (haptools/data/genotypes.py:530-557): raises if an
        unphased heterozygote is present, else keeps a strided view of the two strands."""
        if self.data.shape[2] < 3:
            self.log.warning("Phase information has already been removed from the data")
            return
        het_unphased = np.logical_and(self.data[:, :, 0] != self.data[:, :, 1], ~self.data[:, :, 2].astype(bool))
        if het_unphased.any():
            s, v = np.argwhere(het_unphased)[0]
            raise ValueError(
                f"Variant with ID {self.variants['id'][v]} at POS {self.variants['chrom'][v]}:"
                f"{self.variants['pos'][v]} is unphased for sample {self.samples[s]}"
            )
        self.data = self.data[:, :, :2]


class GenotypesVCF(Genotypes):
    """``Genotypes`` whose ``variants`` also carry the allele strings
    (haptools/data/genotypes.py:712-736)."""

    def __init__(self, fname: Path | str = None, log: logging.Logger = None):
        super().__init__(fname, log)
        self.variants = np.array([], dtype=_BASE_FIELDS + [("alleles", object)])
