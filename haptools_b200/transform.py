"""
Ancestry-aware transform classes (haptools/transform.py:16-149) and the ancestry genotype
container (haptools/transform.py:152-186, 302-351):

    HaplotypeAncestry.transform    haptools/transform.py:31-68
    HaplotypesAncestry.transform   haptools/transform.py:91-149

A haplotype is present on a chromosome copy iff every one of its alleles matches AND the
local-ancestry label at every one of its variants equals the haplotype's population.  The
label comparison is folded into the bit-planes by the pack kernel.
"""
from __future__ import annotations

import logging
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

from . import data
from .data import haplotypes as _haplotypes  # noqa: F401


@dataclass
class HaplotypeAncestry(data.Haplotype):
    ancestry: str
    _extras: tuple = field(default=("ancestry",), init=False, repr=False)

    def transform(self, genotypes) -> np.ndarray:
        """Replaces haptools/transform.py:31-68; returns ``(n, 2)`` bool."""
        from .integrate import haplotype_ancestry_transform

        return haplotype_ancestry_transform(self, genotypes)


class HaplotypesAncestry(data.Haplotypes):
    def __init__(self, fname: Path | str = None, haplotype: type = HaplotypeAncestry,
                 variant: type = data.Variant, log: logging.Logger = None):
        super().__init__(fname, haplotype=haplotype, variant=variant, log=log)

    def transform(self, gts, hap_gts: data.GenotypesVCF = None) -> data.GenotypesVCF:
        """Replaces haptools/transform.py:91-149."""
        return self._transform(gts, hap_gts, ancestry=True)


class GenotypesAncestry(data.GenotypesVCF):
    """GenotypesVCF plus ``ancestry (n, p, 2) uint8`` and its label dictionaries
    (haptools/transform.py:152-186)."""

    def __init__(self, fname: Path | str = None, log: logging.Logger = None):
        super().__init__(fname, log)
        self.ancestry = None
        self.valid_labels = None
        self.ancestry_labels = {}
        self.popnum_ancestry = {}

    _MISSING_MIN = 255  # this class treats only 255 as missing (haptools/transform.py:360)

    def _discard_samples(self, samp_idx, level: str = "info"):
        # haptools/transform.py:363-373: the ancestry rows go too, and the line is logged at INFO
        self.ancestry = np.delete(self.ancestry, samp_idx, axis=0)
        super()._discard_samples(samp_idx, level="info")

    def _copy_fields_to(self, other):
        super()._copy_fields_to(other)
        other.ancestry, other.ancestry_labels = self.ancestry, self.ancestry_labels

    def _take(self, idx, axis):
        super()._take(idx, axis)
        self.ancestry = self.ancestry[idx, :] if axis == 0 else self.ancestry[:, idx]
