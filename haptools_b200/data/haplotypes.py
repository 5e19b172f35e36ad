"""
The in-memory model of a .hap file that the transform consumes
(haptools/data/haplotypes.py:118-164 ``Variant``, :298-370 ``Haplotype``, :722-780 and
:977-991 ``Haplotypes``) and the two transform methods whose bodies this project replaces:

    Haplotype.transform    haptools/data/haplotypes.py:463-511
    Haplotypes.transform   haptools/data/haplotypes.py:1339-1413

Same signatures, return types, side effects on ``hap_gts``, log lines and exceptions; the
arithmetic runs on the GPU (haptools_b200.engine).  .hap parsing/writing is out of scope.
"""
from __future__ import annotations

import logging
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

from .genotypes import GenotypesVCF


@dataclass
class Variant:
    start: int
    end: int
    id: str
    allele: str
    _extras: tuple = field(default=tuple(), init=False, repr=False)

    @property
    def ID(self):
        return self.id


@dataclass
class Repeat:
    """A tandem-repeat line of a .hap file; skipped by transform (haplotypes.py:1364)."""

    chrom: str
    start: int
    end: int
    id: str
    _extras: tuple = field(default=tuple(), init=False, repr=False)


@dataclass
class Haplotype:
    chrom: str
    start: int
    end: int
    id: str
    variants: tuple = field(default_factory=tuple, init=False)
    _extras: tuple = field(default=tuple(), init=False, repr=False)

    def __len__(self):
        return len(self.variants) if self.variants is not None else 0

    @property
    def ID(self):
        return self.id

    @property
    def varIDs(self):
        return tuple(var.id for var in self.variants)

    def transform(self, genotypes: GenotypesVCF) -> np.ndarray:
        """
        Presence of this haplotype on each chromosome copy of each sample: ``(n, 2)`` bool.
        Replaces haptools/data/haplotypes.py:463-511.
        """
        from ..integrate import haplotype_transform

        return haplotype_transform(self, genotypes)


class Haplotypes:
    """A set of haplotypes keyed by ID (haptools/data/haplotypes.py:722-780)."""

    def __init__(self, fname: Path | str = None, haplotype: type = Haplotype, variant: type = Variant,
                 repeat: type = Repeat, log: logging.Logger = None):
        self.fname = Path(fname) if isinstance(fname, str) else fname
        self.data = None
        self.log = log or logging.getLogger(f"haptools_b200.{type(self).__name__}")
        self.types = {"H": haplotype, "V": variant, "R": repeat}
        self.type_ids = None

    def __len__(self):
        return len(self.data) if self.data is not None else 0

    def index(self, force: bool = False):
        """(Re)build ``type_ids``; a no-op once built unless ``force``
        (haptools/data/haplotypes.py:977-991)."""
        if self.type_ids is not None and not force:
            return
        self.type_ids = {"H": [], "R": []}
        for key, value in self.data.items():
            if isinstance(value, Haplotype):
                self.type_ids["H"].append(key)
            if isinstance(value, Repeat):
                self.type_ids["R"].append(key)

    # shared by the ancestry-aware subclass (haptools_b200/transform.py)
    def _transform(self, gts, hap_gts, ancestry: bool):
        from ..integrate import haplotypes_transform

        return haplotypes_transform(self, gts, hap_gts, ancestry, GenotypesVCF)

    def transform(self, gts: GenotypesVCF, hap_gts: GenotypesVCF = None) -> GenotypesVCF:
        """
        Fill (and return) ``hap_gts`` with the ``(n, H, 2)`` bool presence matrix of every
        haplotype.  Replaces haptools/data/haplotypes.py:1339-1413.
        """
        return self._transform(gts, hap_gts, ancestry=False)
