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


def _reject_empty(hap, genotypes):
    """The reference's single-haplotype form cannot broadcast its (1, 0) allele array against
    (n, 0, 2) and raises (haptools/data/haplotypes.py:499-511); only the plural form maps a
    haplotype without variants to all-True."""
    if len(hap.variants) == 0:
        n = genotypes.data.shape[0]
        raise ValueError(f"operands could not be broadcast together with shapes (1,0) ({n},0,2) ")


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
        from .. import engine
        from ..plan import plan_from_haplotypes

        # clip_bool=False: this form compares an INTEGER allele array (haplotypes.py:499-511), so with bool
        # genotypes an allele index >= 2 never matches (the plural form's dtype=bool cast does not apply)
        plan = plan_from_haplotypes([self], genotypes, ancestry=False, who=f"haplotype '{self.id}'", clip_bool=False)
        _reject_empty(self, genotypes)
        return engine.transform_host(genotypes.data, plan)[:, 0]


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
        from .. import engine
        from ..plan import plan_from_haplotypes

        self.index()
        haps = [self.data[hap] for hap in self.type_ids["H"]]
        if hap_gts is None:
            hap_gts = GenotypesVCF(fname=None, log=self.log)
        hap_gts.samples = gts.samples
        # haptools/data/haplotypes.py:1369-1372, column by column instead of a tuple per haplotype
        variants = np.empty(len(haps), dtype=hap_gts.variants.dtype)
        variants["id"] = [hap.id for hap in haps]
        variants["chrom"] = [hap.chrom for hap in haps]
        variants["pos"] = [hap.start for hap in haps]
        variants["alleles"].fill(("A", "T"))
        hap_gts.variants = variants
        # the five log lines of haplotypes.py:1387-1410 / transform.py:122-143, verbatim and in order (the first
        # two are emitted by the planner at the steps they belong to)
        plan = plan_from_haplotypes(haps, gts, ancestry=ancestry, log=self.log)
        self.log.info(f"Transforming genotypes for {len(haps)} haplotypes")
        self.log.debug(
            f"Allocating array with dtype {gts.data.dtype} and size "
            f"{(len(gts.samples), len(haps), 2)}"
        )
        self.log.debug("Computing haplotype genotypes. This may take a while")
        if ancestry and gts.data.shape[2] != 2:
            # haptools/transform.py:137 compares the whole last axis; a leftover phase column
            # makes the reference fail with a broadcast error -- same exception class here
            raise ValueError(
                f"could not broadcast input array from shape {(gts.data.shape[0], gts.data.shape[2])} "
                f"into shape {(gts.data.shape[0], 2)}"
            )
        hap_gts.data = engine.transform_host(gts.data, plan, ancestry=gts.ancestry if ancestry else None)
        return hap_gts

    def transform(self, gts: GenotypesVCF, hap_gts: GenotypesVCF = None) -> GenotypesVCF:
        """
        Fill (and return) ``hap_gts`` with the ``(n, H, 2)`` bool presence matrix of every
        haplotype.  Replaces haptools/data/haplotypes.py:1339-1413.
        """
        return self._transform(gts, hap_gts, ancestry=False)
