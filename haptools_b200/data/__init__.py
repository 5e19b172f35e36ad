"""The slice of ``haptools.data`` that the transform hot path touches."""
from .breakpoints import Breakpoints, HapBlock
from .genotypes import Genotypes, GenotypesVCF
from .haplotypes import Haplotype, Haplotypes, Repeat, Variant

__all__ = ["Breakpoints", "HapBlock", "Genotypes", "GenotypesVCF", "Variant", "Haplotype", "Repeat", "Haplotypes"]
