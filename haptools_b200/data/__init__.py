"""The slice of ``haptools.data`` that the transform hot path touches."""
from .genotypes import Genotypes, GenotypesVCF
from .haplotypes import Haplotype, Haplotypes, Repeat, Variant

__all__ = ["Genotypes", "GenotypesVCF", "Variant", "Haplotype", "Repeat", "Haplotypes"]
