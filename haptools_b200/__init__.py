"""
haptools_b200 -- a B200-native implementation of ONE hot path of CAST-genomics/haptools:
the haplotype transform (Haplotypes.transform / Haplotype.transform and their
ancestry-aware variants).  The classes in haptools_b200.data and haptools_b200.transform
mirror the reference's interface for this path; their transform() bodies run hand-written
sm_100a CUDA kernels through a C-ABI shared library (include/haptools_cuda.h).
There is no CPU fallback.
"""
__version__ = "0.1.0"
