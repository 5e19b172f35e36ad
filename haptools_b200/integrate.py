"""
The reference-side binding, executable: ``install(haptools)`` replaces the bodies of the reference's four
transform methods with calls into this package (planner -> C-ABI library), leaving every signature, return type,
side effect on ``hap_gts``, log line and exception where it was:

    haptools.data.Haplotype.transform            haptools/data/haplotypes.py:463-511
    haptools.data.Haplotypes.transform           haptools/data/haplotypes.py:1339-1413
    haptools.transform.HaplotypeAncestry.transform    haptools/transform.py:31-68
    haptools.transform.HaplotypesAncestry.transform   haptools/transform.py:91-149

The functions below are ALSO the bodies of this package's own mirror classes (haptools_b200/data/haplotypes.py,
haptools_b200/transform.py), so the mirror and a patched reference run the very same code.  They only rely on
the attribute contract the two share: ``Haplotypes.data / .type_ids / .index() / .log``, ``Haplotype.id / .chrom
/ .start / .variants``, ``Variant.id / .allele``, ``Genotypes.data / .samples / .variants / .index() / ._var_idx``
(and ``.ancestry / .ancestry_labels`` for the ancestry classes).

    import haptools
    from haptools_b200 import integrate
    integrate.install(haptools)      # the CLI, `haptools transform`, `haptools ld`, ... now run on the GPU
    integrate.uninstall(haptools)    # puts the numpy bodies back

There is no CPU fallback behind the patch: without a CUDA device the patched methods raise HaptoolsCudaError.
"""
from __future__ import annotations

import numpy as np


def _reject_empty(hap, genotypes):
    """The reference's single-haplotype form cannot broadcast its (1, 0) allele array against
    (n, 0, 2) and raises (haptools/data/haplotypes.py:499-511); only the plural form maps a
    haplotype without variants to all-True."""
    if len(hap.variants) == 0:
        n = genotypes.data.shape[0]
        raise ValueError(f"operands could not be broadcast together with shapes (1,0) ({n},0,2) ")


def haplotype_transform(hap, genotypes) -> np.ndarray:
    """Body of ``Haplotype.transform`` (haptools/data/haplotypes.py:463-511): ``(n, 2)`` bool."""
    from . import engine
    from .plan import plan_from_haplotypes

    # clip_bool=False: this form compares an INTEGER allele array (haplotypes.py:499-511), so with bool
    # genotypes an allele index >= 2 never matches (the plural form's dtype=bool cast does not apply)
    plan = plan_from_haplotypes([hap], genotypes, ancestry=False, who=f"haplotype '{hap.id}'", clip_bool=False)
    _reject_empty(hap, genotypes)
    return engine.transform_host(genotypes.data, plan)[:, 0]


def haplotype_ancestry_transform(hap, genotypes) -> np.ndarray:
    """Body of ``HaplotypeAncestry.transform`` (haptools/transform.py:31-68): ``(n, 2)`` bool."""
    from . import engine
    from .plan import plan_from_haplotypes

    plan, never = plan_from_haplotypes(
        [hap], genotypes, ancestry=True, who=f"haplotype '{hap.id}'", unknown_label="never"
    )
    _reject_empty(hap, genotypes)
    n, depth = genotypes.data.shape[0], genotypes.data.shape[2]
    if depth != 2:
        # haptools/transform.py:61 compares the whole last axis, so a leftover phase
        # column makes the reference's logical_and fail to broadcast
        raise ValueError(f"operands could not be broadcast together with shapes ({n},{depth}) ({n},2) ")
    if never[0]:
        return np.zeros((n, 2), dtype=np.bool_)
    return engine.transform_host(genotypes.data, plan, ancestry=genotypes.ancestry)[:, 0]


def haplotypes_transform(haps_obj, gts, hap_gts, ancestry: bool, genotypes_vcf):
    """Body of ``Haplotypes.transform`` (haptools/data/haplotypes.py:1339-1413) and, with ``ancestry=True``, of
    ``HaplotypesAncestry.transform`` (haptools/transform.py:91-149).  ``genotypes_vcf``: the GenotypesVCF class
    to instantiate when ``hap_gts`` is None (the reference's own when installed on the reference)."""
    from . import engine
    from .plan import plan_from_haplotypes

    haps_obj.index()
    haps = [haps_obj.data[hap] for hap in haps_obj.type_ids["H"]]
    if hap_gts is None:
        hap_gts = genotypes_vcf(fname=None, log=haps_obj.log)
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
    log = haps_obj.log
    plan = plan_from_haplotypes(haps, gts, ancestry=ancestry, log=log)
    log.info(f"Transforming genotypes for {len(haps)} haplotypes")
    log.debug(
        f"Allocating array with dtype {gts.data.dtype} and size "
        f"{(len(gts.samples), len(haps), 2)}"
    )
    log.debug("Computing haplotype genotypes. This may take a while")
    if ancestry and gts.data.shape[2] != 2:
        # haptools/transform.py:137 compares the whole last axis; a leftover phase column
        # makes the reference fail with a broadcast error -- same exception class here
        raise ValueError(
            f"could not broadcast input array from shape {(gts.data.shape[0], gts.data.shape[2])} "
            f"into shape {(gts.data.shape[0], 2)}"
        )
    hap_gts.data = engine.transform_host(gts.data, plan, ancestry=gts.ancestry if ancestry else None)
    return hap_gts


# ----------------------------------------------------------------------------------------
# install / uninstall on the reference package
# ----------------------------------------------------------------------------------------
_ORIGINAL = "_haptools_b200_original_transform"


def _targets(haptools_module):
    import importlib

    rdata = importlib.import_module(haptools_module.__name__ + ".data")
    rtransform = importlib.import_module(haptools_module.__name__ + ".transform")
    genotypes_vcf = rdata.GenotypesVCF

    def plural(self, gts, hap_gts=None):
        return haplotypes_transform(self, gts, hap_gts, False, genotypes_vcf)

    def plural_ancestry(self, gts, hap_gts=None):
        return haplotypes_transform(self, gts, hap_gts, True, genotypes_vcf)

    def single(self, genotypes):
        return haplotype_transform(self, genotypes)

    def single_ancestry(self, genotypes):
        return haplotype_ancestry_transform(self, genotypes)

    return [
        (rdata.Haplotype, single),
        (rdata.Haplotypes, plural),
        (rtransform.HaplotypeAncestry, single_ancestry),
        (rtransform.HaplotypesAncestry, plural_ancestry),
    ]


def install(haptools_module) -> None:
    """Replace the four transform bodies of the given ``haptools`` package (idempotent)."""
    for cls, fn in _targets(haptools_module):
        if _ORIGINAL in cls.__dict__:
            continue
        original = cls.__dict__["transform"]
        fn.__name__, fn.__qualname__ = "transform", f"{cls.__qualname__}.transform"
        fn.__doc__ = original.__doc__
        setattr(cls, _ORIGINAL, original)
        cls.transform = fn


def uninstall(haptools_module) -> None:
    """Put the reference's own numpy bodies back."""
    for cls, _ in _targets(haptools_module):
        if _ORIGINAL in cls.__dict__:
            cls.transform = cls.__dict__[_ORIGINAL]
            delattr(cls, _ORIGINAL)


def installed(haptools_module) -> bool:
    return all(_ORIGINAL in cls.__dict__ for cls, _ in _targets(haptools_module))
