"""
TEST INFRASTRUCTURE ONLY -- never imported by the product path (haptools_b200/).

Imports the UNMODIFIED reference (CAST-genomics/haptools, mounted read-only at
/root/reference) in *this* container so that the numpy restatement in
``oracle/transform_oracle.py`` can be validated against it and so that golden
vectors can be generated (``tests/golden/make_golden.py``).

The reference imports ``pysam``, ``cyvcf2``, ``pgenlib`` and ``matplotlib`` at module
top (haptools/data/genotypes.py:10-14, haptools/data/haplotypes.py:10,
haptools/transform.py:8-10); none is installed here and the transform hot path never
touches them, so empty stub modules are registered before the import.  No reference
file is modified or copied.

/root/reference does not exist on the GPU box: nothing that runs there may import
this module (``available()`` returns False there).
"""
from __future__ import annotations

import sys
import types
from pathlib import Path

REFERENCE_ROOT = Path("/root/reference")

_STUBS = {
    "pysam": ["TabixFile", "VariantFile", "tabix_index"],
    "cyvcf2": ["VCF", "Variant"],
    "pgenlib": ["PgenReader", "PvarReader", "PgenWriter"],
    "matplotlib": [],
}


def available() -> bool:
    return (REFERENCE_ROOT / "haptools" / "data" / "haplotypes.py").exists()


def load():
    """Return a namespace holding the reference's hot-path classes."""
    if not available():
        raise RuntimeError("the reference checkout is not present on this machine")
    for name, attrs in _STUBS.items():
        if name not in sys.modules:
            mod = types.ModuleType(name)
            for attr in attrs:
                setattr(mod, attr, type(attr, (), {}))
            sys.modules[name] = mod
    if str(REFERENCE_ROOT) not in sys.path:
        sys.path.insert(0, str(REFERENCE_ROOT))
    import haptools.data as rdata  # noqa: E402  (the reference's package)
    import haptools.transform as rtransform  # noqa: E402

    ns = types.SimpleNamespace(
        data=rdata,
        transform=rtransform,
        Variant=rdata.Variant,
        Haplotype=rdata.Haplotype,
        Haplotypes=rdata.Haplotypes,
        Genotypes=rdata.Genotypes,
        GenotypesVCF=rdata.GenotypesVCF,
        HaplotypeAncestry=rtransform.HaplotypeAncestry,
        HaplotypesAncestry=rtransform.HaplotypesAncestry,
        GenotypesAncestry=rtransform.GenotypesAncestry,
    )
    return ns
