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
    to instantiate when ``hap_gts`` is None (the reference's own when installed on the reference).

    Everything that is O(H) Python -- the list of haplotypes, the identity check of the cached plan, the result's
    ``variants`` array: 80 ms at 100 k haplotypes -- runs on a helper thread while this one drives the device
    (``transform_host`` spends its time inside the library with the GIL released).  A cached plan is therefore used
    OPTIMISTICALLY: the device starts on it at once, and if the helper finds that something the plan was built from
    has been replaced since, the result is dropped and the call starts over with a fresh plan."""
    haps_obj.index()
    if hap_gts is None:
        hap_gts = genotypes_vcf(fname=None, log=haps_obj.log)
    hap_gts.samples = gts.samples
    n_haps = len(haps_obj.type_ids["H"])
    entry = _cache_probe(haps_obj, gts, ancestry, n_haps)
    if entry is not None:
        # the helper starts when the engine has queued what it can queue without waiting (engine.transform_host
        # `on_queued`): from then on this thread only waits inside the library, and the helper has the interpreter
        job = _Background(_list_check_variants, n_haps, haps_obj, entry, ancestry, hap_gts.variants.dtype, defer=True)
        plan, n_distinct = entry[2], entry[3]
        with _ShortSwitchInterval(job.threaded):
            try:
                data = _logged_transform(haps_obj.log, gts, plan, n_haps, ancestry, n_distinct, on_queued=job.start)
            except BaseException:
                hap_gts.variants = job.result()[2]
                raise
            haps, fresh, variants = job.result()
        if fresh:
            hap_gts.variants, hap_gts.data = variants, data
            return hap_gts
        del data  # built from a plan that no longer describes these haplotypes
        clear_plan_cache(haps_obj)
    else:
        haps = _haplotype_list(haps_obj)
    job = _Background(_result_variants, len(haps), haps_obj, haps, hap_gts.variants.dtype)
    with _ShortSwitchInterval(job.threaded):
        try:
            # the planner emits the first two of the five log lines at the steps they belong to
            plan = _planned(haps_obj, haps, gts, ancestry, haps_obj.log)
            data = _logged_transform(haps_obj.log, gts, plan, len(haps), ancestry, None)
        finally:
            hap_gts.variants = job.result()  # the reference sets `variants` before it can fail (haplotypes.py:1369)
    hap_gts.data = data
    return hap_gts


def _haplotype_list(haps_obj) -> list:
    ids = haps_obj.type_ids["H"]
    if len(ids) == len(haps_obj.data) and list(haps_obj.data) == ids:
        return list(haps_obj.data.values())  # only haplotypes, in index order: no 100 k dict look-ups
    return list(map(haps_obj.data.__getitem__, ids))


def _logged_transform(log, gts, plan, n_haps, ancestry, n_distinct, on_queued=None):
    """The five log lines of haplotypes.py:1387-1410 / transform.py:122-143, verbatim and in order, around the
    library call (``n_distinct`` None: the planner has just emitted the first two)."""
    from . import engine

    if n_distinct is not None:
        log.debug(f"Copying genotypes for {n_distinct} distinct alleles")
        log.debug("Creating array denoting alt allele status")
    log.info(f"Transforming genotypes for {n_haps} haplotypes")
    log.debug(
        f"Allocating array with dtype {gts.data.dtype} and size "
        f"{(len(gts.samples), n_haps, 2)}"
    )
    log.debug("Computing haplotype genotypes. This may take a while")
    if ancestry and gts.data.shape[2] != 2:
        # haptools/transform.py:137 compares the whole last axis; a leftover phase column
        # makes the reference fail with a broadcast error -- same exception class here
        raise ValueError(
            f"could not broadcast input array from shape {(gts.data.shape[0], gts.data.shape[2])} "
            f"into shape {(gts.data.shape[0], 2)}"
        )
    kw = {"on_queued": on_queued} if on_queued is not None else {}
    return engine.transform_host(gts.data, plan, ancestry=gts.ancestry if ancestry else None, **kw)


class _Background:
    """fn(*args) on a helper thread (inline when ``size`` is small); result() joins and re-raises.  ``defer``: the
    work begins at start() (or, if nobody called that, inside result())."""

    def __init__(self, fn, size, *args, defer: bool = False):
        self._out = self._err = self._thread = None
        self._fn, self._args, self._started = fn, args, False
        self.threaded = size >= 4096
        if not defer:
            self.start()

    def _run(self):
        try:
            self._out = self._fn(*self._args)
        except BaseException as e:  # handed to the caller's thread
            self._err = e

    def start(self):
        import threading

        if self._started:
            return
        self._started = True
        if not self.threaded:
            self._run()
        else:
            self._thread = threading.Thread(target=self._run, name="haptools_b200-helper", daemon=True)
            self._thread.start()

    def result(self):
        if not self._started:  # never started (the engine failed before it got there): do the work here
            self.threaded = False
            self.start()
        if self._thread is not None:
            self._thread.join()
            self._thread = None
        if self._err is not None:
            raise self._err
        return self._out


class _ShortSwitchInterval:
    """While a helper thread runs Python next to the thread that drives the device, the driving thread has to win
    the GIL back after every library call; with the interpreter's default switch interval (5 ms) each of the
    ~100 calls of a transform could wait that long.  0.2 ms for the duration of the call, then the old value."""

    def __init__(self, active: bool):
        self._active, self._old = active, None

    def __enter__(self):
        import sys

        if self._active:
            self._old = sys.getswitchinterval()
            sys.setswitchinterval(min(self._old, 2e-4))
        return self

    def __exit__(self, *exc):
        import sys

        if self._old is not None:
            sys.setswitchinterval(self._old)
        return False


# ----------------------------------------------------------------------------------------
# plan cache
# ----------------------------------------------------------------------------------------
# The planner walks every Variant object of every haplotype (the reference's own O(K) dict loop,
# haplotypes.py:1375-1386): 0.5 s for 100 k haplotypes / 1.1 M alleles, against 40 ms for transforming 16 k
# samples.  Callers that transform several genotype sets (or sample batches) with the SAME haplotype set -- what
# `haptools transform` on chunked input and `haptools ld` do -- get the plan of the previous call back when
# nothing it was built from has been replaced: the `data` dict and `type_ids` of the Haplotypes object, the
# `variants` TUPLE of every haplotype (tuples are immutable; the check is by identity, O(H)), the `variants`
# array and ID index of the genotypes, the dtype class of `data` and the ancestry label dict.  The O(1) part of
# that stamp decides whether the entry is tried at all (`_cache_probe`); the O(H) part runs on the helper thread
# NEXT TO the transform that already uses the plan (`_list_check_variants`), and a transform whose plan turns out
# stale is thrown away and repeated with a fresh plan (tests/test_plan_cache.py).  What the check
# cannot see is an IN-PLACE edit of a Variant's fields or of the genotypes' `variants` array between two calls;
# code that does that sets HAPTOOLS_B200_PLAN_CACHE=0 (or calls clear_plan_cache(haps)).
_CACHE_ATTR = "_haptools_b200_plan_cache"


def clear_plan_cache(haps_obj) -> None:
    haps_obj.__dict__.pop(_CACHE_ATTR, None)
    haps_obj.__dict__.pop(_CACHE_ATTR + "_variants", None)


def _result_variants(haps_obj, haps, dtype):
    """``hap_gts.variants`` (haptools/data/haplotypes.py:1369-1372), built column by column instead of a tuple
    per haplotype.  The string -> U50 conversions of the previous call are kept (as plain column arrays) and
    reused when the (id, chrom, start) of every haplotype are what they were; the result is a fresh array
    every time, like the reference's."""
    fields = ([hap.id for hap in haps], [hap.chrom for hap in haps], [hap.start for hap in haps])
    entry = haps_obj.__dict__.get(_CACHE_ATTR + "_variants")
    if entry is None or entry[0] != fields or entry[1] != dtype:
        cols = (np.asarray(fields[0], dtype=dtype["id"]), np.asarray(fields[1], dtype=dtype["chrom"]),
                np.asarray(fields[2], dtype=dtype["pos"])) if len(haps) else None
        entry = (fields, dtype, cols)
        if _cache_enabled():
            haps_obj.__dict__[_CACHE_ATTR + "_variants"] = entry
    variants = np.empty(len(haps), dtype=dtype)
    if len(haps):
        variants["id"], variants["chrom"], variants["pos"] = entry[2]
        variants["alleles"].fill(("A", "T"))
    return variants


def _cache_enabled() -> bool:
    import os

    return os.environ.get("HAPTOOLS_B200_PLAN_CACHE", "1").strip() not in ("0", "off", "no")


def _stamp(haps_obj, gts, ancestry):
    """The O(1) part of what a plan was built from (compared by identity up to index 2, by value after)."""
    return (
        haps_obj.data, gts.variants, gts._var_idx, gts.data.dtype == np.bool_, gts.data.shape[1], bool(ancestry),
        tuple(sorted(gts.ancestry_labels.items())) if ancestry else None,
    )


def _cache_probe(haps_obj, gts, ancestry, n_haps):
    """The cache entry if its O(1) stamp matches this call (the O(H) part is `_list_check_variants`), else None."""
    if not _cache_enabled():
        return None
    entry = haps_obj.__dict__.get(_CACHE_ATTR)
    if entry is None:
        return None
    gts.index(samples=False, variants=True)  # raises on duplicate IDs like the planner would
    old, new = entry[0], _stamp(haps_obj, gts, ancestry)
    if old[0] is new[0] and old[1] is new[1] and old[2] is new[2] and old[3:] == new[3:] and len(entry[1]) == n_haps:
        return entry
    return None


def _list_check_variants(haps_obj, entry, ancestry, dtype):
    """Helper-thread half of a call that found a cache entry: (haplotype list, entry still valid, `variants`)."""
    from operator import is_

    haps = _haplotype_list(haps_obj)
    tuples = [h.variants for h in haps]
    fresh = len(tuples) == len(entry[1]) and all(map(is_, entry[1], tuples))
    if fresh and ancestry:
        fresh = [h.ancestry for h in haps] == entry[4]
    return haps, fresh, _result_variants(haps_obj, haps, dtype)


def _planned(haps_obj, haps, gts, ancestry, log):
    """A fresh plan, remembered for the next call (unless the cache is switched off)."""
    from .plan import plan_from_haplotypes

    if not _cache_enabled():
        return plan_from_haplotypes(haps, gts, ancestry=ancestry, log=log)
    gts.index(samples=False, variants=True)  # raises on duplicate IDs like the planner would
    seen = []
    plan = plan_from_haplotypes(haps, gts, ancestry=ancestry, log=log, n_distinct_out=seen)
    # the entry keeps the tuples (and the two arrays) alive, so an `is` match can never be a recycled address
    haps_obj.__dict__[_CACHE_ATTR] = (
        _stamp(haps_obj, gts, ancestry), [h.variants for h in haps], plan, seen[0],
        [h.ancestry for h in haps] if ancestry else None,
    )
    return plan


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
