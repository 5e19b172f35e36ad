"""
The executable reference-side binding (haptools_b200/integrate.py).

CPU part (build container only, skipped where /root/reference is absent): install() on the LIVE, unmodified
reference, with engine.transform_host replaced by a numpy stand-in that executes the PLAN the planner produced
(so everything above the C-ABI -- planner, key folding, result container, log lines, exceptions -- is exercised
on the reference's real objects), then
  * the reference's OWN unit tests for the path (tests/test_data.py:1616-1695, tests/test_transform.py:167-212),
  * randomised differential runs against the reference's original numpy bodies, log records included.
GPU part: the same install() with the real engine, on the mirror's objects standing in for the package.
"""
import logging
import sys
import types

import numpy as np
import pytest

import _cases
from oracle import ref_harness

needs_ref = pytest.mark.skipif(not ref_harness.available(), reason="the reference checkout is not on this machine")


def plan_executor(calls):
    """numpy stand-in for engine.transform_host: executes the plan's keys and CSR literally."""

    def run(gt, plan, ancestry=None, **kw):
        calls.append(plan)
        key_col = plan.key_col().astype(np.int64)
        eq = gt[:, key_col, :2] == plan.key_allele[np.newaxis, :, np.newaxis]
        if plan.has_ancestry:
            eq &= ancestry[:, key_col, :2] == plan.key_label[np.newaxis, :, np.newaxis]
        out = np.empty((gt.shape[0], plan.n_haps, 2), dtype=np.bool_)
        off = plan.hap_off.astype(np.int64)
        for h in range(plan.n_haps):
            out[:, h] = eq[:, plan.hap_key[off[h]:off[h + 1]].astype(np.int64)].all(axis=1)
        return out

    return run


@pytest.fixture
def patched_reference(monkeypatch):
    ns = ref_harness.load()
    import haptools  # the reference's package

    from haptools_b200 import engine, integrate

    calls = []
    monkeypatch.setattr(engine, "transform_host", plan_executor(calls))
    integrate.install(haptools)
    assert integrate.installed(haptools)
    try:
        yield ns, calls
    finally:
        integrate.uninstall(haptools)
        assert not integrate.installed(haptools)


@needs_ref
def test_reference_own_tests_pass_through_the_binding(patched_reference):
    ns, calls = patched_reference
    from tests.test_data import TestHaplotypes  # the reference's test module (regular package at /root/reference)
    from tests.test_transform import TestHaplotypesAncestry

    t = TestHaplotypes()
    for name in ("test_hap_transform", "test_hap_transform_multiallelic", "test_haps_transform",
                 "test_haps_transform_multiallelic"):
        before = len(calls)
        getattr(t, name)()
        assert len(calls) > before, f"{name} did not go through the installed binding"
    ta = TestHaplotypesAncestry()
    for name in ("test_hap_transform", "test_haps_transform"):
        before = len(calls)
        getattr(ta, name)()
        assert len(calls) > before, f"{name} did not go through the installed binding"


class _Capture(logging.Handler):
    def __init__(self):
        super().__init__(level=logging.DEBUG)
        self.records = []

    def emit(self, record):
        self.records.append((record.levelname, record.getMessage()))


def _logger(name):
    log = logging.getLogger(name)
    log.setLevel(logging.DEBUG)
    log.propagate = False
    cap = _Capture()
    log.handlers = [cap]
    return log, cap


@needs_ref
@pytest.mark.parametrize("seed", range(10))
def test_patched_reference_equals_original_numpy_bodies(patched_reference, seed):
    ns, calls = patched_reference
    from haptools_b200 import integrate

    rng = np.random.default_rng(4400 + seed)
    kind = "ancestry" if seed % 2 else "plain"
    case = _cases.random_case(rng, n=int(rng.integers(1, 60)), p=int(rng.integers(2, 25)), n_haps=int(rng.integers(1, 30)),
                              kind=kind, max_alleles=2 if kind == "ancestry" else int(rng.choice([2, 3])),
                              missing_rate=float(rng.choice([0, 0.05])), as_bool=bool(seed % 3 == 0 and kind == "plain"),
                              empty_hap=bool(seed % 4 == 0), layout=str(rng.choice(["C", "F"])))
    if kind == "ancestry" and seed % 5 == 1:
        case["ancestry_labels"] = {"YRI": 7, **case["ancestry_labels"]}
    haps, gts = _cases.build_objects(case, ns)
    cls = type(haps)
    original = cls.__dict__[integrate._ORIGINAL]
    log_new, cap_new = _logger(f"integrate_new_{seed}")
    log_old, cap_old = _logger(f"integrate_old_{seed}")
    haps.log = log_new
    got = haps.transform(gts)
    haps.log = log_old
    want = original(haps, gts)
    assert type(got) is type(want) is ns.GenotypesVCF
    assert got.data.dtype == want.data.dtype == np.bool_
    np.testing.assert_array_equal(got.data, want.data)
    assert got.samples == want.samples
    assert got.variants.dtype == want.variants.dtype and list(got.variants) == list(want.variants)
    assert cap_new.records == cap_old.records  # the five log lines, verbatim and in order
    assert len(cap_new.records) == 5
    # single forms
    for hap in haps.data.values():
        one_orig = type(hap).__dict__[integrate._ORIGINAL]
        if len(hap.variants) == 0:
            with pytest.raises(ValueError):
                hap.transform(gts)
            with pytest.raises(ValueError):
                one_orig(hap, gts)
            continue
        np.testing.assert_array_equal(hap.transform(gts), one_orig(hap, gts))


@needs_ref
def test_single_form_with_bool_genotypes_and_a_multiallelic_allele(patched_reference):
    """haplotypes.py:499-511 keeps an INTEGER allele array: with bool genotypes an ALT index >= 2 never matches,
    while the plural form (dtype=bool cast, :1398) treats it like 1."""
    ns, _ = patched_reference
    from haptools_b200 import integrate

    rng = np.random.default_rng(3)
    case = _cases.random_case(rng, n=12, p=6, n_haps=5, max_alleles=3, as_bool=True)
    for h in case["haps"]:  # make every haplotype ask for the LAST allele of its first variant
        j = case["var_id"].index(h["variants"][0]["id"])
        h["variants"][0]["allele"] = case["var_alleles"][j][-1]
    haps, gts = _cases.build_objects(case, ns)
    for hap in haps.data.values():
        np.testing.assert_array_equal(hap.transform(gts), type(hap).__dict__[integrate._ORIGINAL](hap, gts))
    np.testing.assert_array_equal(haps.transform(gts).data, type(haps).__dict__[integrate._ORIGINAL](haps, gts).data)


@needs_ref
def test_errors_are_the_references(patched_reference):
    ns, _ = patched_reference
    from haptools_b200 import integrate

    rng = np.random.default_rng(8)
    case = _cases.random_case(rng, n=5, p=4, n_haps=3)
    case["haps"][1]["variants"][0]["allele"] = "NOPE"
    haps, gts = _cases.build_objects(case, ns)
    for fn in (haps.transform, lambda g: type(haps).__dict__[integrate._ORIGINAL](haps, g)):
        with pytest.raises(ValueError, match="Some alleles were not present in the genotypes"):
            fn(gts)
    case = _cases.random_case(rng, n=5, p=4, n_haps=3, kind="ancestry")
    case["haps"][0]["ancestry"] = "UNKNOWN"
    haps, gts = _cases.build_objects(case, ns)
    for fn in (haps.transform, lambda g: type(haps).__dict__[integrate._ORIGINAL](haps, g)):
        with pytest.raises(OverflowError):
            fn(gts)


def test_install_on_a_package_shaped_like_haptools(monkeypatch):
    """install()/uninstall() mechanics without the reference: any package exposing `.data` and `.transform`
    submodules with the four classes (here: this repo's mirror, wrapped as a fake package)."""
    from haptools_b200 import data as hdata
    from haptools_b200 import engine, integrate
    from haptools_b200 import transform as htransform

    pkg = types.ModuleType("fake_haptools")
    d, t = types.ModuleType("fake_haptools.data"), types.ModuleType("fake_haptools.transform")

    class Haplotype(hdata.Haplotype):
        def transform(self, genotypes):
            return "numpy body"

    class Haplotypes(hdata.Haplotypes):
        def transform(self, gts, hap_gts=None):
            return "numpy body"

    class HaplotypeAncestry(htransform.HaplotypeAncestry):
        def transform(self, genotypes):
            return "numpy body"

    class HaplotypesAncestry(htransform.HaplotypesAncestry):
        def transform(self, gts, hap_gts=None):
            return "numpy body"

    d.Haplotype, d.Haplotypes, d.GenotypesVCF = Haplotype, Haplotypes, hdata.GenotypesVCF
    t.HaplotypeAncestry, t.HaplotypesAncestry = HaplotypeAncestry, HaplotypesAncestry
    pkg.data, pkg.transform = d, t
    for m in (pkg, d, t):
        monkeypatch.setitem(sys.modules, m.__name__, m)
    calls = []
    monkeypatch.setattr(engine, "transform_host", plan_executor(calls))
    integrate.install(pkg)
    integrate.install(pkg)  # idempotent
    rng = np.random.default_rng(1)
    case = _cases.random_case(rng, n=9, p=5, n_haps=4)
    ns = types.SimpleNamespace(Variant=hdata.Variant, Haplotype=Haplotype, Haplotypes=Haplotypes, GenotypesVCF=hdata.GenotypesVCF)
    haps, gts = _cases.build_objects(case, ns)
    out = haps.transform(gts)
    assert isinstance(out, hdata.GenotypesVCF) and out.data.shape == (9, 4, 2) and len(calls) == 1
    assert Haplotypes.transform.__qualname__.endswith("Haplotypes.transform")
    integrate.uninstall(pkg)
    assert haps.transform(gts) == "numpy body"


@pytest.mark.gpu
def test_installed_binding_runs_the_cuda_path(monkeypatch):
    """install() on a stand-in package (numpy bodies restated from the oracle) with the REAL engine: the patched
    classes return what the numpy bodies returned, and the planner's work reached the device library."""
    from haptools_b200 import data as hdata
    from haptools_b200 import engine, integrate
    from haptools_b200 import transform as htransform
    from oracle import transform_oracle as orc

    pkg = types.ModuleType("fake_haptools_gpu")
    d, t = types.ModuleType("fake_haptools_gpu.data"), types.ModuleType("fake_haptools_gpu.transform")

    class Haplotypes(hdata.Haplotypes):
        def transform(self, gts, hap_gts=None):  # the "numpy body"
            self.index()
            out = hdata.GenotypesVCF(fname=None)
            out.data = orc.haplotypes_transform([self.data[k] for k in self.type_ids["H"]], gts)
            return out

    class HaplotypesAncestry(htransform.HaplotypesAncestry):
        def transform(self, gts, hap_gts=None):
            self.index()
            out = hdata.GenotypesVCF(fname=None)
            out.data = orc.haplotypes_ancestry_transform([self.data[k] for k in self.type_ids["H"]], gts)
            return out

    d.Haplotype, d.Haplotypes, d.GenotypesVCF = type("Haplotype", (hdata.Haplotype,), {"transform": lambda s, g: None}), Haplotypes, hdata.GenotypesVCF
    t.HaplotypeAncestry = type("HaplotypeAncestry", (htransform.HaplotypeAncestry,), {"transform": lambda s, g: None})
    t.HaplotypesAncestry = HaplotypesAncestry
    pkg.data, pkg.transform = d, t
    for m in (pkg, d, t):
        monkeypatch.setitem(sys.modules, m.__name__, m)
    seen = []
    real = engine.transform_host
    monkeypatch.setattr(engine, "transform_host", lambda *a, **kw: (seen.append(1), real(*a, **kw))[1])
    rng = np.random.default_rng(2)
    for kind, cls, hcls in (("plain", Haplotypes, d.Haplotype), ("ancestry", HaplotypesAncestry, t.HaplotypeAncestry)):
        case = _cases.random_case(rng, n=1500, p=30, n_haps=40, kind=kind, missing_rate=0.02, empty_hap=True)
        ns = types.SimpleNamespace(Variant=hdata.Variant, Haplotype=hcls, Haplotypes=cls, GenotypesVCF=hdata.GenotypesVCF,
                                   HaplotypeAncestry=hcls, HaplotypesAncestry=cls, GenotypesAncestry=htransform.GenotypesAncestry)
        haps, gts = _cases.build_objects(case, ns)
        want = haps.transform(gts).data
        integrate.install(pkg)
        try:
            n_before = len(seen)
            got = haps.transform(gts)
            assert len(seen) == n_before + 1
            np.testing.assert_array_equal(got.data, want)
        finally:
            integrate.uninstall(pkg)
