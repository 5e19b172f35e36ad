"""
The class API's plan cache (haptools_b200/integrate.py): a cached plan is used optimistically -- the device starts
on it while a helper thread checks, haplotype by haplotype, that nothing the plan was built from was replaced --
and a stale plan costs one wasted transform, never a wrong result.  CPU tests: engine.transform_host is replaced by
the numpy stand-in of tests/test_integrate.py that executes the plan literally; the expected arrays come from the
oracle's restatement of haptools/data/haplotypes.py:1375-1412 / haptools/transform.py:103-148.
"""
import logging
import sys
import types

import numpy as np
import pytest

import _cases
from oracle import transform_oracle as orc
from test_integrate import plan_executor


def mirror_ns():
    from haptools_b200 import data as hdata
    from haptools_b200 import transform as htransform

    return types.SimpleNamespace(
        Variant=hdata.Variant, Haplotype=hdata.Haplotype, Haplotypes=hdata.Haplotypes, GenotypesVCF=hdata.GenotypesVCF,
        HaplotypeAncestry=htransform.HaplotypeAncestry, HaplotypesAncestry=htransform.HaplotypesAncestry,
        GenotypesAncestry=htransform.GenotypesAncestry,
    )


def expected(case):
    """The oracle's object-level restatement on freshly built objects of the (possibly edited) case."""
    haps, gts = _cases.build_objects(case, mirror_ns())
    fn = orc.haplotypes_ancestry_transform if case["kind"] == "ancestry" else orc.haplotypes_transform
    return fn(list(haps.data.values()), gts)


@pytest.fixture
def executor(monkeypatch):
    from haptools_b200 import engine

    calls = []
    monkeypatch.setattr(engine, "transform_host", plan_executor(calls))
    return calls


# 5,000 haplotypes: above the size at which the O(H) bookkeeping moves to the helper thread
@pytest.mark.parametrize("n_haps,kind", [(6, "plain"), (5000, "plain"), (7, "ancestry"), (4500, "ancestry")])
def test_hit_then_stale_then_hit(executor, n_haps, kind):
    rng = np.random.default_rng(n_haps)
    case = _cases.random_case(rng, n=5, p=40, n_haps=n_haps, kind=kind, max_hap_len=4)
    ns = mirror_ns()
    haps, gts = _cases.build_objects(case, ns)
    want = expected(case)
    interval = sys.getswitchinterval()

    first = haps.transform(gts)
    assert np.array_equal(first.data, want) and len(executor) == 1
    second = haps.transform(gts)
    assert np.array_equal(second.data, want) and len(executor) == 2
    assert executor[1] is executor[0], "the second call must reuse the plan"
    assert second.variants is not first.variants and np.array_equal(second.variants["id"], first.variants["id"])

    # replace ONE haplotype's variants tuple (what editing a haplotype means: tuples are immutable)
    target = n_haps // 2
    hid = f"H{target}"
    old = case["haps"][target]["variants"]
    donor = case["haps"][(target + 1) % n_haps]["variants"]
    assert donor != old or n_haps < 3
    case["haps"][target]["variants"] = donor
    haps.data[hid].variants = tuple(
        ns.Variant(start=v["start"], end=v["end"], id=v["id"], allele=v["allele"]) for v in donor
    )
    want2 = expected(case)
    third = haps.transform(gts)
    assert np.array_equal(third.data, want2), "a stale plan must never reach the caller"
    # the optimistic run on the old plan, then the run on the fresh one
    assert len(executor) == 4 and executor[2] is executor[0] and executor[3] is not executor[0]
    fourth = haps.transform(gts)
    assert np.array_equal(fourth.data, want2) and len(executor) == 5 and executor[4] is executor[3]
    assert sys.getswitchinterval() == interval


def test_replaced_genotype_metadata_is_a_miss(executor):
    rng = np.random.default_rng(3)
    case = _cases.random_case(rng, n=4, p=12, n_haps=5)
    haps, gts = _cases.build_objects(case, mirror_ns())
    haps.transform(gts)
    haps.transform(gts)
    assert executor[1] is executor[0]
    # another Genotypes object with the columns in another order: the plan's columns no longer apply
    perm = np.arange(12)[::-1]
    case2 = dict(case, gt=case["gt"][:, perm], var_id=[case["var_id"][j] for j in perm],
                 var_pos=[case["var_pos"][j] for j in perm], var_alleles=[case["var_alleles"][j] for j in perm])
    _, gts2 = _cases.build_objects(case2, mirror_ns())
    out = haps.transform(gts2)
    assert executor[2] is not executor[0] and len(executor) == 3, "no optimistic run when the O(1) stamp differs"
    assert np.array_equal(out.data, expected(case))


def test_ancestry_label_of_a_haplotype_changed(executor):
    rng = np.random.default_rng(11)
    case = _cases.random_case(rng, n=6, p=10, n_haps=5000, kind="ancestry", max_hap_len=3)
    haps, gts = _cases.build_objects(case, mirror_ns())
    haps.transform(gts)
    was = case["haps"][17]["ancestry"]
    now = next(pop for pop in case["ancestry_labels"] if pop != was)
    case["haps"][17]["ancestry"] = now
    haps.data["H17"].ancestry = now
    out = haps.transform(gts)
    assert np.array_equal(out.data, expected(case))


def test_five_log_lines_on_a_hit(executor, caplog):
    rng = np.random.default_rng(5)
    case = _cases.random_case(rng, n=4, p=9, n_haps=4)
    haps, gts = _cases.build_objects(case, mirror_ns())
    haps.log = logging.getLogger("plan_cache_test")
    with caplog.at_level(logging.DEBUG, logger="plan_cache_test"):
        haps.transform(gts)
        miss = [r.getMessage() for r in caplog.records]
        caplog.clear()
        haps.transform(gts)
        hit = [r.getMessage() for r in caplog.records]
    assert hit == miss and len(hit) == 5
    assert hit[0].startswith("Copying genotypes for ") and hit[2] == "Transforming genotypes for 4 haplotypes"


def test_engine_error_still_sets_variants(monkeypatch):
    from haptools_b200 import engine

    rng = np.random.default_rng(6)
    case = _cases.random_case(rng, n=4, p=9, n_haps=4)
    ns = mirror_ns()
    haps, gts = _cases.build_objects(case, ns)
    monkeypatch.setattr(engine, "transform_host", plan_executor([]))
    haps.transform(gts)

    def boom(*a, **kw):
        raise RuntimeError("device lost")

    monkeypatch.setattr(engine, "transform_host", boom)
    hap_gts = ns.GenotypesVCF(fname=None)
    with pytest.raises(RuntimeError, match="device lost"):
        haps.transform(gts, hap_gts)
    # the reference fills `variants` before the arithmetic (haptools/data/haplotypes.py:1369)
    assert list(hap_gts.variants["id"]) == [h["id"] for h in case["haps"]]
