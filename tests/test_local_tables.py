"""Host logic of the two-phase ("local") transform: the planner's locality tables
(haptools_b200/plan.py build_local_tables).  CPU only."""
import numpy as np
import pytest

from haptools_b200 import synth
from haptools_b200.plan import build_local_tables, plan_from_refs


def _check_tables(plan, lt, dmax, max_block):
    H, K = plan.n_haps, plan.n_refs
    order, pos = lt.order.astype(np.int64), lt.pos.astype(np.int64)
    assert sorted(order.tolist()) == list(range(H))
    np.testing.assert_array_equal(pos[order], np.arange(H))
    boff = lt.blk_hap_off.astype(np.int64)
    assert boff[0] == 0 and boff[-1] == H and np.all(np.diff(boff) >= 1) and np.all(np.diff(boff) <= max_block)
    assert len(lt.blk_key_lo) == len(lt.blk_key_cnt) == len(boff) - 1
    assert np.all(lt.blk_key_cnt <= dmax)
    s_off = lt.s_hap_off.astype(np.int64)
    assert s_off[0] == 0 and s_off[-1] == K
    off, key = plan.hap_off.astype(np.int64), plan.hap_key.astype(np.int64)
    reads = 0
    for b in range(lt.n_blocks):
        lo, cnt = int(lt.blk_key_lo[b]), int(lt.blk_key_cnt[b])
        for j in range(boff[b], boff[b + 1]):
            h = order[j]
            rel = lt.s_hap_key[s_off[j]:s_off[j + 1]].astype(np.int64)
            want = key[off[h]:off[h + 1]]
            if cnt:
                assert np.all(rel < cnt)
                np.testing.assert_array_equal(rel + lo, want)
            else:
                np.testing.assert_array_equal(rel, want)
                reads += len(want)
        reads += cnt
        assert lo + cnt <= max(plan.n_keys, 1)
    assert reads == lt.staged_reads


@pytest.mark.parametrize("dmax,max_block", [(288, 128), (16, 128), (40, 7), (1, 128)])
def test_local_tables_invariants(dmax, max_block):
    plan = synth.synth_plan(600, 900, seed=3, max_alleles=9)
    lt = build_local_tables(plan, dmax, max_block)
    _check_tables(plan, lt, dmax, max_block)


def test_local_tables_empty_and_wide_haplotypes():
    rng = np.random.default_rng(5)
    p, H = 500, 300
    k = rng.integers(0, 5, size=H)
    k[::17] = 0
    k[5] = 200  # spans far more than dmax keys: its block cannot be staged
    off = np.concatenate([[0], np.cumsum(k)])
    cols = np.concatenate([np.sort(rng.choice(p, size=int(x), replace=False)) if x > 4 else
                           np.sort(rng.integers(0, 6, size=int(x)) + rng.integers(0, p - 6)) for x in k] + [np.zeros(0, np.int64)])
    alle = rng.integers(0, 2, size=len(cols))
    plan = plan_from_refs(cols, alle, off, p)
    lt = build_local_tables(plan, 32, 16)
    _check_tables(plan, lt, 32, 16)
    assert (lt.blk_key_cnt == 0).any()
    # no haplotypes at all / no alleles at all
    for hap_off in (np.zeros(1, np.int64), np.zeros(8, np.int64)):
        pl = plan_from_refs(np.zeros(0, np.int64), np.zeros(0, np.int64), hap_off, p)
        t = build_local_tables(pl)
        _check_tables(pl, t, 288, 128)
        assert not t.worthwhile(pl.n_refs)


def test_worthwhile_follows_allele_sharing():
    dense = synth.synth_plan(5_000, 10_000, seed=2)   # K >> A: many haplotypes per variant
    sparse = synth.synth_plan(100_000, 1_000, seed=1)  # K ~ A: nothing to share
    assert build_local_tables(dense).worthwhile(dense.n_refs)
    assert not build_local_tables(sparse).worthwhile(sparse.n_refs)
