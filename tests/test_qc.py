"""
Genotype QC.  CPU: the numpy restatement (oracle/qc_oracle.py) against the scenarios recorded
from the unmodified reference (tests/golden/qc_*.npz, make_golden_qc.py).  GPU: the mirror classes -- whose
checks run as ONE device pass, ht_qc_u8 -- against the same scenarios in every memory layout, and the raw
kernel outputs against numpy on random matrices.
"""
import json
import logging

import numpy as np
import pytest

import _cases
from oracle import qc_oracle

FIXTURES = sorted(_cases.GOLDEN_DIR.glob("qc_*.npz"))


def _load(path):
    z = np.load(path, allow_pickle=False)
    meta = json.loads(str(z["meta"]))
    return z, meta


def _variants(p, dtype):
    return np.array([(f"v{j}", "1" if j % 3 else "2", 100 + 7 * j, ("A", "C", "G")) for j in range(p)], dtype=dtype)


def _check_outcome(z, meta, outcomes, rets, final_data, final_samples, final_ids, log, final_ancestry=None):
    want = meta["outcomes"]
    assert len(outcomes) == len(want)
    for i, (got, exp) in enumerate(zip(outcomes, want)):
        if "raises" in exp:
            assert got.get("raises") == exp["raises"]
        else:
            assert "raises" not in got
            if exp["ret"]:
                np.testing.assert_array_equal(rets[i], z[f"ret{i}"])  # MAF: same float64 arithmetic -> same bits
            else:
                assert rets[i] is None
    assert final_data.dtype == z["final_data"].dtype
    np.testing.assert_array_equal(final_data, z["final_data"])
    assert list(final_samples) == meta["final_samples"]
    assert list(final_ids) == list(z["final_ids"])
    assert [tuple(r) for r in log] == [tuple(r) for r in meta["log"]]
    if final_ancestry is not None:
        np.testing.assert_array_equal(final_ancestry, z["final_ancestry"])


@pytest.mark.parametrize("path", FIXTURES, ids=lambda p: p.stem)
def test_oracle_matches_reference_scenarios(path):
    z, meta = _load(path)
    data = z["data"]
    dtype = [("id", "U50"), ("chrom", "U10"), ("pos", np.uint32), ("alleles", object)]
    state = {"data": data.copy(), "samples": tuple(f"S{i}" for i in range(data.shape[0])),
             "variants": _variants(data.shape[1], dtype), "prephased": meta["prephased"], "log": [],
             "ancestry": z["ancestry"].copy() if meta["kind"] == "ancestry" else None}
    outcomes = qc_oracle.run_scenario(state, [(m, kw) for m, kw in meta["calls"]])
    rets = [o.get("ret") for o in outcomes]
    _check_outcome(z, meta, outcomes, rets, state["data"], state["samples"], state["variants"]["id"], state["log"],
                   state["ancestry"] if meta["kind"] == "ancestry" else None)


class _Capture(logging.Handler):
    def __init__(self):
        super().__init__(level=logging.DEBUG)
        self.records = []

    def emit(self, record):
        self.records.append((record.levelname, record.getMessage()))


def _layouts3(d):
    """Memory layouts the readers produce for an (n, p, depth) matrix."""
    out = {"C": np.ascontiguousarray(d), "F": np.ascontiguousarray(d.transpose(1, 0, 2)).transpose(1, 0, 2)}
    if d.shape[2] == 2 and d.dtype == np.uint8:
        c3 = np.zeros(d.shape[:2] + (3,), dtype=d.dtype)
        c3[:, :, :2] = d
        out["C3view"] = c3[:, :, :2]
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("path", FIXTURES, ids=lambda p: p.stem)
def test_mirror_classes_match_reference_scenarios(path):
    from haptools_b200 import data as hdata
    from haptools_b200 import transform as htransform

    z, meta = _load(path)
    for lay, arr in _layouts3(z["data"]).items():
        log = logging.getLogger(f"qc_test_{path.stem}_{lay}")
        log.setLevel(logging.DEBUG)
        log.propagate = False
        cap = _Capture()
        log.handlers = [cap]
        anc = meta["kind"] == "ancestry"
        gts = (htransform.GenotypesAncestry if anc else hdata.GenotypesVCF)(fname=None, log=log)
        gts.data = arr.copy(order="K") if lay != "C3view" else arr
        gts.samples = tuple(f"S{i}" for i in range(arr.shape[0]))
        gts.variants = _variants(arr.shape[1], gts.variants.dtype)
        gts._prephased = meta["prephased"]
        if anc:
            gts.ancestry = z["ancestry"].copy()
        outcomes, rets = [], []
        for method, kwargs in meta["calls"]:
            try:
                rets.append(getattr(gts, method)(**kwargs))
                outcomes.append({"ret": rets[-1] is not None})
            except ValueError as e:
                rets.append(None)
                outcomes.append({"raises": str(e)})
                break
        _check_outcome(z, meta, outcomes, rets, gts.data, gts.samples, gts.variants["id"], cap.records,
                       gts.ancestry if anc else None)


@pytest.mark.gpu
@pytest.mark.parametrize("n,p,depth", [(1, 1, 2), (5, 3, 3), (130, 77, 3), (1100, 260, 2), (3000, 1030, 3), (257, 4100, 3)])
def test_qc_kernel_vs_numpy(n, p, depth):
    """Every output of ht_qc_u8 on random matrices with missing / multi-allelic / unphased calls sprinkled in, for
    host arrays in both major orders (chunked upload) and for a device tensor."""
    import torch

    from haptools_b200 import engine

    rng = np.random.default_rng(n + p)
    d = rng.integers(0, 2, size=(n, p, depth), dtype=np.uint8)
    d[rng.random((n, p, depth)) < 0.003] = 255
    d[rng.random((n, p, depth)) < 0.003] = 254
    d[rng.random((n, p, depth)) < 0.004] = 2
    for missing_min in (254, 255):
        g = d[:, :, :2]
        missing = np.any(g >= missing_min, axis=2)
        multi = np.any(g > 1, axis=2)
        b = d.astype(np.bool_)
        unphased = (b[:, :, 0] ^ b[:, :, 1]) & ~b[:, :, 2] if depth == 3 else np.zeros((n, p), dtype=bool)
        first = lambda m: tuple(int(x[0]) for x in np.nonzero(m)) if m.any() else None  # noqa: E731
        for lay, arr in _layouts3(d).items():
            for chunk_bytes in (256 << 20, 64 << 10):
                r = engine.qc_host(arr, missing_min=missing_min, chunk_bytes=chunk_bytes)
                assert r.first_missing == first(missing) and r.first_multiallelic == first(multi)
                assert r.first_unphased == first(unphased)
                assert r.n_missing == int(missing.sum()) and r.n_multiallelic == int(multi.sum())
                np.testing.assert_array_equal(r.samp_missing, missing.any(axis=1))
                np.testing.assert_array_equal(r.var_multiallelic, multi.any(axis=0))
                np.testing.assert_array_equal(r.alt_counts, g.astype(bool).sum(axis=(0, 2)))
        t = torch.from_numpy(d).cuda()
        for view in (t, t.permute(1, 0, 2).contiguous().permute(1, 0, 2)):
            r = engine.qc_device(view, missing_min=missing_min)
            assert r.first_missing == first(missing) and r.first_unphased == first(unphased)
            np.testing.assert_array_equal(r.alt_counts, g.astype(bool).sum(axis=(0, 2)))
            np.testing.assert_array_equal(r.samp_missing, missing.any(axis=1))
