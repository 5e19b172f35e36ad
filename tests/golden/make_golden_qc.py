#!/usr/bin/env python
"""
Golden scenarios for the genotype checks (check_missing / check_biallelic / check_phase / check_maf), generated
by running the UNMODIFIED reference classes (haptools.data.GenotypesVCF and haptools.transform.GenotypesAncestry
from /root/reference, imported through oracle/ref_harness.py).  Build container only:

    python tests/golden/make_golden_qc.py

qc_<name>.npz: the input matrix (and ancestry), the calls made, and what the reference did: per call the
ValueError message or the returned MAF, then the object's data / samples / variant IDs afterwards and every log
record it emitted.
"""
from __future__ import annotations

import json
import logging
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
from oracle import ref_harness  # noqa: E402

ns = ref_harness.load()


class _Capture(logging.Handler):
    def __init__(self):
        super().__init__(level=logging.DEBUG)
        self.records = []

    def emit(self, record):
        self.records.append((record.levelname, record.getMessage()))


def build(kind, data, ancestry=None, prephased=False):
    log = logging.getLogger(f"qc_golden_{id(data)}")
    log.setLevel(logging.DEBUG)
    log.propagate = False
    cap = _Capture()
    log.addHandler(cap)
    gts = (ns.GenotypesAncestry if kind == "ancestry" else ns.GenotypesVCF)(fname=None, log=log)
    n, p = data.shape[:2]
    gts.data = data.copy()
    gts.samples = tuple(f"S{i}" for i in range(n))
    gts.variants = np.array([(f"v{j}", "1" if j % 3 else "2", 100 + 7 * j, ("A", "C", "G")) for j in range(p)],
                            dtype=gts.variants.dtype)
    gts._prephased = prephased
    if kind == "ancestry":
        gts.ancestry = ancestry.copy()
    return gts, cap


def run(name, kind, data, calls, ancestry=None, prephased=False):
    gts, cap = build(kind, data, ancestry, prephased)
    outcomes, rets = [], {}
    for i, (method, kwargs) in enumerate(calls):
        try:
            r = getattr(gts, method)(**kwargs)
            outcomes.append({"ret": r is not None})
            if r is not None:
                rets[f"ret{i}"] = np.asarray(r)
        except ValueError as e:
            outcomes.append({"raises": str(e)})
            break
    arrays = dict(
        data=data, final_data=gts.data, final_ids=np.array(list(gts.variants["id"]), dtype="U50"),
        meta=np.array(json.dumps({"kind": kind, "calls": calls, "outcomes": outcomes, "prephased": prephased,
                                  "final_samples": list(gts.samples), "log": cap.records})),
        **rets,
    )
    if ancestry is not None:
        arrays["ancestry"] = ancestry
        arrays["final_ancestry"] = gts.ancestry
    np.savez_compressed(HERE / f"qc_{name}.npz", **arrays)
    print(f"qc_{name}: {outcomes} log={cap.records}")


def main():
    rng = np.random.default_rng(20)
    pipeline = [("check_missing", {}), ("check_biallelic", {}), ("check_phase", {}), ("check_maf", {})]

    def clean(n, p):
        d = rng.integers(0, 2, size=(n, p, 3), dtype=np.uint8)
        d[:, :, 2] = 1
        return d

    run("clean_pipeline", "plain", clean(6, 5), pipeline)
    d = clean(7, 6)
    d[2, 3, 1] = 255
    d[1, 4, 0] = 255
    run("missing_raises_first_pair", "plain", d, pipeline)
    d = clean(5, 4)
    d[3, 1, 0] = 254
    run("missing_254_plain_raises", "plain", d, [("check_missing", {})])
    anc = rng.integers(0, 3, size=(5, 4, 2), dtype=np.uint8)
    run("missing_254_ancestry_passes", "ancestry", d[:, :, :2].copy(), [("check_missing", {})], ancestry=anc)
    d2 = d[:, :, :2].copy()
    d2[0, 2, 1] = 255
    d2[4, 0, 0] = 255
    run("missing_255_ancestry_discard", "ancestry", d2, [("check_missing", {"discard_also": True})], ancestry=anc)
    run("missing_255_ancestry_raises", "ancestry", d2, [("check_missing", {})], ancestry=anc)
    d = clean(8, 5)
    d[1, 1, 0] = 255
    d[1, 3, 1] = 254
    d[6, 0, 1] = 255
    d[2, 2, 0] = 2
    d[5, 2, 1] = 2
    d[5, 4, 0] = 3
    run("discard_then_discard", "plain", d, [("check_missing", {"discard_also": True}), ("check_biallelic", {"discard_also": True}),
                                             ("check_phase", {}), ("check_maf", {})])
    d = clean(3, 4)
    d[:, 1, 0] = 255
    run("all_samples_discarded", "plain", d, [("check_missing", {"discard_also": True})])
    d = clean(4, 3)
    d[2, 0, 0] = 2
    d[0, 1, 1] = 5
    d[3, 2, 1] = 2
    run("all_variants_discarded", "plain", d, [("check_biallelic", {"discard_also": True})])
    run("multiallelic_raises", "plain", d, [("check_biallelic", {})])
    d = clean(5, 4)
    d[4, 1, 0] = 255  # a missing call is also "> 1"
    run("missing_counts_as_multiallelic", "plain", d, [("check_biallelic", {})])
    d = clean(6, 5)
    d[3, 2] = (1, 0, 0)
    d[1, 4] = (0, 1, 0)
    run("unphased_het_raises", "plain", d, [("check_phase", {})])
    d = clean(6, 5)
    d[3, 2] = (1, 2, 0)   # both non-zero: the bool cast makes it "homozygous" -> passes
    d[2, 1] = (1, 1, 0)
    d[0, 0] = (0, 0, 0)
    run("unphased_1_2_passes", "plain", d, [("check_phase", {})])
    d = clean(6, 5)
    d[5, 0] = (0, 2, 0)
    d[4, 4] = (1, 0, 7)   # any non-zero phase byte is "phased"
    run("unphased_0_2_raises", "plain", d, [("check_phase", {})])
    run("prephased_flag", "plain", clean(4, 3), [("check_phase", {})], prephased=True)
    run("depth2_already_removed", "plain", clean(4, 3)[:, :, :2].copy(), [("check_phase", {})])
    run("bool_already_biallelic", "plain", clean(4, 3)[:, :, :2].astype(np.bool_), [("check_biallelic", {}), ("check_maf", {})])
    d = clean(40, 6)
    d[:, 2, :2] = 0
    d[0, 2, 0] = 1          # MAF 1/80
    d[:, 4, :2] = 1
    d[3, 4, 1] = 0          # MAF 1/80 from the other side
    run("maf_threshold_raises", "plain", d, [("check_maf", {"threshold": 0.05})])
    run("maf_threshold_warns", "plain", d, [("check_maf", {"threshold": 0.05, "warn_only": True})])
    run("maf_threshold_discards", "plain", d, [("check_maf", {"threshold": 0.05, "discard_also": True})])
    run("maf_of_multiallelic_counts_nonzero", "plain", rng.integers(0, 4, size=(30, 7, 2), dtype=np.uint8), [("check_maf", {})])
    d = rng.integers(0, 2, size=(300, 70, 3), dtype=np.uint8)
    d[:, :, 2] = rng.random((300, 70)) < 0.98
    hom = d[:, :, 0] == d[:, :, 1]
    d[:, :, 2] |= ~hom & (rng.random((300, 70)) < 0.97)
    run("random_unphased", "plain", d, [("check_missing", {}), ("check_biallelic", {}), ("check_phase", {})])
    d = clean(300, 70)
    d[rng.random((300, 70, 3)) < 0.002] = 255
    d[:, :, 2] = 1
    run("random_missing_discard", "plain", d, pipeline[:0] + [("check_missing", {"discard_also": True}), ("check_biallelic", {}),
                                                             ("check_phase", {}), ("check_maf", {})])


if __name__ == "__main__":
    main()
