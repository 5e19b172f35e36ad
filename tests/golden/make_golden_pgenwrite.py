#!/usr/bin/env python
"""
Golden vectors for the PGEN-writer hand-off: the arrays GenotypesPLINK.write passes to
pgenlib.PgenWriter.append_alleles_batch for one chunk of variants, produced by executing the
reference's OWN statements (haptools/data/genotypes.py:1579, 1618-1634) and its own
GenotypesPLINK._num_unique_alleles (:1538-1567) from /root/reference (imported through
oracle/ref_harness.py; pgenlib itself is not installed, so write() cannot run as a whole).
Build container only:   python tests/golden/make_golden_pgenwrite.py
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
from oracle import ref_harness  # noqa: E402

ns = ref_harness.load()
from haptools.data import GenotypesPLINK  # noqa: E402  (the reference's)


def reference_batch(gts, start, end):
    # the statements of GenotypesPLINK.write, verbatim in order
    data = gts.data.transpose((1, 0, 2))[:, :, :2]
    size = end - start
    missing = np.ascontiguousarray(data[start:end] == np.iinfo(np.uint8).max)
    allele_cts = gts._num_unique_alleles(data[start:end])
    subset_data = np.ascontiguousarray(data[start:end], dtype=np.int32)
    subset_data.resize((size, len(gts.samples) * 2))
    missing.resize((size, len(gts.samples) * 2))
    subset_data[missing] = -9
    return subset_data, allele_cts


def main():
    rng = np.random.default_rng(11)
    cases = {
        "bool_result": (rng.random((37, 9, 2)) < 0.3),                       # what transform produces
        "bool_constant_rows": np.stack([np.zeros((21, 2), bool), np.ones((21, 2), bool), rng.random((21, 2)) < 0.5], axis=1),
        "uint8_multiallelic_missing": rng.choice(np.array([0, 1, 2, 255], dtype=np.uint8), size=(16, 7, 3), p=[0.5, 0.3, 0.1, 0.1]),
    }
    for name, arr in cases.items():
        gts = GenotypesPLINK.__new__(GenotypesPLINK)
        gts.data = arr
        gts.samples = tuple(f"s{i}" for i in range(arr.shape[0]))
        p = arr.shape[1]
        out = {"data": arr}
        for k, (a, b) in enumerate([(0, p), (1, max(2, p // 2)), (p - 1, p)]):
            sub, cts = reference_batch(gts, a, b)
            out[f"range{k}"] = np.array([a, b])
            out[f"subset{k}"] = sub
            out[f"cts{k}"] = cts
        np.savez_compressed(HERE / f"pgenwrite_{name}.npz", **out)
        print(name, arr.shape, arr.dtype)


if __name__ == "__main__":
    main()
