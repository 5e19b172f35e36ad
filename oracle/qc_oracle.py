"""
TEST INFRASTRUCTURE ONLY -- never imported by the product path (haptools_b200/).

numpy restatement of the reference's genotype checks, the host passes that ht_qc_u8 replaces:

    Genotypes.check_missing            haptools/data/genotypes.py:444-485
    GenotypesAncestry.check_missing    haptools/transform.py:353-383   (== 255 only, INFO log, ancestry rows too)
    Genotypes.check_biallelic          haptools/data/genotypes.py:487-528
    Genotypes.check_phase              haptools/data/genotypes.py:530-557
    Genotypes.check_maf                haptools/data/genotypes.py:559-625

Pinned to fixtures generated from the UNMODIFIED reference (tests/golden/make_golden_qc.py ->
tests/golden/qc_*.npz) and to the live reference in the build container (tests/test_qc.py).

A "state" is a dict: data (n, p, 2|3) uint8/bool, samples tuple, variants structured array (id, chrom, pos, ...),
ancestry (ancestry kind only), prephased bool, log list of (level, message).  Every function mutates the state
the way the reference method mutates its object, and raises what it raises.
"""
from __future__ import annotations

import numpy as np


def _first(mask):
    s, v = np.nonzero(mask)
    return int(s[0]), int(v[0])


def _where(state, v):
    return tuple(state["variants"][v])[:3]


def check_missing(state, discard_also=False):
    anc = state.get("ancestry") is not None
    d = state["data"][:, :, :2]
    # genotypes.py:463 (>= 254) vs transform.py:360 (== 255)
    missing = np.any(d == 255, axis=2) if anc else np.any(d >= 254, axis=2)
    if missing.any():
        if discard_also:
            rows = np.nonzero(missing)[0]
            before = len(state["samples"])
            state["data"] = np.delete(state["data"], rows, axis=0)
            if anc:
                state["ancestry"] = np.delete(state["ancestry"], rows, axis=0)
            state["samples"] = tuple(np.delete(state["samples"], rows))
            state["log"].append(("INFO" if anc else "WARNING",
                                 f"Ignoring missing genotypes from {before - len(state['samples'])} samples"))
        else:
            s, v = _first(missing)
            raise ValueError("Genotype with ID {} at POS {}:{} is missing for sample {}".format(*_where(state, v), state["samples"][s]))
    if discard_also and not state["data"].shape[0]:
        state["log"].append(("WARNING", "All samples were discarded! Check that that none of your variants are"
                                        " missing genotypes (GT: '.|.')."))


def check_biallelic(state, discard_also=False):
    if state["data"].dtype == np.bool_:
        state["log"].append(("WARNING", "All genotypes are already biallelic"))
        return
    multi = np.any(state["data"][:, :, :2] > 1, axis=2)
    if multi.any():
        if discard_also:
            cols = np.nonzero(multi)[1]
            state["log"].append(("INFO", f"Ignoring {len(cols)} multiallelic variants"))  # pairs, as the reference counts
            state["data"] = np.delete(state["data"], cols, axis=1)
            state["variants"] = np.delete(state["variants"], cols)
        else:
            s, v = _first(multi)
            raise ValueError("Variant with ID {} at POS {}:{} is multiallelic for sample {}".format(*_where(state, v), state["samples"][s]))
    if discard_also and not state["data"].shape[1]:
        state["log"].append(("WARNING", "All variants were discarded! Check that there are biallelic variants in your dataset."))
    state["data"] = state["data"].astype(np.bool_)


def check_phase(state):
    if state.get("prephased") or state["data"].shape[2] < 3:
        state["log"].append(("WARNING", "Phase information has already been removed from the data"))
        return
    b = state["data"].astype(np.bool_)
    unphased = (b[:, :, 0] ^ b[:, :, 1]) & ~b[:, :, 2]
    if unphased.any():
        s, v = _first(unphased)
        raise ValueError("Variant with ID {} at POS {}:{} is unphased for sample {}".format(*_where(state, v), state["samples"][s]))
    state["data"] = state["data"][:, :, :2]


def check_maf(state, threshold=None, discard_also=False, warn_only=False):
    d = state["data"]
    ref_af = d[:, :, :2].astype(np.bool_).sum(axis=(0, 2)) / (2 * d.shape[0])
    maf = np.array([ref_af, 1 - ref_af]).min(axis=0)
    if threshold is None:
        return maf
    rare = np.nonzero(maf < threshold)[0]
    if len(rare):
        if discard_also:
            before = len(state["variants"])
            state["data"] = np.delete(d, rare, axis=1)
            state["variants"] = np.delete(state["variants"], rare)
            maf = np.delete(maf, rare)
            state["log"].append(("INFO", f"Ignoring {before - len(state['variants'])} variants with MAF < {threshold}"))
        else:
            msg = "Variant with ID {} at POS {}:{} has MAF {} < {}".format(*(_where(state, rare[0]) + (maf[rare[0]], threshold)))
            if warn_only:
                state["log"].append(("WARNING", msg))
            else:
                raise ValueError(msg)
    return maf


CHECKS = {"check_missing": check_missing, "check_biallelic": check_biallelic, "check_phase": check_phase, "check_maf": check_maf}


def run_scenario(state, calls):
    """calls: list of (method name, kwargs).  Returns one outcome per call: {"raises": message} (the scenario stops
    there) or {"ret": maf or None}; the state is mutated along the way."""
    outcomes = []
    for name, kwargs in calls:
        try:
            outcomes.append({"ret": CHECKS[name](state, **kwargs)})
        except ValueError as e:
            outcomes.append({"raises": str(e)})
            break
    return outcomes
