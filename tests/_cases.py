"""
Portable test-case files: inputs and expected outputs of the transform hot path as plain
arrays + JSON, so golden vectors generated from the reference in the build container
(tests/golden/make_golden.py) can be replayed anywhere (the GPU box has no reference).

A case is a dict:
  kind            "plain" | "ancestry"
  gt              (n, p, 2|3) uint8 or bool
  samples         list[str]
  var_id/var_chrom/var_pos  per-variant metadata
  var_alleles     list[list[str]]
  haps            list[dict(id, chrom, start, end, ancestry?, variants=[dict(start,end,id,allele)])]
  ancestry        (n, p, 2) uint8            (kind == "ancestry")
  ancestry_labels dict[str, int]             (kind == "ancestry")
  expected_plural (n, H, 2) bool             output of Haplotypes[Ancestry].transform
  expected_single (H, n, 2) bool             output of each Haplotype[Ancestry].transform
"""
from __future__ import annotations

import json
from pathlib import Path

import numpy as np

GOLDEN_DIR = Path(__file__).parent / "golden"


def save_case(path, case: dict) -> None:
    arrays = {
        "gt": case["gt"],
        "var_pos": np.asarray(case["var_pos"], dtype=np.uint32),
        "expected_plural": case["expected_plural"],
        "expected_single": case["expected_single"],
        "meta": np.array(
            json.dumps(
                {
                    "kind": case["kind"],
                    "samples": list(case["samples"]),
                    "var_id": list(case["var_id"]),
                    "var_chrom": list(case["var_chrom"]),
                    "var_alleles": [list(a) for a in case["var_alleles"]],
                    "haps": case["haps"],
                    "ancestry_labels": case.get("ancestry_labels"),
                    "source": case.get("source", ""),
                }
            )
        ),
    }
    if case.get("ancestry") is not None:
        arrays["ancestry"] = case["ancestry"]
    np.savez_compressed(path, **arrays)


def load_case(path) -> dict:
    with np.load(path, allow_pickle=False) as z:
        meta = json.loads(str(z["meta"]))
        case = dict(meta)
        case["gt"] = z["gt"]
        case["var_pos"] = z["var_pos"]
        case["expected_plural"] = z["expected_plural"]
        case["expected_single"] = z["expected_single"]
        case["ancestry"] = z["ancestry"] if "ancestry" in z.files else None
    return case


def list_golden() -> list[Path]:
    """Transform fixtures (the bp_*.npz breakpoint fixtures have their own loader)."""
    return sorted(p for p in GOLDEN_DIR.glob("*.npz") if not p.name.startswith(("bp_", "pgenwrite_", "qc_")))


def list_golden_bp() -> list[Path]:
    return sorted(GOLDEN_DIR.glob("bp_*.npz"))


def load_bp_case(path, breakpoints_cls, hapblock_dtype):
    """Rebuild a Breakpoints object (un-encoded) + the variants array from a bp_*.npz fixture."""
    with np.load(path, allow_pickle=False) as z:
        meta = json.loads(str(z["meta"]))
        n = int(z["n"])
        data = {}
        for s in range(n):
            strands = []
            for c in range(2):
                sel = (z["blk_sample"] == s) & (z["blk_strand"] == c)
                rows = [(str(pop), str(chrom), int(bp), float(bp) / 1e4)
                        for pop, chrom, bp in zip(z["blk_pop"][sel], z["blk_chrom"][sel], z["blk_bp"][sel])]
                strands.append(np.array(rows, dtype=hapblock_dtype))
            data[f"S{s}"] = strands
        variants = np.array(list(zip(z["var_chrom"], z["var_pos"])), dtype=[("chrom", "U10"), ("pos", np.uint32)])
        expected = z["expected"]
    bps = breakpoints_cls(fname=None)
    bps.data = data
    return bps, variants, meta, expected


def build_objects(case: dict, ns, with_alleles: bool = True):
    """
    Build (haps_container, genotypes) objects out of a case using the classes in the
    namespace ``ns`` -- either the reference's (oracle/ref_harness.load()) or the
    haptools_b200 host mirror; both expose Variant/Haplotype/Haplotypes/GenotypesVCF/
    HaplotypeAncestry/HaplotypesAncestry/GenotypesAncestry.
    """
    anc = case["kind"] == "ancestry"
    hap_cls = ns.HaplotypeAncestry if anc else ns.Haplotype
    haps_cls = ns.HaplotypesAncestry if anc else ns.Haplotypes
    gts_cls = ns.GenotypesAncestry if anc else ns.GenotypesVCF
    hapdict = {}
    for h in case["haps"]:
        kwargs = dict(chrom=h["chrom"], start=h["start"], end=h["end"], id=h["id"])
        if anc:
            kwargs["ancestry"] = h["ancestry"]
        hap = hap_cls(**kwargs)
        hap.variants = tuple(
            ns.Variant(start=v["start"], end=v["end"], id=v["id"], allele=v["allele"])
            for v in h["variants"]
        )
        hapdict[h["id"]] = hap
    haps = haps_cls(fname=None)
    haps.data = hapdict
    gts = gts_cls(fname=None)
    gts.data = case["gt"]
    gts.samples = tuple(case["samples"])
    gts.variants = np.array(
        [
            (vid, chrom, pos, tuple(alleles))
            for vid, chrom, pos, alleles in zip(
                case["var_id"], case["var_chrom"], case["var_pos"], case["var_alleles"]
            )
        ],
        dtype=[
            ("id", "U50"),
            ("chrom", "U10"),
            ("pos", np.uint32),
            ("alleles", object),
        ],
    )
    if anc:
        gts.ancestry = case["ancestry"]
        gts.ancestry_labels = dict(case["ancestry_labels"])
    return haps, gts


def random_case(
    rng: np.random.Generator,
    n: int,
    p: int,
    n_haps: int,
    kind: str = "plain",
    max_alleles: int = 2,
    missing_rate: float = 0.0,
    phase_column: bool = False,
    as_bool: bool = False,
    max_hap_len: int = 5,
    empty_hap: bool = False,
    n_pops: int = 3,
    layout: str = "C",
) -> dict:
    """A random case WITHOUT expected outputs (the generator fills those in)."""
    bases = ["A", "C", "G", "T", "AT", "GC"]
    var_alleles = []
    for _ in range(p):
        m = int(rng.integers(2, max_alleles + 1)) if max_alleles > 2 else 2
        var_alleles.append(list(rng.permutation(bases)[:m]))
    depth = 3 if phase_column else 2
    gt = np.zeros((n, p, depth), dtype=np.uint8)
    for j in range(p):
        m = len(var_alleles[j])
        gt[:, j, :2] = rng.integers(0, m, size=(n, 2), dtype=np.uint8)
    if missing_rate > 0:
        miss = rng.random((n, p, 2)) < missing_rate
        gt[:, :, :2][miss] = np.where(rng.random(miss.sum()) < 0.5, 255, 254)
    if phase_column:
        gt[:, :, 2] = rng.integers(0, 2, size=(n, p), dtype=np.uint8)
    if as_bool:
        gt = gt.astype(np.bool_)
    if layout == "F":
        # variant-major physical order, as produced by the reference's VCF reader
        # (haptools/data/genotypes.py:183-186, 210): a transposed view of (p, n, d)
        gt = np.ascontiguousarray(gt.transpose((1, 0, 2))).transpose((1, 0, 2))
    var_id = [f"1:{1000 + 3 * j}:{var_alleles[j][0]}:{var_alleles[j][1]}" for j in range(p)]
    haps = []
    pops = [f"POP{k}" for k in range(n_pops)]
    for h in range(n_haps):
        k = int(rng.integers(1, max(2, min(max_hap_len, p) + 1)))
        if empty_hap and h == n_haps // 2:
            k = 0
        cols = np.sort(rng.choice(p, size=k, replace=False)) if k else []
        variants = []
        for j in cols:
            j = int(j)
            allele = var_alleles[j][int(rng.integers(0, len(var_alleles[j])))]
            variants.append(
                dict(start=1000 + 3 * j, end=1001 + 3 * j, id=var_id[j], allele=allele)
            )
        hap = dict(
            id=f"H{h}",
            chrom="1",
            start=variants[0]["start"] if variants else 1000,
            end=variants[-1]["end"] if variants else 1001,
            variants=variants,
        )
        if kind == "ancestry":
            hap["ancestry"] = pops[int(rng.integers(0, n_pops))]
        haps.append(hap)
    case = dict(
        kind=kind,
        gt=gt,
        samples=[f"S{i}" for i in range(n)],
        var_id=var_id,
        var_chrom=["1"] * p,
        var_pos=[1000 + 3 * j for j in range(p)],
        var_alleles=var_alleles,
        haps=haps,
    )
    if kind == "ancestry":
        # piecewise-constant local ancestry along variants
        anc = np.zeros((n, p, 2), dtype=np.uint8)
        for s in range(n):
            for c in range(2):
                j = 0
                while j < p:
                    run = int(rng.integers(1, max(2, p // 2)))
                    anc[s, j : j + run, c] = rng.integers(0, n_pops)
                    j += run
        case["ancestry"] = anc
        case["ancestry_labels"] = {pop: k for k, pop in enumerate(pops)}
    return case
