"""
Latency of the class API (Haplotypes.transform on host objects, result back on the host) at the sizes of
the reference's own benchmark script (tests/bench_transform.py: 5,000 samples x 115 variants x 20
haplotypes of 1..115 alleles) and a few larger ones, next to the numpy restatement of the reference on this
box's host (1 thread).  Small cases are launch / allocation / copy latency, not bandwidth.
"""
import os, sys, time, types, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from haptools_b200 import data
from oracle import transform_oracle as orc

def build(n, p, H, kmax, seed=0):
    rng = np.random.default_rng(seed)
    gts = data.GenotypesVCF(fname=None)
    gts.data = rng.integers(0, 2, size=(n, p, 2), dtype=np.uint8)
    gts.variants = np.array([(f"SNP{i}", "chr1", i, ("T", "C")) for i in range(p)],
                            dtype=[("id", "U50"), ("chrom", "U10"), ("pos", np.uint32), ("alleles", object)])
    gts.samples = tuple(f"sample{i}" for i in range(n))
    haps = data.Haplotypes(fname=None)
    haps.data = {}
    for i in range(H):
        h = data.Haplotype(chrom="chr1", start=0, end=0, id=f"H{i}")
        k = int(rng.integers(1, kmax + 1))
        cols = rng.choice(p, size=k, replace=False)
        h.variants = tuple(data.Variant(0, 0, f"SNP{c}", str(rng.choice(["T", "C"]))) for c in cols)
        haps.data[h.id] = h
    return haps, gts

for (n, p, H, kmax) in ((5000, 115, 20, 115), (5000, 115, 90, 115), (3200, 80, 20, 80), (2504, 2000, 200, 20), (50_000, 2000, 1000, 20)):
    haps, gts = build(n, p, H, kmax)
    hap_list = list(haps.data.values())
    got = haps.transform(gts).data  # warm-up (CUDA context, library load, allocator)
    t = []
    for _ in range(5):
        t0 = time.perf_counter(); got = haps.transform(gts).data; t.append(time.perf_counter() - t0)
    c = []
    for _ in range(3):
        t0 = time.perf_counter(); want = orc.haplotypes_transform(hap_list, gts); c.append(time.perf_counter() - t0)
    assert np.array_equal(got, want)
    print(f"n={n} p={p} H={H} k<= {kmax}: B200 class API {min(t)*1e3:.2f} ms, numpy restatement {min(c)*1e3:.2f} ms, ratio {min(c)/min(t):.1f}x", flush=True)

if os.environ.get("HT_PROFILE"):
    import cProfile, pstats
    haps, gts = build(50_000, 2000, 1000, 20)
    haps.transform(gts)
    pr = cProfile.Profile(); pr.enable()
    for _ in range(3): haps.transform(gts)
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
