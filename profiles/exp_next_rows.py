"""Measurements for the SURVEY.md 8(f) "next" rows: N1 streamed PGEN chunks, N2 fused allele counts,
N3 local ancestry from breakpoints on the device."""
import os, sys, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from haptools_b200 import engine, synth, stream, _cuda, data
from haptools_b200.data.breakpoints import HapBlockEncoded
dev = torch.device("cuda")

# ---- N1: a bounded sample of BASELINE configs[4] (200k samples, 1,024-variant int32 chunks)
n, chunk, n_chunks = 200_000, 1024, 12
V = chunk * n_chunks
plan = synth.synth_plan(V, 2000, seed=3)
pool = np.empty((chunk, 2 * n), dtype=np.int32)
pool[:] = np.random.default_rng(0).integers(0, 2, size=(64, 2 * n), dtype=np.int32)[np.arange(chunk) % 64]
def fill(idx, out):  # stands for pgenlib decoding into the caller's buffer: a plain memcpy here
    np.copyto(out, pool[: len(idx)])
reader = stream.ArrayPgenReader(fill=fill, n_samples=n)
for rep in range(2):
    st = {}
    t0 = time.perf_counter()
    res = stream.transform_pgen_stream(reader, np.arange(V, dtype=np.uint32), n, plan, chunk_variants=chunk, stats=st, to_host=False)
    torch.cuda.synchronize()
wall = st["stream_s"]
print(f"N1 stream: {st['chunks']} chunks x {chunk} variants x {n} samples: H2D {st['h2d_bytes']/1e9:.1f} GB; "
      f"reader {st['read_s']:.3f} s, H2D {st['h2d_s']:.3f} s ({st['h2d_bytes']/st['h2d_s']/1e9:.1f} GB/s), pack {st['pack_s']:.4f} s "
      f"({st['h2d_bytes']/st['pack_s']/1e9:.0f} GB/s), stream wall {wall:.3f} s -> overlap efficiency "
      f"{max(st['read_s'], st['h2d_s'], st['pack_s'])/wall:.2f}; {V/wall:.0f} variants/s")
del res, pool
torch.cuda.empty_cache()

# ---- N2: cost of the fused allele counts
n, p, H = 200_000, 50_000, 100_000
plan = synth.synth_plan(p, H, seed=2)
dp = engine.DevicePlan(plan)
planes = engine.alloc_planes(n, plan.n_keys, dev); planes.random_(0, 2**31 - 1)
out = torch.empty((n, 2 * H), dtype=torch.uint8, device=dev)
counts = torch.zeros(H, dtype=torch.int64, device=dev)
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
a = timeit(lambda: engine.transform_u8(dp, planes, n, out.data_ptr(), 2 * H))
b = timeit(lambda: engine.transform_u8(dp, planes, n, out.data_ptr(), 2 * H, counts=counts))
def separate():
    acc = torch.zeros(H, dtype=torch.int32, device=dev)
    v = out.view(n, H, 2)
    for i in range(0, n, 4096):
        acc += v[i : i + 4096].sum(dim=(0, 2), dtype=torch.int32)
    return acc
c = timeit(separate, reps=2)
print(f"N2 counts: transform {a:.2f} ms, with fused counts {b:.2f} ms (+{100*(b-a)/a:.1f} %), separate pass over the result {c:.2f} ms")
del planes, out
torch.cuda.empty_cache()

# ---- N3: breakpoints -> ancestry labels
n, p, nblk = 20_000, 200_000, 12
rng = np.random.default_rng(1)
bps = data.Breakpoints(fname=None)
bps.data = {}
for s in range(n):
    strands = []
    for c in range(2):
        ends = np.sort(rng.choice(np.arange(1, 10_000_000, dtype=np.uint32), size=nblk, replace=False)); ends[-1] = 10_000_000
        blk = np.zeros(nblk, dtype=HapBlockEncoded)
        blk["pop"] = rng.integers(0, 5, size=nblk); blk["chrom"] = "1"; blk["bp"] = ends
        strands.append(blk)
    bps.data[f"S{s}"] = strands
bps.labels = {f"P{i}": i for i in range(5)}
variants = np.zeros(p, dtype=[("chrom", "U10"), ("pos", np.uint32)])
variants["chrom"] = "1"; variants["pos"] = np.sort(rng.integers(1, 10_000_000, size=p))
t0 = time.perf_counter(); anc = bps.population_tensor(variants); torch.cuda.synchronize(); t1 = time.perf_counter()
names, chroms, var_key, keys, pops, seg_off = bps._flatten(variants, None)
d = [torch.from_numpy(x.view(t)).to(dev) for x, t in ((keys, np.int64), (pops, np.uint8), (seg_off, np.int32), (var_key, np.int64))]
bad = torch.full((1,), -1, dtype=torch.int64, device=dev)
k = timeit(lambda: _cuda.check(_cuda.lib().ht_bp_ancestry(d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(), anc.data_ptr(), *anc.stride(), n, p, bad.data_ptr(), 0), "bp"))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import breakpoints_oracle as borc
sub = {k_: bps.data[k_] for k_ in list(bps.data)[:200]}
t2 = time.perf_counter(); want = borc.population_array(sub, variants); t3 = time.perf_counter()
assert np.array_equal(anc[:200].cpu().numpy(), want)
print(f"N3 bp->ancestry: {n} samples x {p} variants x {nblk} blocks/strand: kernel {k:.2f} ms ({n*p*2/k/1e6:.0f} G labels/s), "
      f"whole call incl. host flattening {t1-t0:.2f} s; numpy restatement of the reference {1e3*(t3-t2)/200*n/1e3:.1f} s extrapolated from 200 samples")
