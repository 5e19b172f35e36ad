"""Measurements for the N4 / N1b rows: planner (reference loop vs vectorised), VCF text writer,
haplotype-major int32 batches for the PGEN writer."""
import io, os, sys, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from haptools_b200 import data, engine, stream, synth, writers
from haptools_b200.plan import plan_from_haplotypes, plan_from_haplotypes_loop

# ---- planner at UKB scale (host only) -------------------------------------------------
p, H = 50_000, 100_000
ref_col, ref_allele, hap_off, _ = synth.synth_haplotype_refs(p, H, 2)
al = ("A", "C")
haps = []
for h in range(H):
    hp = data.Haplotype("1", h + 1, h + 2, f"H{h}")
    hp.variants = tuple(data.Variant(int(c) + 1, int(c) + 2, f"v{c}", al[a]) for c, a in zip(ref_col[hap_off[h]:hap_off[h + 1]], ref_allele[hap_off[h]:hap_off[h + 1]]))
    haps.append(hp)
g = data.GenotypesVCF(None)
g.data = np.zeros((4, p, 2), dtype=np.uint8)
g.variants = np.array([(f"v{i}", "1", i + 1, ("A", "C")) for i in range(p)], dtype=g.variants.dtype)
g.samples = tuple("s%d" % i for i in range(4))
for name, fn in (("reference loop", plan_from_haplotypes_loop), ("vectorised", plan_from_haplotypes)):
    best = min(_t for _t in [(lambda t0: (fn(haps, g), time.perf_counter() - t0)[1])(time.perf_counter()) for _ in range(3)])
    print(f"planner, {name}: {best:.2f} s for K = {len(ref_col)} alleles, H = {H}")

# ---- writers ------------------------------------------------------------------------------
dev = torch.device("cuda")
n, Hw = 100_000, 4_096
tiles = (n + 1023) // 1024
bits = torch.empty((tiles, Hw, 32, 2), dtype=torch.int32, device=dev); bits.random_(0, 2**31 - 1)
variants = np.empty(Hw, dtype=data.GenotypesVCF(None).variants.dtype)
variants["id"] = [f"H{i}" for i in range(Hw)]; variants["chrom"] = "1"; variants["pos"] = np.arange(Hw) + 1; variants["alleles"].fill(("A", "T"))
samples = [f"S{i}" for i in range(n)]
class Sink(io.RawIOBase):
    def __init__(self): self.n = 0
    def write(self, b): self.n += len(b); return len(b)
for _ in range(2):
    sink = Sink(); t0 = time.perf_counter()
    writers.write_vcf_from_bits(sink, variants, samples, bits, n)
    dt = time.perf_counter() - t0
print(f"VCF writer: {Hw} records x {n} samples = {sink.n/1e9:.2f} GB of text in {dt:.2f} s ({sink.n/dt/1e9:.2f} GB/s, {Hw*n*2/dt/1e9:.2f} G hap-calls/s) to a null sink")
# kernel alone
out = torch.empty((1024, 4 * n), dtype=torch.uint8, device=dev)
from haptools_b200 import _cuda
def k(): _cuda.check(_cuda.lib().ht_format_vcf_gt(bits.data_ptr(), n, Hw, 0, 1024, out.data_ptr(), 4 * n, 0), "fmt")
k(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); [k() for _ in range(10)]; e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"ht_format_vcf_gt alone: {out.numel()/ms/1e6:.0f} GB/s of text")
# the reference's per-sample Python loop (restated without pysam: string formatting only) on a slice
mat = (np.random.default_rng(0).random((2000, 8, 2)) < 0.3)
t0 = time.perf_counter()
for j in range(mat.shape[1]):
    "\t".join(f"{int(a)}|{int(b)}" for a, b in mat[:, j])
dt = time.perf_counter() - t0
print(f"Python per-sample loop (formatting only, no pysam): {mat.shape[0]*mat.shape[1]*2/dt/1e6:.1f} M hap-calls/s")
t0 = time.perf_counter(); rows = 0
for a, b, sub, cts in stream.pgen_write_batches(bits, n, chunk_haps=256): rows += b - a
dt = time.perf_counter() - t0
print(f"PGEN batches (int32, {rows} rows x {2*n}): {rows*2*n*4/dt/1e9:.1f} GB/s to the host, {rows*n*2/dt/1e9:.2f} G hap-calls/s")
