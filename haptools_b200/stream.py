"""
Streamed PGEN input (BASELINE.json configs[4]).

The reference materialises the whole `(n, p, 3)` uint8 matrix on the host before transforming
(`GenotypesPLINK.read`, haptools/data/genotypes.py:1298-1406: pgenlib fills `(chunk, 2n)` int32
buffers that are cast and transposed chunk by chunk).  Here those same int32 chunk buffers are
the unit of work: the reader fills page-locked ring buffers, each chunk is DMA'd on a copy
stream while the next one is being read, and `ht_pack_i32` turns it straight into bit-planes
-- `gts.data` is never built.  Only the variants some haplotype references are read (as
`transform_haps` already asks for, haptools/transform.py:602-620).  One transform launch
follows once all planes exist.
"""
from __future__ import annotations

import time

import numpy as np
import torch

from . import _cuda, engine
from ._cuda import check
from .plan import TransformPlan


class ArrayPgenReader:
    """
    Stand-in for `pgenlib.PgenReader` with its buffer contract (pgenlib is not installed in
    this image): `read_alleles_list(variant_idxs uint32[c], allele_int32_out int32[c, 2n])`,
    allele codes 0/1/2..., -9 = missing.  Backed by a `(p, 2n)` int32 array or by a callable
    `fill(variant_idxs, out)`.
    """

    def __init__(self, alleles=None, n_samples: int = None, fill=None):
        self._alleles, self._fill = alleles, fill
        self._n = n_samples if n_samples is not None else alleles.shape[1] // 2

    def get_raw_sample_ct(self) -> int:
        return self._n

    def read_alleles_list(self, variant_idxs, allele_int32_out, hap_maj: bool = False):
        if hap_maj:
            raise NotImplementedError("haplotype-major mode is not part of the contract used")
        if self._fill is not None:
            self._fill(variant_idxs, allele_int32_out)
        else:
            np.take(self._alleles, variant_idxs, axis=0, out=allele_int32_out)


def _ancestry_chunks(ancestry, n: int, p: int, dev):
    """Normalise the `ancestry` argument of transform_pgen_stream into `labels(cols) -> CUDA uint8 (n, c, 2)`
    (launched on the current stream) and `check()`."""
    if hasattr(ancestry, "labels") and hasattr(ancestry, "check"):  # Breakpoints.population_provider
        if ancestry.n_samples != n:
            raise ValueError("the breakpoints and the genotypes differ in their samples")
        return (lambda cols: ancestry.labels(cols, variant_major=True)), ancestry.check
    if callable(ancestry):
        return ancestry, (lambda: None)
    if isinstance(ancestry, torch.Tensor):
        if not ancestry.is_cuda or ancestry.dtype != torch.uint8 or tuple(ancestry.shape[:2]) != (n, p) or ancestry.shape[2] < 2:
            raise ValueError("ancestry must be a CUDA uint8 tensor shaped (samples, variants, 2)")
        return (lambda cols: ancestry.index_select(1, torch.from_numpy(np.ascontiguousarray(cols, dtype=np.int64)).to(dev))), (lambda: None)
    a = np.asarray(ancestry)
    if a.dtype != np.uint8 or a.shape[:2] != (n, p) or a.shape[2] < 2:
        raise ValueError("ancestry must be a uint8 array shaped (samples, variants, 2)")
    # a host matrix (GenotypesAncestry.ancestry): the chunk's columns are gathered on the host and uploaded
    return (lambda cols: torch.from_numpy(np.ascontiguousarray(a[:, cols, :2])).to(dev, non_blocking=False)), (lambda: None)


def transform_pgen_stream(reader, variant_idxs: np.ndarray, n_samples: int, plan: TransformPlan,
                          chunk_variants: int = 1024, device=None, n_host_buffers: int = 3,
                          fmt: str = "u8", to_host: bool = True, stats: dict = None, ancestry=None):
    """
    reader        object with pgenlib's `read_alleles_list(idx, int32[c, 2n])`
    ancestry      for ancestry plans (what `transform --ancestry` does on PGEN input, haptools/transform.py:
                  610-660): where the local-ancestry labels of a chunk's variants come from -- a
                  `Breakpoints.population_provider(...)` (labels computed on the device from the .bp blocks, chunk
                  by chunk: neither gts.data nor the (n, p, 2) ancestry matrix is ever built), a `(n, p, 2)` uint8
                  matrix (CUDA tensor or host array, columns = the plan's), or any callable
                  `cols -> CUDA uint8 tensor (n, len(cols), 2)`
    variant_idxs  uint32 indices (into the PGEN) of the variants to read; column j of the plan
                  is variant_idxs[j] (so plan.n_cols == len(variant_idxs))
    Returns the `(n, H, 2)` bool result (numpy if `to_host`, else a device uint8 tensor), or the
    packed result with fmt="bits".  `stats` (optional dict) receives timings and byte counts.
    """
    engine.require_cuda()
    variant_idxs = np.ascontiguousarray(variant_idxs, dtype=np.uint32)
    p = len(variant_idxs)
    if p != plan.n_cols:
        raise ValueError(f"plan was made for {plan.n_cols} variant columns, {p} variants are to be read")
    if plan.has_ancestry != (ancestry is not None):
        raise ValueError("ancestry plan and ancestry source must be given together")
    n, H = int(n_samples), plan.n_haps
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    anc_labels, anc_check = _ancestry_chunks(ancestry, n, p, dev) if ancestry is not None else (None, None)
    lib = _cuda.lib()
    t_start = time.perf_counter()
    with torch.cuda.device(dev):
        dplan = engine.DevicePlan(plan, dev)
        planes = engine.alloc_planes(n, plan.n_keys, dev)
        # only referenced columns are read: chunk over plan.var_col
        cols = plan.var_col.astype(np.int64)
        V = len(cols)
        chunk = max(1, min(int(chunk_variants), max(V, 1)))
        row_in_chunk = torch.from_numpy((np.arange(V, dtype=np.int64) % chunk).astype(np.int32)).to(dev)
        host = [engine.pinned_empty((chunk, 2 * n), np.int32) for _ in range(min(n_host_buffers, max(1, -(-V // chunk))))]
        devbuf = [torch.empty((chunk, 2 * n), dtype=torch.int32, device=dev) for _ in range(2)]
        s_in, s_cmp = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        cur = torch.cuda.current_stream(dev)
        s_in.wait_stream(cur)
        s_cmp.wait_stream(cur)
        ev_h2d = [None] * len(host)   # host buffer may be refilled once its copy has left
        ev_pack = [None] * 2          # device buffer may be overwritten once it has been packed
        t_read = 0.0
        h2d_bytes = 0
        timing = stats is not None
        ev_pairs = []  # (h2d start, h2d end, pack start, pack end) CUDA events per chunk
        ev_start = s_in.record_event(torch.cuda.Event(enable_timing=True)) if timing else None
        for k, v0 in enumerate(range(0, V, chunk)):
            v1 = min(V, v0 + chunk)
            c = v1 - v0
            hb, db = k % len(host), k % 2
            if ev_h2d[hb] is not None:
                ev_h2d[hb].synchronize()
            t0 = time.perf_counter()
            reader.read_alleles_list(variant_idxs[cols[v0:v1]], host[hb][:c])
            t_read += time.perf_counter() - t0
            if ev_pack[db] is not None:
                s_in.wait_event(ev_pack[db])
            nbytes = c * 2 * n * 4
            e0 = s_in.record_event(torch.cuda.Event(enable_timing=True)) if timing else None
            check(lib.ht_copy2d(devbuf[db].data_ptr(), nbytes, host[hb].ctypes.data, nbytes, nbytes, 1,
                                _cuda.HT_COPY_H2D, s_in.cuda_stream), "ht_copy2d")
            h2d_bytes += nbytes
            ev_h2d[hb] = s_in.record_event(torch.cuda.Event(enable_timing=True)) if timing else s_in.record_event()
            s_cmp.wait_event(ev_h2d[hb])
            e2 = s_cmp.record_event(torch.cuda.Event(enable_timing=True)) if timing else None
            if anc_labels is None:
                check(
                    lib.ht_pack_i32(devbuf[db].data_ptr(), 2 * n, n, row_in_chunk.data_ptr() + 4 * v0,
                                    dplan.var_key_off.data_ptr() + 4 * v0, dplan.key_allele.data_ptr(), c, 0,
                                    plan.n_keys, planes.data_ptr(), 0, s_cmp.cuda_stream),
                    "ht_pack_i32",
                )
            else:
                with torch.cuda.stream(s_cmp):  # allocated and filled in the compute stream's order
                    anc_t = anc_labels(cols[v0:v1])
                if anc_t.dtype != torch.uint8 or tuple(anc_t.shape[:2]) != (n, c):
                    raise ValueError("the ancestry source must return a CUDA uint8 tensor (samples, chunk variants, 2)")
                check(
                    lib.ht_pack_i32_anc(devbuf[db].data_ptr(), 2 * n, n, anc_t.data_ptr(), *map(int, anc_t.stride()),
                                        row_in_chunk.data_ptr() + 4 * v0, dplan.var_key_off.data_ptr() + 4 * v0,
                                        dplan.key_allele.data_ptr(), dplan.key_label.data_ptr(), c, 0, plan.n_keys,
                                        planes.data_ptr(), 0, s_cmp.cuda_stream),
                    "ht_pack_i32_anc",
                )
                anc_t.record_stream(s_cmp)
            ev_pack[db] = s_cmp.record_event(torch.cuda.Event(enable_timing=True)) if timing else s_cmp.record_event()
            if timing:
                ev_pairs.append((e0, ev_h2d[hb], e2, ev_pack[db]))
        t_packed = time.perf_counter()  # everything ENQUEUED; the copies and packs may still be running
        ev_packed = s_cmp.record_event(torch.cuda.Event(enable_timing=True)) if timing else None
        with torch.cuda.stream(s_cmp):
            if fmt == "bits":
                tiles = (n + engine.TILE - 1) // engine.TILE
                res = torch.empty((tiles, H, 32, 2), dtype=torch.int32, device=dev)
                engine.transform_bits(dplan, planes, n, res, s_cmp)
            else:
                pitch = engine.out_pitch(H)
                buf = torch.empty((n, pitch), dtype=torch.uint8, device=dev)
                engine.transform_u8(dplan, planes, n, buf.data_ptr(), pitch, s_cmp)
                res = buf[:, : 2 * H].unflatten(1, (H, 2))
        s_cmp.synchronize()
        cur.wait_stream(s_cmp)
        if anc_check is not None:
            anc_check()  # e.g. a variant that no breakpoint block covers: the reference's ValueError
        t_done = time.perf_counter()
        if to_host:
            if fmt == "bits":
                res = res.cpu().numpy().view(np.uint32)
            else:
                res = engine.download_result(buf, n, H, pitch, cur).view(np.bool_)
        if stats is not None:
            stats.update(h2d_s=sum(a.elapsed_time(b) for a, b, _, _ in ev_pairs) * 1e-3,
                         pack_s=sum(c.elapsed_time(d) for _, _, c, d in ev_pairs) * 1e-3)
            ev_packed.synchronize()
            # packed_s: device time from the first copy's turn on the copy stream to the last pack -- the loop above
            # only ENQUEUES, so its own end (stream_s) says nothing when the reader is fast
            stats.update(packed_s=ev_start.elapsed_time(ev_packed) * 1e-3)
            stats.update(read_s=t_read, stream_s=t_packed - t_start, total_device_s=t_done - t_start,
                         total_s=time.perf_counter() - t_start, h2d_bytes=h2d_bytes, chunks=-(-V // chunk) if V else 0)
    return res


# ----------------------------------------------------------------------------------------
# hand-off to the PGEN writer (second half)
# ----------------------------------------------------------------------------------------
def hapmajor_from_bits(bits: torch.Tensor, n_samples: int, h0: int, h1: int, dtype=torch.int32,
                       out: torch.Tensor = None, stream=None) -> torch.Tensor:
    """
    Rows h0..h1-1 of the bit-packed result `(tiles, H, 32, 2)` as the haplotype-major
    `(h1 - h0, 2 n)` matrix `PgenWriter.append_alleles_batch` takes (element 2 s + c), uint8 or
    int32, on the device (ht_unpack_bits_hapmajor).
    """
    if dtype not in (torch.int32, torch.uint8):
        raise ValueError("dtype must be torch.int32 or torch.uint8")
    H = int(bits.shape[1])
    pitch = (2 * n_samples + 3) // 4 * 4
    if out is None:
        out = torch.empty((h1 - h0, pitch), dtype=dtype, device=bits.device)
    check(
        _cuda.lib().ht_unpack_bits_hapmajor(
            bits.data_ptr(), n_samples, H, h0, h1 - h0, out.element_size(), out.data_ptr(), int(out.stride(0)),
            (stream or torch.cuda.current_stream(bits.device)).cuda_stream,
        ),
        "ht_unpack_bits_hapmajor",
    )
    return out[:, : 2 * n_samples]


def pgen_write_batches(bits: torch.Tensor, n_samples: int, chunk_haps: int = 1024, dtype=np.int32):
    """
    What `GenotypesPLINK.write` feeds `pgenlib.PgenWriter.append_alleles_batch` chunk by chunk
    (haptools/data/genotypes.py:1611-1642), produced from the transform's bit-packed result
    without ever building, transposing or widening the `(n, H, 2)` matrix on the host: yields
    `(start, end, subset_data, allele_cts)` with `subset_data` a C-contiguous `(end - start, 2 n)`
    host array of `dtype` (one of THREE page-locked buffers used in turn: batch k stays valid while
    batch k + 1 is asked for and is overwritten when batch k + 2 is -- a writer thread may hold one
    batch while the next is fetched) and `allele_cts` = 2 for every row (a 0/1 matrix never has missing calls
    and `_num_unique_alleles` floors at 2, genotypes.py:1538-1567).  The expansion of chunk
    k + 1 and its device->host copy overlap the consumer's work on chunk k.
    """
    engine.require_cuda()
    H = int(bits.shape[1])
    tdt = {np.dtype(np.int32): torch.int32, np.dtype(np.uint8): torch.uint8}[np.dtype(dtype)]
    dev = bits.device
    chunk_haps = max(1, min(chunk_haps, H))
    pitch = (2 * n_samples + 3) // 4 * 4
    with torch.cuda.device(dev):
        s_cmp, cur = torch.cuda.Stream(dev), torch.cuda.current_stream(dev)
        s_cmp.wait_stream(cur)
        dbuf = [torch.empty((chunk_haps, pitch), dtype=tdt, device=dev) for _ in range(2)]
        # three host buffers: launch(k + 1) runs BEFORE batch k is yielded and writes asynchronously, so with
        # two it would overwrite batch k - 1 while the consumer may still hold it
        hbuf = [torch.empty((chunk_haps, pitch), dtype=tdt, pin_memory=True) for _ in range(3)]
        events = [None, None, None]

        def launch(k):
            a, b = k * chunk_haps, min(H, (k + 1) * chunk_haps)
            with torch.cuda.stream(s_cmp):
                hapmajor_from_bits(bits, n_samples, a, b, tdt, out=dbuf[k % 2], stream=s_cmp)
                hbuf[k % 3][: b - a].copy_(dbuf[k % 2][: b - a], non_blocking=True)
                events[k % 3] = s_cmp.record_event()

        n_chunks = -(-H // chunk_haps) if n_samples > 0 else 0
        if n_chunks:
            launch(0)
        for k in range(n_chunks):
            if k + 1 < n_chunks:
                launch(k + 1)
            events[k % 3].synchronize()
            a, b = k * chunk_haps, min(H, (k + 1) * chunk_haps)
            host = hbuf[k % 3].numpy()[: b - a, : 2 * n_samples]
            if pitch != 2 * n_samples:
                host = np.ascontiguousarray(host)
            yield a, b, host, np.full(b - a, 2, dtype=np.uint32)
        cur.wait_stream(s_cmp)
