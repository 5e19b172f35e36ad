"""
Device execution of the haplotype transform through the C-ABI (include/haptools_cuda.h).

torch is plumbing only: device memory, streams and events.  All arithmetic happens in the
hand-written kernels of libhaptools_cuda.so; there is no CPU fallback -- without a CUDA
device (or without the built library) every entry point raises.

Two entry points:
  transform_device  genotypes already resident in HBM (a CUDA torch tensor with arbitrary
                    strides) -> result tensor on the device
  transform_host    genotypes in a host numpy array (any of the reference's layouts,
                    SURVEY.md 4c) -> numpy bool array; sample chunks are streamed with
                    double-buffered H2D / compute / D2H on three streams, the analogue of
                    the reference's variant-chunked PGEN reads
                    (haptools/data/genotypes.py:1353-1406)
"""
from __future__ import annotations

import ctypes
import os

import numpy as np
import torch

from . import _cuda
from ._cuda import HaptoolsCudaError, check
from .plan import TransformPlan, build_local_tables

TILE = _cuda.TILE_SAMPLES
# The two-phase ("local") transform is bit-exact and tested, but as measured on B200 (DESIGN.md
# section 5) it does not beat the one-kernel form yet (12.7 vs 10.1 ms on the UKB-scale plan at
# 200 k samples): method="auto" keeps to the one-kernel form unless this is switched on.
LOCAL_AUTO = False
# transform_host(result="auto"): results of this many bytes or more leave the device bit-packed and are widened
# by the host copy engine; smaller ones as bytes (one DMA, no extra kernel, no hand-off to the pool)
RESULT_BITS_MIN = 8 << 20
# The shared-memory staged transform (ht_transform_u8_staged) is bit-exact and cuts the L2 -> SM record reads of
# position-sorted haplotype sets from 3.1 x to 2.1 x the plane bytes, but on B200 it is SLOWER than the plain
# half-tile kernel on those very sets (8.6 vs 8.2 ms at 200 k samples: 3 instead of 4 CTAs per SM and a CTA-wide
# wait for the staged run): method="auto" keeps to the plain kernel unless this is set.
STAGED_AUTO = False


def require_cuda() -> None:
    if not torch.cuda.is_available():
        raise HaptoolsCudaError(
            "haptools_b200 needs a CUDA device (B200, sm_100a): the transform has no CPU fallback"
        )
    _cuda.lib()


def _ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


def _stream_ptr(stream=None) -> int:
    return (stream or torch.cuda.current_stream()).cuda_stream


class DevicePlan:
    """A TransformPlan uploaded to one device (a few MB; replicated on every rank)."""

    def __init__(self, plan: TransformPlan, device=None, local_dmax: int = None, local_max_block: int = None):
        require_cuda()
        self.plan = plan
        self._local_kw = {k: v for k, v in (("dmax", local_dmax), ("max_block", local_max_block)) if v}
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)

        def up(a, dtype):
            # uint32 tables are carried as int32 tensors (same bits)
            return None if a is None else torch.from_numpy(np.ascontiguousarray(a).view(dtype)).to(self.device)

        self.var_col = up(plan.var_col, np.int32)
        self.var_key_off = up(plan.var_key_off, np.int32)
        self.key_allele = up(plan.key_allele, np.uint8)
        self.key_label = up(plan.key_label, np.uint8)
        self.hap_off = up(plan.hap_off, np.int32)
        self.hap_key = up(plan.hap_key, np.int32)
        quad_off16, quads = plan.quad_tables()
        self.quad_off16 = up(quad_off16, np.int32)
        self.quads = up(quads, np.int32)
        self.col_var = up(plan.col_var(), np.int32)
        col_key01, col_key_range, var_other = plan.fast_tables()
        self.col_key01 = up(col_key01, np.int32)
        self.col_key_range = up(col_key_range, np.int32)
        self.var_other = up(var_other, np.int32)
        self.tables = _cuda.PackTables(
            plan.n_cols, plan.n_vars, plan.n_keys, _ptr(self.var_col), _ptr(self.var_key_off),
            _ptr(self.key_allele), _ptr(self.key_label), _ptr(self.col_var), _ptr(self.col_key01),
            _ptr(self.col_key_range), _ptr(self.var_other) if len(var_other) else None, len(var_other),
        )

        # key runs of the 128-haplotype blocks, for the shared-memory staged transform (0 lines: does not pay)
        env = os.environ.get("HT_STAGE_LINES", "").strip()
        self.stage_lines = int(env) if env else plan.stage_lines()
        self.blk_key_lo = self.blk_key_cnt = None
        if self.stage_lines > 0 and quads is not None:
            lo, cnt = plan.block_spans()
            self.blk_key_lo, self.blk_key_cnt = up(lo, np.int32), up(cnt, np.int32)
        else:
            self.stage_lines = 0
        self._local = None  # (LocalTables, device tensors, ctypes struct), built on first use
        self._compact = None  # tables for a genotype buffer that holds only the referenced variants
        self._up = up

    def compact_tables(self):
        """Pack tables for a device buffer whose "columns" are just the plan's referenced variants, in
        var_col order (what transform_host uploads from a variant-major host array when few of its
        variants are referred to): variant v sits at column v.  Only the kernels that go through var_col
        (variant-major, generic) apply, which is what that layout takes anyway."""
        if self._compact is None:
            ident = torch.arange(self.plan.n_vars, dtype=torch.int32, device=self.device)
            tables = _cuda.PackTables(
                self.plan.n_vars, self.plan.n_vars, self.plan.n_keys, _ptr(ident), _ptr(self.var_key_off),
                _ptr(self.key_allele), _ptr(self.key_label), None, None, None, None, -1,
            )
            self._compact = (ident, tables)
        return self._compact[1]

    n_haps = property(lambda self: self.plan.n_haps)
    n_keys = property(lambda self: self.plan.n_keys)
    n_vars = property(lambda self: self.plan.n_vars)

    def local(self):
        """Locality tables of the two-phase transform (plan.build_local_tables), uploaded once:
        returns (LocalTables, _cuda.LocalTables)."""
        if self._local is None:
            lt = build_local_tables(self.plan, **self._local_kw)
            dev = [self._up(a, np.int32) for a in
                   (lt.blk_hap_off, lt.blk_key_lo, lt.blk_key_cnt, lt.s_hap_off, lt.s_hap_key, lt.order)]
            struct = _cuda.LocalTables(lt.n_blocks, lt.dmax, *[_ptr(t) if t.numel() else None for t in dev])
            self._local = (lt, dev, struct)
        return self._local[0], self._local[2]

    def use_local(self) -> bool:
        """Should "auto" pick the two-phase transform?  Only when it is enabled (LOCAL_AUTO) and
        the plan shares enough alleles between neighbouring haplotypes (LocalTables.worthwhile)."""
        if not LOCAL_AUTO:
            return False
        if self.plan.n_haps == 0 or self.plan.n_refs < 4 * self.plan.n_haps:
            return False  # cannot pay for the packed intermediate: skip building the tables
        return self.local()[0].worthwhile(self.plan.n_refs)


def device_plan(plan: TransformPlan, device) -> DevicePlan:
    """The plan's tables on `device`, uploaded once per (plan object, device) and kept with the plan."""
    device = torch.device(device)
    cache = plan.__dict__.setdefault("_device_plans", {})
    key = (device.type, device.index)
    if key not in cache:
        cache[key] = DevicePlan(plan, device)
    return cache[key]


def plane_bytes(n_samples: int, n_keys: int) -> int:
    return int(_cuda.lib().ht_plane_bytes(n_samples, n_keys))


def alloc_planes(n_samples: int, n_keys: int, device) -> torch.Tensor:
    return torch.empty(max(plane_bytes(n_samples, n_keys), 16) // 4, dtype=torch.int32, device=device)


def out_pitch(n_haps: int) -> int:
    """Row pitch (bytes) of a device result buffer: 16-byte aligned rows keep the transform
    kernel on its 128-bit store path for any number of haplotypes."""
    return max(16, (2 * n_haps + 15) // 16 * 16)


def pack(dplan: DevicePlan, gt_ptr: int, gt_strides, n_samples: int, planes: torch.Tensor,
         anc_ptr: int = 0, anc_strides=(0, 0, 0), flags: torch.Tensor = None,
         mode: int = _cuda.HT_PACK_AUTO, stream=None, tables=None) -> None:
    """K1 on raw device pointers; strides are (sample, variant, strand) in bytes.  `tables`:
    DevicePlan.compact_tables() when the buffer holds only the referenced variants."""
    if dplan.plan.has_ancestry != bool(anc_ptr):
        raise ValueError("ancestry plan and ancestry array must be given together")
    check(
        _cuda.lib().ht_pack_u8(
            gt_ptr, *map(int, gt_strides), anc_ptr, *map(int, anc_strides), n_samples,
            ctypes.byref(tables if tables is not None else dplan.tables), _ptr(planes), _ptr(flags), mode, _stream_ptr(stream),
        ),
        "ht_pack_u8",
    )


def local_workspace(dplan: DevicePlan, n_samples: int, device) -> torch.Tensor:
    """Scratch of the two-phase transform: the bit-packed result of one group of tiles."""
    nbytes = int(_cuda.lib().ht_local_workspace_bytes(n_samples, dplan.n_haps))
    return torch.empty(max(nbytes, 16) // 4, dtype=torch.int32, device=device)


def transform_u8(dplan: DevicePlan, planes: torch.Tensor, n_samples: int, out_ptr: int, pitch: int,
                 stream=None, counts: torch.Tensor = None, method: str = "auto",
                 workspace: torch.Tensor = None, use_quads: bool = True) -> None:
    """K2 on raw device pointers.  `counts` (optional int64[H] tensor) receives the number of
    present copies of every haplotype (the numerator of check_maf).

    method  "direct": one kernel, one L2 line read per haplotype allele (ht_transform_u8);
            "staged": the same kernel with the key runs of its haplotype blocks staged in shared memory
            (ht_transform_u8_staged; pays for position-sorted haplotype sets, DevicePlan.stage_lines);
            "local": two phases through shared memory and a bit-packed intermediate
            (ht_transform_u8_local); "auto": direct -- the other two are bit-exact but measured slower on
            B200 (STAGED_AUTO / LOCAL_AUTO switch them on for plans they suit).  Results are identical.
    use_quads  False: let the kernel build its padded key quads from the CSR itself (the form
            without planner support; kept for callers that only have hap_off / hap_key)."""
    if method not in ("auto", "direct", "local", "staged"):
        raise ValueError("method must be 'auto', 'direct', 'staged' or 'local'")
    if (method == "staged" or (method == "auto" and STAGED_AUTO)) and use_quads and dplan.stage_lines > 0:
        check(
            _cuda.lib().ht_transform_u8_staged(
                _ptr(planes), n_samples, dplan.n_keys, _ptr(dplan.hap_off), _ptr(dplan.hap_key),
                _ptr(dplan.quad_off16), _ptr(dplan.quads), _ptr(dplan.blk_key_lo), _ptr(dplan.blk_key_cnt),
                dplan.stage_lines, dplan.n_haps, out_ptr, pitch, _ptr(counts), _stream_ptr(stream),
            ),
            "ht_transform_u8_staged",
        )
        return
    if method == "local" or (method == "auto" and dplan.use_local()):
        if dplan.n_haps == 0 or n_samples == 0:
            return
        _, tables = dplan.local()
        if workspace is None:
            # allocated on the launching stream: the caching allocator then reuses the block
            # only in that stream's order
            with torch.cuda.stream(stream or torch.cuda.current_stream(dplan.device)):
                workspace = local_workspace(dplan, n_samples, dplan.device)
        check(
            _cuda.lib().ht_transform_u8_local(
                _ptr(planes), n_samples, dplan.n_keys, ctypes.byref(tables), dplan.n_haps, out_ptr, pitch,
                _ptr(counts), _ptr(workspace), workspace.numel() * workspace.element_size(), _stream_ptr(stream),
            ),
            "ht_transform_u8_local",
        )
        return
    check(
        _cuda.lib().ht_transform_u8(
            _ptr(planes), n_samples, dplan.n_keys, _ptr(dplan.hap_off), _ptr(dplan.hap_key),
            _ptr(dplan.quad_off16) if use_quads else None, _ptr(dplan.quads) if use_quads else None,
            dplan.n_haps, out_ptr, pitch, _ptr(counts), _stream_ptr(stream),
        ),
        "ht_transform_u8",
    )


def transform_bits(dplan: DevicePlan, planes: torch.Tensor, n_samples: int, out_bits: torch.Tensor,
                   stream=None, method: str = "auto", counts: torch.Tensor = None) -> None:
    """Bit-packed result (tiles, H, 32, 2); `method` as in transform_u8 (`counts` needs "local")."""
    if method == "local" or (method == "auto" and dplan.n_haps and dplan.use_local()):
        if dplan.n_haps == 0 or n_samples == 0:
            return
        _, tables = dplan.local()
        check(
            _cuda.lib().ht_transform_bits_local(
                _ptr(planes), n_samples, dplan.n_keys, ctypes.byref(tables), dplan.n_haps, _ptr(out_bits),
                _ptr(counts), _stream_ptr(stream),
            ),
            "ht_transform_bits_local",
        )
        return
    if counts is not None:
        raise ValueError("counts are only produced by the local form; use count_bits")
    check(
        _cuda.lib().ht_transform_bits(
            _ptr(planes), n_samples, dplan.n_keys, _ptr(dplan.hap_off), _ptr(dplan.hap_key),
            dplan.n_haps, _ptr(out_bits), _stream_ptr(stream),
        ),
        "ht_transform_bits",
    )


def unpack_bits_u8(bits: torch.Tensor, n_samples: int, n_haps: int, out_ptr: int, pitch: int, stream=None) -> None:
    """Bit-packed result -> (n, H, 2) bytes (ht_unpack_bits_u8)."""
    check(_cuda.lib().ht_unpack_bits_u8(_ptr(bits), n_samples, n_haps, out_ptr, pitch, _stream_ptr(stream)), "ht_unpack_bits_u8")


def count_bits(bits: torch.Tensor, n_samples: int, n_haps: int, stream=None) -> torch.Tensor:
    """Per-haplotype allele counts of a bit-packed result (numerator of check_maf,
    haptools/data/genotypes.py:559-625)."""
    counts = torch.empty(n_haps, dtype=torch.int64, device=bits.device)
    check(_cuda.lib().ht_count_bits(_ptr(bits), n_samples, n_haps, _ptr(counts), _stream_ptr(stream)), "ht_count_bits")
    return counts


def _check_gt_tensor(gt: torch.Tensor, what: str) -> None:
    if not isinstance(gt, torch.Tensor) or not gt.is_cuda:
        raise TypeError(f"{what} must be a CUDA tensor")
    if gt.dtype not in (torch.uint8, torch.bool) or gt.dim() != 3 or gt.shape[2] < 2:
        raise ValueError(f"{what} must have shape (samples, variants, >=2) and dtype uint8 or bool")


def maf_from_counts(counts, n_samples: int) -> np.ndarray:
    """Minor allele frequency of each haplotype from its allele count, exactly as
    Genotypes.check_maf computes it (haptools/data/genotypes.py:600-602)."""
    counts = counts.cpu().numpy() if isinstance(counts, torch.Tensor) else np.asarray(counts)
    ref_af = counts / (2 * n_samples)
    return np.array([ref_af, 1 - ref_af]).min(axis=0)


def transform_device(gt: torch.Tensor, dplan: DevicePlan, ancestry: torch.Tensor = None,
                     out: torch.Tensor = None, planes: torch.Tensor = None, fmt: str = "u8",
                     flags: torch.Tensor = None, pack_mode: int = _cuda.HT_PACK_AUTO,
                     counts: torch.Tensor = None):
    """
    Transform genotypes resident on the device.

    gt        (n, p, >=2) uint8/bool CUDA tensor, any strides (never copied)
    ancestry  (n, p, 2) uint8 CUDA tensor for ancestry plans
    fmt       "u8": returns an (n, H, 2) uint8 tensor (0/1 bytes == numpy bool bytes);
              "bits": returns the bit-packed result (tiles, H, 32, 2) int32
    """
    require_cuda()
    _check_gt_tensor(gt, "gt")
    n, p = int(gt.shape[0]), int(gt.shape[1])
    if p != dplan.plan.n_cols:
        raise ValueError(f"plan was made for {dplan.plan.n_cols} variant columns, genotypes have {p}")
    if ancestry is not None:
        _check_gt_tensor(ancestry, "ancestry")
        if tuple(ancestry.shape[:2]) != (n, p):
            raise ValueError("ancestry and genotypes differ in shape")
    H = dplan.n_haps
    with torch.cuda.device(gt.device):
        if planes is None:
            planes = alloc_planes(n, dplan.n_keys, gt.device)
        pack(dplan, gt.data_ptr(), gt.stride(), n, planes,
             _ptr(ancestry), ancestry.stride() if ancestry is not None else (0, 0, 0), flags, pack_mode)
        if fmt == "bits":
            tiles = (n + TILE - 1) // TILE
            if out is None:
                out = torch.empty((tiles, H, 32, 2), dtype=torch.int32, device=gt.device)
            transform_bits(dplan, planes, n, out)
            return out
        if fmt != "u8":
            raise ValueError("fmt must be 'u8' or 'bits'")
        if out is None:
            pitch = out_pitch(H)
            buf = torch.empty((n, pitch), dtype=torch.uint8, device=gt.device)
            out = buf[:, : 2 * H].view(n, H, 2) if pitch == 2 * H else buf[:, : 2 * H].unflatten(1, (H, 2))
        else:
            if out.dtype != torch.uint8 or tuple(out.shape) != (n, H, 2) or (H and out.stride()[1:] != (2, 1)):
                raise ValueError("out must be a uint8 tensor of shape (n, H, 2) with contiguous rows")
        transform_u8(dplan, planes, n, out.data_ptr(), int(out.stride(0)) if n else 2 * H, counts=counts)
        return out


# --------------------------------------------------------------------------------------
# host arrays
# --------------------------------------------------------------------------------------
def _fast_layout(strides) -> bool:
    """Do these device strides (sample, variant, strand) reach one of the fast pack kernels?"""
    ss, sv, sc = strides
    return sc == 1 and ((sv == 2 and ss % 16 == 0) or (ss == 2 and sv % 16 == 0))


class _Region:
    """The bytes of a[s0:s1, :, :2] as a pitched 2-D region of host memory."""

    __slots__ = ("ptr", "rows", "width", "pitch", "sample_major", "ss", "sv", "sc", "keep")

    def dev_pitch(self) -> int:
        return max(16, (self.width + 15) // 16 * 16)

    def dev_strides(self):
        dp = self.dev_pitch()
        return (dp, self.sv, self.sc) if self.sample_major else (self.ss, dp, self.sc)

    def dev_bytes(self) -> int:
        return self.rows * self.dev_pitch()

    def dense(self, n: int, p: int):
        """(strides, bytes) of the dense 2-bytes-per-genotype re-layout with the same major
        axis and 16-byte aligned rows (ht_relayout_u8 target)."""
        inner = p if self.sample_major else n
        pitch = max(16, (2 * inner + 15) // 16 * 16)
        strides = (pitch, 2, 1) if self.sample_major else (2, pitch, 1)
        return strides, pitch * (n if self.sample_major else p)


def _region(a: np.ndarray, s0: int, s1: int, depth: int = 2) -> _Region:
    """`depth`: strand indices covered (3 = with the phase byte, for the QC pass)."""
    n, p = s1 - s0, a.shape[1]
    ss, sv, sc = a.strides
    r = _Region()
    r.keep = None
    ok = ss > 0 and sv > 0 and sc > 0
    if ok and ss >= sv:
        width = (p - 1) * sv + (depth - 1) * sc + 1
        ok = width <= ss or n <= 1
        sample_major = True
    elif ok:
        width = (n - 1) * ss + (depth - 1) * sc + 1
        ok = width <= sv or p <= 1
        sample_major = False
    if not ok:
        # exotic views (negative or overlapping strides): one contiguous host copy
        c = np.ascontiguousarray(a[s0:s1, :, :depth])
        r.keep = c
        a, s0 = c, 0
        ss, sv, sc = c.strides
        width, sample_major = (p - 1) * sv + (depth - 1) * sc + 1, True
    r.ptr = a.ctypes.data + s0 * ss
    r.sample_major = sample_major
    r.ss, r.sv, r.sc = ss, sv, sc
    r.width = width
    r.rows, r.pitch = (n, ss) if sample_major else (p, sv)
    return r


def _copy2d(dst: int, dpitch: int, src: int, spitch: int, width: int, rows: int, kind: int, stream) -> None:
    check(_cuda.lib().ht_copy2d(dst, dpitch, src, spitch, width, rows, kind, stream.cuda_stream), "ht_copy2d")


def default_chunk(n: int, p: int, n_haps: int, target_bytes: int = 512 << 20) -> int:
    per_sample = max(2 * n_haps, 3 * p, 1)
    rows = max(TILE, target_bytes // per_sample // TILE * TILE)
    return int(min(max(n, 1), rows))


def _ramped_bounds(n: int, chunk: int, n_devs: int = 1):
    """Sample ranges of at most `chunk` samples, the first ones of every device short (chunk/8, chunk/4,
    chunk/2): the result's D2H stream -- the bound of the host pipeline -- can only start when the first
    chunk has been uploaded and transformed, so the lead-in is one SHORT upload instead of a full one."""
    bounds, s = [], 0
    for div in (8, 4, 2):
        size = max(256, chunk // div // 256 * 256)
        for _ in range(n_devs):
            if s < n and n - s > size:
                bounds.append((s, s + size))
                s += size
    while s < n:
        bounds.append((s, min(n, s + chunk)))
        s += chunk
    return bounds


def _half_chunk(n: int) -> int:
    """Chunk size that cuts n samples into two pieces: half of them, rounded up to a multiple of 256."""
    return -(-(-(-n // 2)) // 256) * 256


# Result arrays of transform_host(out=None): like the reference's np.empty (haplotypes.py:1408), except that the
# memory of a result the caller has DROPPED is handed out again.  A fresh multi-GB np.empty is first touched
# page by page while it is filled (the kernel zeroes every page: 30 ms of the 72 ms a 3.3 GB result takes), and
# glibc gives such blocks back to the kernel on free, so a loop of transforms pays that every time.  An owner
# array is idle when nothing but this list refers to it (every view holds a reference to its base); at most one
# idle buffer is kept.  HAPTOOLS_B200_RESULT_POOL=0 switches the recycling off.
_result_owners: list = []


def _result_array(shape) -> np.ndarray:
    import sys

    nbytes = int(np.prod(shape))
    if os.environ.get("HAPTOOLS_B200_RESULT_POOL", "1").strip() in ("0", "off", "no") or nbytes < (64 << 20):
        return np.empty(shape, dtype=np.uint8)
    pick = None
    for i in range(len(_result_owners) - 1, -1, -1):
        if sys.getrefcount(_result_owners[i]) > 2:  # the list + getrefcount's argument: views are alive
            continue
        owner = _result_owners.pop(i)
        if pick is None and nbytes <= owner.nbytes <= 2 * nbytes:
            pick = owner
        del owner  # idle buffers that do not fit are released
    if pick is None:
        pick = np.empty(nbytes, dtype=np.uint8)
    _result_owners.append(pick)
    return pick[:nbytes].reshape(shape)


def pinned_empty(shape, dtype=np.uint8) -> np.ndarray:
    """A page-locked numpy array (owned by a torch pinned tensor kept alive through .base)."""
    require_cuda()
    nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
    t = torch.empty(max(nbytes, 1), dtype=torch.uint8, pin_memory=True)
    return t.numpy()[:nbytes].view(dtype).reshape(shape)


def _submit(ticket: int, what: str) -> int:
    if ticket < 0:
        raise HaptoolsCudaError(f"{what}: {_cuda.lib().ht_last_error().decode('utf-8', 'replace')}")
    return ticket


def _host_wait(ticket) -> None:
    if ticket:
        check(_cuda.lib().ht_host_wait(ticket), "ht_host_wait")


def _pageable(ptr: int, nbytes: int) -> bool:
    """Ordinary host memory of a size worth the host copy engine (pool of threads + page-locked slots)?"""
    return nbytes >= (4 << 20) and bool(_cuda.lib().ht_host_is_pageable(ptr))


class _HostPipe:
    """One device's share of transform_host: its chunks flow through upload -> kernels -> download, two of
    everything, each stage asynchronous to the one Python thread that drives all devices.

    upload    page-locked source: a pitched DMA (or the row-gather kernel) on the copy stream; PAGEABLE source:
              a job of the host copy engine (ht_host_submit_upload), waited for just before the chunk's kernels
    kernels   [re-layout] -> pack -> transform on the compute stream
    download  result="bits" (default for results of 8 MB or more): ht_transform_bits -> ht_bits_to_rowbits, and
              the host copy engine brings 1 bit per hap-call over PCIe and widens it into the caller's array
              (ht_host_submit_download_expand) -- page-locked or not; result="u8": the bytes themselves, by
              DMA into a page-locked array or through the engine's slots into a pageable one
    """

    def __init__(self, dev, plan, dplan, gt, ancestry, bounds, chunk, pack_mode, out, H, result):
        self.dev, self.plan, self.pack_mode, self.H = dev, plan, pack_mode, H
        self.gt, self.ancestry, self.out_ptr = gt, ancestry, out.ctypes.data
        self.bounds = bounds
        self.p = gt.shape[1]
        self.result = result
        lib = _cuda.lib()
        with torch.cuda.device(dev):
            self.dplan = dplan if dplan is not None and dplan.device == dev else device_plan(plan, dev)
            self.pitch = out_pitch(H)
            ends = sorted(bounds, key=lambda b: b[0] - b[1])[:2] + bounds[-1:]  # the largest chunks (and the last)
            # Variant-major host array (VCF reader layout: a variant is a contiguous row) of which the plan
            # refers to at most half the variants: upload only those rows (ht_copy_rows_h2d) -- the device-side
            # counterpart of the reference's gts.subset(variants=...) gather (haplotypes.py:1388)
            r0 = _region(gt, *bounds[0])
            self.gather = (not r0.sample_major and r0.keep is None and 0 < 2 * plan.n_vars <= self.p
                           and r0.width <= (8 << 20))
            if self.gather and plan.has_ancestry:
                # the row gather needs the ancestry array in the same (variant-major) layout; any other layout
                # takes the full upload, like the reference accepts any layout
                ra0 = _region(ancestry, *bounds[0])
                self.gather = not (ra0.sample_major or ra0.keep is not None or ra0.width > (8 << 20))
            self.rows_idx = np.ascontiguousarray(plan.var_col, dtype=np.uint32) if self.gather else None
            self.pv = plan.n_vars if self.gather else self.p  # variant rows held by the device buffers
            self.tables = self.dplan.compact_tables() if self.gather else None
            if self.gather:
                gt_bytes = max(plan.n_vars * _region(gt, *b).dev_pitch() for b in ends)
            else:
                gt_bytes = max(_region(gt, *b).dev_bytes() for b in ends)
            anc_bytes = 0
            if plan.has_ancestry:
                if self.gather:
                    anc_bytes = max(plan.n_vars * _region(ancestry, *b).dev_pitch() for b in ends)
                else:
                    anc_bytes = max(_region(ancestry, *b).dev_bytes() for b in ends)
            if result == "bits":
                self.bits_pitch = int(lib.ht_rowbits_pitch(H))
                out_bytes = chunk * self.bits_pitch
            else:
                out_bytes = chunk * self.pitch
            # Buffers per stage.  Two are what the DEVICE needs (upload k+1 || kernels k || download k-1).  More of
            # them take the host waits out of the loop that queues the chunks: with a buffer set per chunk nothing
            # is reused, so the one Python thread queues the whole call in a few ms and then sits in finish() with
            # the GIL released -- whatever else runs in the interpreter (the class API's helper thread) no longer
            # delays a chunk.  Bounded by a quarter of the free device memory and 16 GB (HAPTOOLS_B200_PIPE_DEPTH
            # overrides): enough for 128 k samples of the UKB-scale plan per call.
            depth = os.environ.get("HAPTOOLS_B200_PIPE_DEPTH", "").strip()
            if depth:
                want = max(2, int(depth))
            else:
                free = torch.cuda.mem_get_info(dev)[0]
                want = max(2, int(min(free // 4, 16 << 30) // max(gt_bytes + anc_bytes + out_bytes, 1)))
            nbuf = min(want, len(bounds))
            while True:
                try:
                    self.gt_buf = [torch.empty(gt_bytes, dtype=torch.uint8, device=dev) for _ in range(nbuf)]
                    self.anc_buf = [None] * nbuf
                    if plan.has_ancestry:
                        self.anc_buf = [torch.empty(anc_bytes, dtype=torch.uint8, device=dev) for _ in range(nbuf)]
                    self.out_buf = [torch.empty(out_bytes, dtype=torch.uint8, device=dev) for _ in range(nbuf)]
                    break
                except torch.cuda.OutOfMemoryError:
                    # the extra sets are a convenience of the host side: fall back to what the device needs
                    self.gt_buf = self.anc_buf = self.out_buf = None
                    if nbuf <= 2:
                        raise
                    nbuf = 2
                    torch.cuda.empty_cache()
            self.nbuf = nbuf
            self.queues_everything = nbuf >= len(bounds)
            if result == "bits":
                tiles = (chunk + TILE - 1) // TILE
                self.bits = torch.empty((tiles, H, 32, 2), dtype=torch.int32, device=dev)
            self.out_pageable = _pageable(self.out_ptr, out.nbytes)
            self.src_pageable = _pageable(r0.ptr, r0.width * r0.rows)  # its uploads are host copy engine jobs the loop waits for
            self.planes = alloc_planes(chunk, plan.n_keys, dev)
            self.s_in, self.s_cmp, self.s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)
            self.cur = torch.cuda.current_stream(dev)
            for st in (self.s_in, self.s_cmp, self.s_out):
                st.wait_stream(self.cur)
        self.dense_gt = self.dense_anc = None  # scratch of the on-device re-layout (compute stream only)
        self.ev_h2d, self.ev_cmp, self.ev_d2h = [None] * nbuf, [None] * nbuf, [None] * nbuf
        self.up_ticket = [[] for _ in range(nbuf)]  # host copy engine jobs of the chunk being uploaded into buffer b
        self.dl_ticket = [0] * nbuf                 # ... and of the result leaving buffer b
        self.regions = [None] * nbuf
        self.keep = []
        self.k = 0
        self.uploads_queued = False

    def step(self) -> None:
        """Queue this device's next chunk; everything is asynchronous except the waits for host copy engine
        jobs whose buffers are needed again."""
        k = self.k
        self.k += 1
        if not self.uploads_queued:
            if k == 0:
                self._start_upload(0)
            if k + 1 < len(self.bounds):
                self._start_upload(k + 1)  # flies while chunk k's kernels are queued
        self._compute(k)

    def queue_all_uploads(self) -> None:
        """With a buffer set per chunk every upload can be queued before the first kernel: the copy stream (or the
        host copy engine, for a pageable source) then never waits for this thread to come back from a wait."""
        if self.queues_everything and not self.uploads_queued:
            for k in range(len(self.bounds)):
                self._start_upload(k)
            self.uploads_queued = True

    def _start_upload(self, k: int) -> None:
        b = k % self.nbuf
        s0, s1 = self.bounds[k]
        with torch.cuda.device(self.dev):
            # the previous user of this input buffer must have been packed
            if self.ev_cmp[b] is not None:
                self.s_in.wait_event(self.ev_cmp[b])
            rg = _region(self.gt, s0, s1)
            ra = _region(self.ancestry, s0, s1) if self.plan.has_ancestry else None
            self.keep += [rg.keep, ra.keep if ra else None]
            self.up_ticket[b] = [self._upload(self.gt_buf[b], rg)]
            if ra is not None:
                self.up_ticket[b].append(self._upload(self.anc_buf[b], ra))
            self.ev_h2d[b] = self.s_in.record_event()
            self.regions[b] = (rg, ra)

    def _upload(self, buf: torch.Tensor, reg: "_Region") -> int:
        """Host region -> device buffer with 16-byte aligned rows; in gather mode only the referenced variants.
        Returns the host copy engine's ticket (0: the copy is in stream order on s_in)."""
        lib, st = _cuda.lib(), self.s_in.cuda_stream
        pageable = _pageable(reg.ptr, reg.width * reg.rows) and reg.width <= (8 << 20)
        if self.gather:
            if pageable:
                return _submit(lib.ht_host_submit_upload(buf.data_ptr(), reg.dev_pitch(), reg.ptr, reg.pitch, reg.width,
                                                         len(self.rows_idx), self.rows_idx.ctypes.data, st),
                               "ht_host_submit_upload")
            check(lib.ht_copy_rows_h2d(buf.data_ptr(), reg.dev_pitch(), reg.ptr, reg.pitch, reg.width,
                                       self.rows_idx.ctypes.data, len(self.rows_idx), st), "ht_copy_rows_h2d")
            return 0
        if pageable:
            return _submit(lib.ht_host_submit_upload(buf.data_ptr(), reg.dev_pitch(), reg.ptr, reg.pitch, reg.width,
                                                     reg.rows, None, st), "ht_host_submit_upload")
        _copy2d(buf.data_ptr(), reg.dev_pitch(), reg.ptr, reg.pitch, reg.width, reg.rows, _cuda.HT_COPY_H2D, self.s_in)
        return 0

    def _compute(self, k: int) -> None:
        plan, dev, H = self.plan, self.dev, self.H
        b = k % self.nbuf
        s0, s1 = self.bounds[k]
        rows = s1 - s0
        s_cmp = self.s_cmp
        lib = _cuda.lib()
        with torch.cuda.device(dev):
            for t in self.up_ticket[b]:  # pageable source: its data is in device memory when the job is done
                _host_wait(t)
            self.up_ticket[b] = []
            s_cmp.wait_event(self.ev_h2d[b])
            # the previous result in this output buffer must have left
            _host_wait(self.dl_ticket[b])
            self.dl_ticket[b] = 0
            if self.ev_d2h[b] is not None:
                s_cmp.wait_event(self.ev_d2h[b])
            rg, ra = self.regions[b]
            g_ptr, g_str = self.gt_buf[b].data_ptr(), rg.dev_strides()
            a_ptr, a_str = (_ptr(self.anc_buf[b]), ra.dev_strides()) if ra else (0, (0, 0, 0))
            if self.pack_mode == _cuda.HT_PACK_AUTO:
                # arrays that interleave the phase column (or other odd strides): dense re-layout
                # on the device so that the fast pack kernels apply (PCIe is the bound here)
                if not _fast_layout(g_str):
                    d_str, d_bytes = rg.dense(rows, self.pv)
                    if self.dense_gt is None or self.dense_gt.numel() < d_bytes:
                        with torch.cuda.stream(s_cmp):
                            self.dense_gt = torch.empty(d_bytes, dtype=torch.uint8, device=dev)
                    check(lib.ht_relayout_u8(g_ptr, *map(int, g_str), rows, self.pv, self.dense_gt.data_ptr(),
                                             d_str[0], d_str[1], s_cmp.cuda_stream), "ht_relayout_u8")
                    g_ptr, g_str = self.dense_gt.data_ptr(), d_str
                if ra and not _fast_layout(a_str):
                    d_str, d_bytes = ra.dense(rows, self.pv)
                    if self.dense_anc is None or self.dense_anc.numel() < d_bytes:
                        with torch.cuda.stream(s_cmp):
                            self.dense_anc = torch.empty(d_bytes, dtype=torch.uint8, device=dev)
                    check(lib.ht_relayout_u8(a_ptr, *map(int, a_str), rows, self.pv, self.dense_anc.data_ptr(),
                                             d_str[0], d_str[1], s_cmp.cuda_stream), "ht_relayout_u8")
                    a_ptr, a_str = self.dense_anc.data_ptr(), d_str
            pack(self.dplan, g_ptr, g_str, rows, self.planes, a_ptr, a_str, None, self.pack_mode, s_cmp, tables=self.tables)
            dst = self.out_ptr + s0 * 2 * H
            if self.result == "bits":
                transform_bits(self.dplan, self.planes, rows, self.bits, s_cmp)
                check(lib.ht_bits_to_rowbits(self.bits.data_ptr(), rows, H, self.out_buf[b].data_ptr(), self.bits_pitch,
                                             s_cmp.cuda_stream), "ht_bits_to_rowbits")
                self.ev_cmp[b] = s_cmp.record_event()
                self.dl_ticket[b] = _submit(
                    lib.ht_host_submit_download_expand(dst, 2 * H, 2 * H, self.out_buf[b].data_ptr(), self.bits_pitch, rows,
                                                       s_cmp.cuda_stream), "ht_host_submit_download_expand")
                return
            transform_u8(self.dplan, self.planes, rows, self.out_buf[b].data_ptr(), self.pitch, s_cmp)
            self.ev_cmp[b] = s_cmp.record_event()
            if self.out_pageable and 2 * H <= (8 << 20):
                self.dl_ticket[b] = _submit(
                    lib.ht_host_submit_download(dst, 2 * H, self.out_buf[b].data_ptr(), self.pitch, 2 * H, rows,
                                                s_cmp.cuda_stream), "ht_host_submit_download")
                return
            self.s_out.wait_event(self.ev_cmp[b])
            _copy2d(dst, 2 * H, self.out_buf[b].data_ptr(), self.pitch, 2 * H, rows, _cuda.HT_COPY_D2H, self.s_out)
            self.ev_d2h[b] = self.s_out.record_event()

    def finish(self) -> None:
        for b in range(self.nbuf):
            _host_wait(self.dl_ticket[b])
            self.dl_ticket[b] = 0
        with torch.cuda.device(self.dev):
            self.s_out.synchronize()
            self.s_cmp.synchronize()
            self.s_in.synchronize()
            self.cur.wait_stream(self.s_out)


def download_result(buf: torch.Tensor, n: int, H: int, pitch: int, stream=None) -> np.ndarray:
    """A device result `(n, pitch)` uint8 (0/1 bytes, 2H used per row) -> a host `(n, H, 2)` uint8 array, by host
    copy engine jobs of at most 256 MB each (page-locked slots, several host threads, asynchronous DMA) -- all
    submitted at once, so the blocks of one job overlap the staging of the others."""
    lib = _cuda.lib()
    out = _result_array((n, H, 2))
    if n == 0 or H == 0:
        return out
    st = (stream or torch.cuda.current_stream(buf.device)).cuda_stream
    if 2 * H > (8 << 20) or not _pageable(out.ctypes.data, out.nbytes):
        _copy2d(out.ctypes.data, 2 * H, buf.data_ptr(), pitch, 2 * H, n, _cuda.HT_COPY_D2H, stream or torch.cuda.current_stream(buf.device))
        (stream or torch.cuda.current_stream(buf.device)).synchronize()
        return out
    rows = max(1, (256 << 20) // max(2 * H, 1))
    tickets = [_submit(lib.ht_host_submit_download(out.ctypes.data + s0 * 2 * H, 2 * H, buf.data_ptr() + s0 * pitch, pitch,
                                                   2 * H, min(rows, n - s0), st), "ht_host_submit_download")
               for s0 in range(0, n, rows)]
    for t in tickets:
        _host_wait(t)
    return out


def _gatherable(a: np.ndarray, plan: TransformPlan) -> bool:
    """Variant-major host array (a variant is a contiguous row of at most 8 MB) of which the plan refers to at
    most half the variants: only the referenced rows need to be uploaded."""
    if a.shape[0] == 0:
        return False
    r = _region(a, 0, min(a.shape[0], TILE))
    return not r.sample_major and r.keep is None and 0 < 2 * plan.n_vars <= a.shape[1] and \
        ((a.shape[0] - 1) * r.ss + r.sc + 1) <= (8 << 20)


def upload_bytes(gt: np.ndarray, plan: TransformPlan, ancestry: np.ndarray = None) -> int:
    """Bytes transform_host copies to the device for these arrays: the bytes a[:, :, :2] spans, or -- for a
    variant-major array of which the plan refers to at most half the variants -- only the referenced rows."""
    total = 0
    gather = _gatherable(gt, plan) and (ancestry is None or _gatherable(ancestry, plan))
    for a in (gt, ancestry):
        if a is None or a.shape[0] == 0:
            continue
        r = _region(a, 0, a.shape[0])
        total += r.width * (plan.n_vars if gather else r.rows)
    return int(total)


def transform_host(gt: np.ndarray, plan: TransformPlan, ancestry: np.ndarray = None, device=None,
                   chunk_samples: int = None, out: np.ndarray = None, pin_output: bool = None,
                   dplan: DevicePlan = None, pack_mode: int = _cuda.HT_PACK_AUTO, devices=None,
                   result: str = "auto", on_queued=None) -> np.ndarray:
    """
    gt (n, p, 2|3) uint8/bool host array (a view is fine: VCF-style transposed, PGEN-style
    with a phase column, C-contiguous ...) -> (n, H, 2) numpy bool, bit-identical to the
    reference's ``hap_gts.data``.

    devices  optional list of CUDA devices (or "all"): sample chunks are dealt round-robin to
             them -- samples shard with no communication, and every GPU brings its own PCIe
             link, which is what bounds this call.  Default: the one current (or `device`) GPU.
    result   how the result leaves the device.  "bits": one BIT per hap-call crosses PCIe
             (ht_transform_bits -> ht_bits_to_rowbits) and the host copy engine widens it to bool bytes
             straight into the result array, page-locked or not (1/8 of the D2H bytes; the host's store
             bandwidth becomes the bound); "u8": the bytes themselves (DMA into a page-locked array, through
             the engine's slots into a pageable one); "auto": "bits" for results of RESULT_BITS_MIN bytes or
             more (HAPTOOLS_B200_RESULT in the environment overrides).
    on_queued  optional callable, invoked once from this thread at the point after which it only waits (GIL
             released) for the device and the host copy engine: the class API starts its O(H) Python work there.
    """
    require_cuda()
    if gt.ndim != 3 or gt.shape[2] < 2 or gt.dtype not in (np.uint8, np.bool_):
        raise ValueError("genotypes must have shape (samples, variants, >=2) and dtype uint8 or bool")
    n, p = gt.shape[0], gt.shape[1]
    if p != plan.n_cols:
        raise ValueError(f"plan was made for {plan.n_cols} variant columns, genotypes have {p}")
    if plan.has_ancestry:
        if ancestry is None or ancestry.shape[:2] != gt.shape[:2] or ancestry.shape[2] < 2 or ancestry.dtype != np.uint8:
            raise ValueError("ancestry must be a uint8 array shaped like the genotypes")
    H = plan.n_haps
    if devices is None and device is None and os.environ.get("HAPTOOLS_B200_DEVICES"):
        # e.g. HAPTOOLS_B200_DEVICES=all or =0,1,2,3: lets Haplotypes.transform (which has no device
        # argument, like the reference's) use every GPU of the box
        env = os.environ["HAPTOOLS_B200_DEVICES"].strip()
        devices = "all" if env.lower() == "all" else [int(x) for x in env.split(",") if x.strip()]
    if devices == "all":
        devs = [torch.device("cuda", i) for i in range(torch.cuda.device_count())]
    elif devices:
        devs = [torch.device(d) if not isinstance(d, int) else torch.device("cuda", d) for d in devices]
    else:
        devs = [torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)]
    devs = [torch.device("cuda", d.index if d.index is not None else torch.cuda.current_device()) for d in devs]
    if out is None:
        nbytes = n * H * 2
        # A plain numpy result by default, like the reference's np.empty: page-locking a fresh buffer costs far
        # more than it saves in ONE call (3.3 GB: 1.66 s to allocate + 63 ms, against 314 ms with the staged
        # download into pageable memory, profiles/exp_first_call.py).  pin_output=True, or a page-locked `out`
        # that is reused (bench.py), gives the full PCIe rate.
        out = pinned_empty((n, H, 2)) if pin_output else _result_array((n, H, 2))
    elif out.shape != (n, H, 2) or out.dtype.itemsize != 1 or not out.flags.c_contiguous:
        raise ValueError("out must be a C-contiguous 1-byte array of shape (n, H, 2)")
    view = out.view(np.bool_)
    if n == 0 or H == 0:
        return view
    chunk = chunk_samples or default_chunk(n, p, H)
    if len(devs) > 1 and not chunk_samples:  # at least two chunks per device keep its three streams busy
        chunk = min(chunk, max(TILE, -(-n // (2 * len(devs))) // TILE * TILE))
    chunk = min(n, max(TILE, chunk // TILE * TILE)) if chunk < n else n
    # Row-gather uploads (variant-major array, few referenced variants) of a small call go in TWO even chunks instead
    # of the ramp: cutting the samples cuts every gathered row into pieces (5 KB rows of the 1000G-scale call ->
    # 0.5-2 KB in a ramp of five), and with 0.13 ms of kernels there is little to overlap (13.3-13.9 ms against
    # 14.8-15.2 ms; one chunk: 13.9-14.2 ms; profiles/r2/e2e_1000g_chunks.json)
    even = (not chunk_samples and len(devs) == 1 and n > 512 and _gatherable(gt, plan)
            and (ancestry is None or _gatherable(ancestry, plan))
            and upload_bytes(gt, plan, ancestry) <= (1 << 30) and out.nbytes <= (256 << 20))
    if even:
        chunk = _half_chunk(n)
    bounds = _ramped_bounds(n, chunk, len(devs)) if not (chunk_samples or even) else [(s, min(n, s + chunk)) for s in range(0, n, chunk)]
    devs = devs[: len(bounds)]
    if result == "auto":
        result = os.environ.get("HAPTOOLS_B200_RESULT", "").strip().lower() or "auto"
    if result == "auto":
        result = "bits" if out.nbytes >= RESULT_BITS_MIN and 2 * H <= (8 << 20) * 8 else "u8"
    if result not in ("bits", "u8"):
        raise ValueError("result must be 'auto', 'bits' or 'u8'")
    pipes = [_HostPipe(d, plan, dplan, gt, ancestry, bounds[i :: len(devs)], chunk, pack_mode, out, H, result)
             for i, d in enumerate(devs)]
    # on_queued: called once, when this thread has no more than waiting left to do -- after the loop when every pipe
    # holds a buffer set per chunk (nothing in the loop waits then), else before it (the loop is most of the call)
    for pp in pipes:
        pp.queue_all_uploads()
    early = on_queued is not None and not all(pp.queues_everything and not pp.src_pageable for pp in pipes)
    if early:
        on_queued()
    for k in range(len(bounds)):
        pipes[k % len(pipes)].step()
    if on_queued is not None and not early:
        on_queued()
    for pipe in pipes:
        pipe.finish()
    return view


# --------------------------------------------------------------------------------------
# QC: one device pass instead of the reference's host passes
# --------------------------------------------------------------------------------------
class QCResult:
    """What Genotypes.check_missing / check_biallelic / check_phase / check_maf need to know about a genotype
    matrix (haptools/data/genotypes.py:444-625), from ONE pass on the device (ht_qc_u8).  `first_*` are
    `(sample index, variant index)` of the pair np.nonzero would list first, or None."""

    __slots__ = ("n_samples", "n_variants", "first_missing", "first_multiallelic", "first_unphased", "n_missing",
                 "n_multiallelic", "samp_missing", "var_multiallelic", "alt_counts")

    def maf(self) -> np.ndarray:
        """haptools/data/genotypes.py:599-602."""
        ref_af = self.alt_counts / (2 * self.n_samples)
        return np.array([ref_af, 1 - ref_af]).min(axis=0)


_QC_NONE = 0xFFFFFFFFFFFFFFFF


def _qc_tensors(n: int, p: int, dev):
    summary = torch.zeros(5, dtype=torch.int64, device=dev)
    summary[:3] = -1  # first_* = ~0
    return (summary, torch.zeros(max(n, 1), dtype=torch.uint8, device=dev), torch.zeros(max(p, 1), dtype=torch.uint8, device=dev),
            torch.zeros(max(p, 1), dtype=torch.int64, device=dev))


def _qc_result(n: int, p: int, tensors) -> QCResult:
    summary, samp, varm, alt = tensors
    vals = summary.cpu().numpy().view(np.uint64)
    r = QCResult()
    r.n_samples, r.n_variants = n, p
    pair = lambda k: None if int(k) == _QC_NONE else (int(k) >> 32, int(k) & 0xFFFFFFFF)  # noqa: E731
    r.first_missing, r.first_multiallelic, r.first_unphased = pair(vals[0]), pair(vals[1]), pair(vals[2])
    r.n_missing, r.n_multiallelic = int(vals[3]), int(vals[4])
    r.samp_missing = samp[:n].cpu().numpy().astype(np.bool_)
    r.var_multiallelic = varm[:p].cpu().numpy().astype(np.bool_)
    r.alt_counts = alt[:p].cpu().numpy()
    return r


def qc_device(gt: torch.Tensor, missing_min: int = 254) -> QCResult:
    """QC of a genotype tensor resident in HBM: (n, p, 2|3) uint8/bool, any strides."""
    require_cuda()
    _check_gt_tensor(gt, "gt")
    n, p, depth = int(gt.shape[0]), int(gt.shape[1]), min(int(gt.shape[2]), 3)
    with torch.cuda.device(gt.device):
        tensors = _qc_tensors(n, p, gt.device)
        check(_cuda.lib().ht_qc_u8(gt.data_ptr(), *map(int, gt.stride()), depth, n, p, 0, missing_min, tensors[0].data_ptr(),
                                   tensors[1].data_ptr(), tensors[2].data_ptr(), tensors[3].data_ptr(), _stream_ptr()), "ht_qc_u8")
        return _qc_result(n, p, tensors)


def qc_host(data: np.ndarray, missing_min: int = 254, device=None, chunk_bytes: int = 256 << 20) -> QCResult:
    """
    QC of a host genotype matrix `(n, p, 2|3)` uint8/bool (any of the reference's layouts): sample chunks are
    uploaded (DMA from page-locked memory, host copy engine from pageable memory) while the previous chunk is
    checked by ht_qc_u8; the host never looks at a genotype byte.  Replaces the numpy passes of
    Genotypes.check_missing / check_biallelic / check_phase / check_maf (haptools/data/genotypes.py:444-625).
    """
    require_cuda()
    if data.ndim != 3 or data.shape[2] < 2 or data.dtype not in (np.uint8, np.bool_):
        raise ValueError("genotypes must have shape (samples, variants, >=2) and dtype uint8 or bool")
    n, p, depth = data.shape[0], data.shape[1], min(data.shape[2], 3)
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    lib = _cuda.lib()
    with torch.cuda.device(dev):
        tensors = _qc_tensors(n, p, dev)
        if n == 0 or p == 0:
            return _qc_result(n, p, tensors)
        rows = max(1, min(n, chunk_bytes // max(1, depth * p)))
        if rows < n:
            rows = max(TILE, rows // TILE * TILE)
        bounds = [(s, min(n, s + rows)) for s in range(0, n, rows)]
        nbuf = min(2, len(bounds))
        nbytes = max(_region(data, *b, depth=depth).dev_bytes() for b in bounds[:1] + bounds[-1:])
        bufs = [torch.empty(nbytes, dtype=torch.uint8, device=dev) for _ in range(nbuf)]
        s_in, s_cmp, cur = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.current_stream(dev)
        s_in.wait_stream(cur)
        s_cmp.wait_stream(cur)
        ev_cmp, keep = [None] * nbuf, []

        def upload(k):
            b = k % nbuf
            if ev_cmp[b] is not None:
                s_in.wait_event(ev_cmp[b])
            reg = _region(data, *bounds[k], depth=depth)
            keep.append(reg.keep)
            ticket = 0
            if _pageable(reg.ptr, reg.width * reg.rows) and reg.width <= (8 << 20):
                ticket = _submit(lib.ht_host_submit_upload(bufs[b].data_ptr(), reg.dev_pitch(), reg.ptr, reg.pitch, reg.width,
                                                           reg.rows, None, s_in.cuda_stream), "ht_host_submit_upload")
            else:
                _copy2d(bufs[b].data_ptr(), reg.dev_pitch(), reg.ptr, reg.pitch, reg.width, reg.rows, _cuda.HT_COPY_H2D, s_in)
            return reg, ticket, s_in.record_event()

        nxt = upload(0)
        for k, (s0, s1) in enumerate(bounds):
            reg, ticket, ev = nxt
            if k + 1 < len(bounds):
                nxt = upload(k + 1)
            _host_wait(ticket)
            s_cmp.wait_event(ev)
            ss, sv, sc = reg.dev_strides()
            check(lib.ht_qc_u8(bufs[k % nbuf].data_ptr(), ss, sv, sc, depth, s1 - s0, p, s0, missing_min, tensors[0].data_ptr(),
                               tensors[1].data_ptr() + s0, tensors[2].data_ptr(), tensors[3].data_ptr(), s_cmp.cuda_stream),
                  "ht_qc_u8")
            ev_cmp[k % nbuf] = s_cmp.record_event()
        s_cmp.synchronize()
        cur.wait_stream(s_cmp)
        return _qc_result(n, p, tensors)


def transform_host_bits(gt: np.ndarray, plan: TransformPlan, ancestry: np.ndarray = None, device=None,
                        chunk_samples: int = None, counts: bool = False):
    """
    Host genotypes -> the bit-packed result `(tiles, H, 32, 2)` int32 RESIDENT ON THE DEVICE (1/8 of
    the bytes of the uint8 form: 12.5 GB for 500 k samples x 100 k haplotypes), for the consumers
    that never need the `(n, H, 2)` byte matrix (writers.write_vcf_from_bits,
    stream.pgen_write_batches, dist.allgather_bits).  Sample chunks of whole tiles are uploaded on
    a copy stream while the previous one is packed and transformed; a chunk's result is a
    contiguous slice of the output because the layout is tile-major.  With `counts=True` also
    returns the per-haplotype allele counts (int64 tensor).
    """
    require_cuda()
    if gt.ndim != 3 or gt.shape[2] < 2 or gt.dtype not in (np.uint8, np.bool_):
        raise ValueError("genotypes must have shape (samples, variants, >=2) and dtype uint8 or bool")
    n, p = gt.shape[0], gt.shape[1]
    if p != plan.n_cols:
        raise ValueError(f"plan was made for {plan.n_cols} variant columns, genotypes have {p}")
    if plan.has_ancestry and (ancestry is None or ancestry.shape[:2] != gt.shape[:2] or ancestry.shape[2] < 2
                              or ancestry.dtype != np.uint8):
        raise ValueError("ancestry must be a uint8 array shaped like the genotypes")
    H = plan.n_haps
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    tiles = (n + TILE - 1) // TILE
    with torch.cuda.device(dev):
        dplan = device_plan(plan, dev)
        bits = torch.empty((tiles, H, 32, 2), dtype=torch.int32, device=dev)
        total = torch.zeros(H, dtype=torch.int64, device=dev) if counts else None
        if n == 0 or H == 0:
            return (bits, total) if counts else bits
        chunk = chunk_samples or max(TILE, (256 << 20) // max(2 * p, 1) // TILE * TILE)
        chunk = max(TILE, chunk // TILE * TILE)
        bounds = [(s, min(n, s + chunk)) for s in range(0, n, chunk)]
        nbuf = min(2, len(bounds))
        gt_bytes = max(_region(gt, *b).dev_bytes() for b in bounds[:1] + bounds[-1:])
        gt_buf = [torch.empty(gt_bytes, dtype=torch.uint8, device=dev) for _ in range(nbuf)]
        anc_buf = [None] * nbuf
        if plan.has_ancestry:
            anc_bytes = max(_region(ancestry, *b).dev_bytes() for b in bounds[:1] + bounds[-1:])
            anc_buf = [torch.empty(anc_bytes, dtype=torch.uint8, device=dev) for _ in range(nbuf)]
        planes = alloc_planes(min(chunk, n), plan.n_keys, dev)
        s_in, s_cmp, cur = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.current_stream(dev)
        s_in.wait_stream(cur)
        s_cmp.wait_stream(cur)
        ev_cmp = [None] * nbuf
        keep, dense = [], [None, None]
        for k, (s0, s1) in enumerate(bounds):
            b, rows = k % nbuf, s1 - s0
            if ev_cmp[b] is not None:
                s_in.wait_event(ev_cmp[b])
            rg = _region(gt, s0, s1)
            keep.append(rg.keep)
            _copy2d(gt_buf[b].data_ptr(), rg.dev_pitch(), rg.ptr, rg.pitch, rg.width, rg.rows, _cuda.HT_COPY_H2D, s_in)
            ra = None
            if plan.has_ancestry:
                ra = _region(ancestry, s0, s1)
                keep.append(ra.keep)
                _copy2d(anc_buf[b].data_ptr(), ra.dev_pitch(), ra.ptr, ra.pitch, ra.width, ra.rows, _cuda.HT_COPY_H2D, s_in)
            s_cmp.wait_event(s_in.record_event())
            ptrs = []
            for which, (reg, buf) in enumerate(((rg, gt_buf[b]), (ra, anc_buf[b]))):
                if reg is None:
                    ptrs.append((0, (0, 0, 0)))
                    continue
                ptr, strides = buf.data_ptr(), reg.dev_strides()
                if not _fast_layout(strides):  # phase column interleaved: dense re-layout on the device
                    d_str, d_bytes = reg.dense(rows, p)
                    if dense[which] is None or dense[which].numel() < d_bytes:
                        with torch.cuda.stream(s_cmp):
                            dense[which] = torch.empty(d_bytes, dtype=torch.uint8, device=dev)
                    check(_cuda.lib().ht_relayout_u8(ptr, *map(int, strides), rows, p, dense[which].data_ptr(),
                                                    d_str[0], d_str[1], s_cmp.cuda_stream), "ht_relayout_u8")
                    ptr, strides = dense[which].data_ptr(), d_str
                ptrs.append((ptr, strides))
            pack(dplan, ptrs[0][0], ptrs[0][1], rows, planes, ptrs[1][0], ptrs[1][1], None, _cuda.HT_PACK_AUTO, s_cmp)
            part = bits[s0 // TILE : (s1 + TILE - 1) // TILE]
            transform_bits(dplan, planes, rows, part, s_cmp)
            if counts:
                with torch.cuda.stream(s_cmp):
                    total += count_bits(part, rows, H, s_cmp)
            ev_cmp[b] = s_cmp.record_event()
        cur.wait_stream(s_cmp)
        s_cmp.synchronize()
        s_in.synchronize()
    return (bits, total) if counts else bits
