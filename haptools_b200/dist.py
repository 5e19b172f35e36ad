"""
Multi-GPU execution of the transform: one process per GPU, samples sharded, NO communication
on the hot path (out[s, :, :] depends only on sample s).  The index tables
(a few MB) are replicated on every rank.  The only collective is an optional all-gather of
the bit-packed result (1/8 of the uint8 bytes): `allgather_bits` over torch.distributed (NCCL on GPUs; the
host-side logic is covered with gloo on CPU), or `PeerGather`, where the transform kernel itself stores every
record into all ranks' gathered buffers over NVLink peer memory (compute and transfer in one kernel).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from ._cuda import TILE_SAMPLES as TILE


def shard_bounds(n_samples: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous sample range of `rank`: equal shards of whole 1024-sample tiles (so the
    packed results of the ranks concatenate into the packed result of the cohort)."""
    tiles = (n_samples + TILE - 1) // TILE
    per = (tiles + world - 1) // world
    lo = min(n_samples, rank * per * TILE)
    hi = min(n_samples, (rank + 1) * per * TILE)
    return lo, hi


def shard_tiles(n_samples: int, world: int) -> int:
    """Tiles per rank (the padded, equal shard size of the all-gather)."""
    tiles = (n_samples + TILE - 1) // TILE
    return (tiles + world - 1) // world


def allgather_bits(local_bits: torch.Tensor, n_samples: int, group=None) -> torch.Tensor:
    """
    local_bits: this rank's packed result (tiles_local, H, 32, 2) int32 for its shard_bounds
    range.  Returns the cohort's packed result (tiles_total, H, 32, 2) on every rank.
    Shards are padded with zero tiles to equal size; the gather is a plain concatenation
    because the layout is tile-major.
    """
    world = dist.get_world_size(group)
    per = shard_tiles(n_samples, world)
    H = local_bits.shape[1]
    send = local_bits
    if local_bits.shape[0] != per:
        send = torch.zeros((per, H, 32, 2), dtype=local_bits.dtype, device=local_bits.device)
        send[: local_bits.shape[0]] = local_bits
    out = torch.empty((world * per, H, 32, 2), dtype=local_bits.dtype, device=local_bits.device)
    dist.all_gather_into_tensor(out, send.contiguous(), group=group)
    tiles = (n_samples + TILE - 1) // TILE
    return out[:tiles]


class _DevMem:
    """A raw device address as something torch.as_tensor understands (__cuda_array_interface__)."""

    def __init__(self, ptr: int, shape, typestr: str = "<i4"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 3}


class PeerGather:
    """
    The gathered bit-packed result (tiles_total, H, 32, 2) of a sample-sharded cohort on EVERY rank, filled by
    `ht_transform_bits_gather`: each rank's transform kernel stores the records of its own shard into its own buffer
    and, through NVLink / NVSwitch peer memory, into every other rank's -- the all-gather is the kernel's stores,
    overlapped with the AND arithmetic record by record, instead of a collective that starts when the kernel ends.

    One process per GPU of one node (torch.distributed initialised, backend NCCL).  Buffers are allocated by the
    library (cudaMalloc: CUDA IPC handles need that), exchanged as 64-byte handles through the process group and
    mapped with lazy peer access.  `run()` is asynchronous on the current stream; the buffer is complete on every
    rank once the stream has passed the one-element all-reduce `run()` ends with (stream-ordered on every rank: it
    completes only when all ranks have reached it, i.e. when all their kernels -- and their remote stores -- are
    done).  world_size 1 works without a process group.
    """

    def __init__(self, n_samples: int, n_haps: int, device=None, group=None):
        import ctypes

        from . import _cuda

        self.lib, self._check = _cuda.lib(), _cuda.check
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if self.world > 16:
            raise ValueError("PeerGather serves the GPUs of one node (at most 16 ranks)")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.n_samples, self.n_haps = int(n_samples), int(n_haps)
        self.per = shard_tiles(n_samples, self.world)  # tiles per rank (equal shards)
        self.tiles = self.per * self.world
        nbytes = max(self.tiles * self.n_haps * 256, 256)
        own, handle = ctypes.c_void_p(), ctypes.create_string_buffer(64)
        self._own, self.ptrs, self._opened = None, [], []
        with torch.cuda.device(self.device):
            # Every step that can fail on ONE rank is followed by an exchange, so that all ranks raise together
            # (a rank that left alone would leave the others waiting in the next collective).
            err = None
            rc = self.lib.ht_peer_alloc(nbytes, ctypes.byref(own), handle)
            if rc == 0:
                self._own = own.value
            else:
                err = f"ht_peer_alloc: {self.lib.ht_last_error().decode('utf-8', 'replace')}"
            handles = [handle.raw if err is None else None]
            if self.world > 1:
                handles = [None] * self.world
                dist.all_gather_object(handles, handle.raw if err is None else None, group=group)
            if err is None and any(h is None for h in handles):
                err = "another rank could not allocate its buffer"
            if err is None:
                for r, h in enumerate(handles):
                    if r == self.rank:
                        self.ptrs.append(self._own)
                        continue
                    peer = ctypes.c_void_p()
                    rc = self.lib.ht_peer_open(ctypes.create_string_buffer(h, 64), ctypes.byref(peer))
                    if rc != 0:
                        err = f"ht_peer_open(rank {r}): {self.lib.ht_last_error().decode('utf-8', 'replace')}"
                        break
                    self.ptrs.append(peer.value)
                    self._opened.append(peer.value)
            if self.world > 1:
                errs = [None] * self.world
                dist.all_gather_object(errs, err, group=group)
                err = next((f"rank {r}: {e}" for r, e in enumerate(errs) if e is not None), None)
            if err is not None:
                for pp in self._opened:
                    self.lib.ht_peer_close(pp)
                if self._own is not None:
                    self.lib.ht_peer_free(self._own)
                self._own, self._opened = None, []
                from ._cuda import HaptoolsCudaError

                raise HaptoolsCudaError(f"PeerGather: {err}")
        self._dsts = (ctypes.c_void_p * self.world)(*self.ptrs)
        self._flag = torch.zeros(1, dtype=torch.int32, device=self.device)
        total_tiles = (self.n_samples + TILE - 1) // TILE
        self.bits = torch.as_tensor(_DevMem(self._own, (self.tiles, self.n_haps, 32, 2)), device=self.device)[:total_tiles]

    @staticmethod
    def key_order(dplan) -> torch.Tensor:
        """Haplotypes by their first key (stable; haplotypes without alleles first): the order in which the kernel
        computes them.  Cached on the DevicePlan."""
        order = getattr(dplan, "_key_order", None)
        if order is None:
            plan = dplan.plan
            off = plan.hap_off.astype(np.int64)
            first = np.full(plan.n_haps, -1, dtype=np.int64)
            has = off[1:] > off[:-1]
            first[has] = plan.hap_key[off[:-1][has]]
            order = torch.from_numpy(np.argsort(first, kind="stable").astype(np.int32)).to(dplan.hap_off.device)
            dplan._key_order = order
        return order

    def run(self, dplan, planes: torch.Tensor, n_local: int, stream=None, ordered: bool = False) -> torch.Tensor:
        """This rank's shard (n_local samples, `shard_bounds`) -> its records in every rank's buffer; returns the
        gathered tensor of THIS rank (complete when the stream has passed the closing all-reduce).  ordered: compute
        the haplotypes in key order (the same records; neighbouring warps share their plane reads, but every warp
        then stores its own 256-byte record instead of one 2 KB store per CTA: 11.4 against 10.5 ms at 2 GPUs, so
        it is off by default)."""
        if n_local > self.per * TILE:
            raise ValueError("shard larger than shard_tiles(n_samples, world) tiles")
        st = stream or torch.cuda.current_stream(self.device)
        order = self.key_order(dplan) if ordered and self.n_haps > 1 else None
        self._check(self.lib.ht_transform_bits_gather(
            planes.data_ptr() if planes is not None and planes.numel() else None, n_local, dplan.plan.n_keys,
            dplan.hap_off.data_ptr(), dplan.hap_key.data_ptr() if dplan.hap_key.numel() else None,
            order.data_ptr() if order is not None else None, self.n_haps,
            self._dsts, self.world, self.rank * self.per, st.cuda_stream), "ht_transform_bits_gather")
        if self.world > 1:
            with torch.cuda.stream(st):
                dist.all_reduce(self._flag, group=self.group)
        return self.bits

    def close(self) -> None:
        """Unmap the peers' buffers and free one's own (all ranks, after a barrier of the caller's)."""
        if getattr(self, "_own", None) is None:
            return
        torch.cuda.synchronize(self.device)
        with torch.cuda.device(self.device):
            for p in self._opened:
                self.lib.ht_peer_close(p)
            self._opened = []
            self.bits = None
            self.lib.ht_peer_free(self._own)
            self._own = None


def pack_result_bits(out_bool: np.ndarray) -> np.ndarray:
    """numpy restatement of the packed result layout (tiles, H, 32 words, 2 strands) uint32
    from an (n, H, 2) bool/0-1 array; bit b of word w of tile t = sample t*1024 + 32*w + b."""
    n, H, _ = out_bool.shape
    tiles = (n + TILE - 1) // TILE
    padded = np.zeros((tiles * TILE, H, 2), dtype=np.uint8)
    padded[:n] = out_bool
    bits = padded.reshape(tiles, 32, 32, H, 2).transpose(0, 3, 1, 4, 2)  # [t, h, w, c, bit]
    packed = np.packbits(bits, axis=-1, bitorder="little")  # (t, H, 32, 2, 4) bytes
    return np.ascontiguousarray(packed).view(np.uint32).reshape(tiles, H, 32, 2)


def unpack_result_bits(bits: np.ndarray, n_samples: int) -> np.ndarray:
    """Inverse of pack_result_bits: (tiles, H, 32, 2) uint32 -> (n, H, 2) bool."""
    t, H = bits.shape[:2]
    by = np.ascontiguousarray(bits).view(np.uint8).reshape(t, H, 32, 2, 4)
    u = np.unpackbits(by, axis=-1, bitorder="little").reshape(t, H, 32, 2, 32)
    return u.transpose(0, 2, 4, 1, 3).reshape(t * TILE, H, 2)[:n_samples].astype(np.bool_)
