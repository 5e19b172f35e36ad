"""
Multi-GPU execution of the transform: one process per GPU, samples sharded, NO communication
on the hot path (out[s, :, :] depends only on sample s).  The index tables
(a few MB) are replicated on every rank.  The only collective is an optional all-gather of
the bit-packed result (1/8 of the uint8 bytes) over torch.distributed (NCCL on GPUs; the
host-side logic is covered with gloo on CPU).
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
