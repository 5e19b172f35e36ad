"""Host-side logic of the pipeline that needs no GPU: the chunk schedule and the upload accounting."""
import numpy as np
import pytest

from haptools_b200 import engine
from haptools_b200.plan import plan_from_refs


@pytest.mark.parametrize("n,chunk,devs", [(1, 1024, 1), (1000, 1024, 1), (16384, 2048, 1), (16385, 2048, 1),
                                          (100_000, 4096, 3), (5000, 1024, 8), (2048, 2048, 2)])
def test_ramped_bounds_cover_every_sample_once(n, chunk, devs):
    bounds = engine._ramped_bounds(n, chunk, devs)
    assert bounds[0][0] == 0 and bounds[-1][1] == n
    for (a0, a1), (b0, b1) in zip(bounds, bounds[1:]):
        assert a1 == b0  # contiguous, in order
    sizes = [b - a for a, b in bounds]
    assert all(0 < s <= chunk for s in sizes)
    if n > 4 * chunk:
        # the lead-in is short: the first chunk of every device is at most an eighth of a full one
        assert all(s <= max(256, chunk // 8) for s in sizes[:devs])
        assert max(sizes) == chunk or len(bounds) <= 3 * devs + 2  # (many devices: the ramp alone may cover n)


@pytest.mark.parametrize("n", [513, 1025, 2048, 2049, 2504, 4096, 5000, 100_001])
def test_half_chunk_makes_two_pieces(n):
    chunk = engine._half_chunk(n)
    bounds = [(s, min(n, s + chunk)) for s in range(0, n, chunk)]
    assert chunk % 256 == 0 and len(bounds) == 2 and bounds[0] == (0, chunk) and bounds[1][1] == n
    assert bounds[0][1] - bounds[0][0] >= bounds[1][1] - bounds[1][0] > 0


def _plan(p, cols):
    cols = np.asarray(cols, dtype=np.int64)
    return plan_from_refs(cols, np.ones(len(cols), dtype=np.int64), np.array([0, len(cols)]), p, None)


def test_upload_bytes_counts_what_is_copied():
    n, p = 100, 400
    sample_major = np.zeros((n, p, 2), dtype=np.uint8)
    variant_major = np.zeros((p, n, 2), dtype=np.uint8).transpose(1, 0, 2)
    with_phase = np.zeros((n, p, 3), dtype=np.uint8)[:, :, :2]
    sparse, dense = _plan(p, [3, 7, 250]), _plan(p, np.arange(0, p, 1))
    # sample-major: every byte the view spans, whatever the plan refers to
    assert engine.upload_bytes(sample_major, sparse) == n * p * 2
    assert engine.upload_bytes(with_phase, sparse) == n * (p * 3 - 1)  # a row spans up to its last genotype byte
    # variant-major: only the referenced variant rows when they are at most half of the variants
    assert engine.upload_bytes(variant_major, sparse) == 3 * n * 2
    assert engine.upload_bytes(variant_major, dense) == p * n * 2
    assert engine.upload_bytes(variant_major, sparse, ancestry=variant_major) == 2 * 3 * n * 2
