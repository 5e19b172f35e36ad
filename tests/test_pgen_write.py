"""PGEN-writer hand-off (N1b): oracle vs the reference-generated fixtures (CPU) and the
device expansion vs the oracle (GPU)."""
import numpy as np
import pytest

import _cases
from oracle import pgen_write_oracle as pwo

FIXTURES = sorted(_cases.GOLDEN_DIR.glob("pgenwrite_*.npz"))


@pytest.mark.parametrize("path", FIXTURES, ids=lambda p: p.stem)
def test_oracle_matches_reference_statements(path):
    z = np.load(path)
    for k in range(3):
        a, b = (int(x) for x in z[f"range{k}"])
        sub, cts = pwo.pgen_batch(z["data"], a, b)
        assert sub.dtype == np.int32 and sub.flags.c_contiguous
        np.testing.assert_array_equal(sub, z[f"subset{k}"])
        np.testing.assert_array_equal(cts, z[f"cts{k}"])
        assert cts.dtype == z[f"cts{k}"].dtype


@pytest.mark.gpu
@pytest.mark.parametrize("n,H", [(1, 1), (37, 9), (1023, 40), (1024, 33), (2600, 130), (4097, 17)])
def test_hapmajor_expansion_vs_oracle(n, H):
    import torch

    from haptools_b200 import dist as hdist
    from haptools_b200 import stream

    rng = np.random.default_rng(n + H)
    result = rng.random((n, H, 2)) < 0.35
    bits = torch.from_numpy(hdist.pack_result_bits(result).view(np.int32)).cuda()
    want, want_cts = pwo.pgen_batch(result, 0, H)
    for dtype in (torch.int32, torch.uint8):
        got = stream.hapmajor_from_bits(bits, n, 0, H, dtype)
        np.testing.assert_array_equal(got.cpu().numpy().astype(np.int32), want)
        a, b = H // 3, max(H // 3 + 1, (2 * H) // 3)
        part = stream.hapmajor_from_bits(bits, n, a, b, dtype)
        np.testing.assert_array_equal(part.cpu().numpy().astype(np.int32), want[a:b])
    for chunk in (1, 7, H):
        seen = 0
        for a, b, sub, cts in stream.pgen_write_batches(bits, n, chunk_haps=chunk):
            assert a == seen and sub.dtype == np.int32 and sub.flags.c_contiguous and sub.shape == (b - a, 2 * n)
            np.testing.assert_array_equal(sub, want[a:b])
            np.testing.assert_array_equal(cts, want_cts[a:b])
            seen = b
        assert seen == H
    # nothing outside the rows / beyond 2n
    pitch = (2 * n + 3) // 4 * 4 + 8
    buf = torch.full((H + 1, pitch), 77, dtype=torch.int32, device="cuda")
    stream.hapmajor_from_bits(bits, n, 0, H, torch.int32, out=buf)
    assert bool((buf[:H, 2 * n:] == 77).all()) and bool((buf[H] == 77).all())


@pytest.mark.gpu
def test_transform_to_pgen_batches_end_to_end():
    """genotypes -> planes -> bit-packed result -> writer batches == reference statements on the
    oracle's (n, H, 2) matrix."""
    import torch

    from haptools_b200 import engine, stream, synth
    from oracle import transform_oracle as orc

    n, p, H = 3000, 200, 150
    plan = synth.synth_plan(p, H, seed=3, max_alleles=5)
    gt = synth.synth_gt(n, p, seed=5, mode=1)
    want_mat = orc.transform_from_csr(gt, plan.key_col().astype(np.int64), plan.key_allele,
                                      plan.hap_off.astype(np.int64), plan.hap_key.astype(np.int64))
    dplan = engine.DevicePlan(plan)
    bits = engine.transform_device(torch.from_numpy(gt).cuda(), dplan, fmt="bits")
    rows = [sub.copy() for _, _, sub, _ in stream.pgen_write_batches(bits, n, chunk_haps=64)]
    np.testing.assert_array_equal(np.concatenate(rows), pwo.pgen_batch(want_mat, 0, H)[0])


@pytest.mark.gpu
@pytest.mark.parametrize("n", [64, 1026])  # 2n % 4 == 0: the yield is a VIEW of the page-locked buffer
def test_batch_stays_valid_while_the_next_one_is_fetched(n):
    """The documented contract of pgen_write_batches: a consumer (e.g. a writer thread) may hold batch
    k while it asks for batch k + 1; only batch k + 2 reuses k's buffer."""
    import torch

    from haptools_b200 import dist as hdist
    from haptools_b200 import stream

    H = 45
    rng = np.random.default_rng(n)
    result = rng.random((n, H, 2)) < 0.5
    bits = torch.from_numpy(hdist.pack_result_bits(result).view(np.int32)).cuda()
    want, _ = pwo.pgen_batch(result, 0, H)
    gen = stream.pgen_write_batches(bits, n, chunk_haps=4)
    held = next(gen)
    for nxt in gen:
        torch.cuda.synchronize()  # whatever the generator launched for later batches has landed
        a, b, sub, _ = held
        np.testing.assert_array_equal(sub, want[a:b])  # batch k, inspected AFTER batch k + 1 was produced
        held = nxt
    a, b, sub, _ = held
    np.testing.assert_array_equal(sub, want[a:b])
