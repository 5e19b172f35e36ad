"""Host half of the bit-packed result hand-off (csrc/ht_expand_host.cpp): bits -> numpy bool bytes.
Runs on the CPU (no CUDA call is made): every vector form this machine has, every alignment."""
import ctypes

import numpy as np
import pytest

from haptools_b200 import _cuda


def _expand(bits2d, n_out, out_pitch, out_offset=0):
    lib = _cuda.lib()
    rows = bits2d.shape[0]
    buf = np.full(rows * out_pitch + out_offset + 64, 0xAB, dtype=np.uint8)
    out = buf[out_offset:]
    rc = lib.ht_expand_rowbits_host(bits2d.ctypes.data, bits2d.strides[0], rows, n_out, out.ctypes.data, out_pitch)
    assert rc == 0
    return buf, out


@pytest.mark.parametrize("isa", [0, 1, 2])
def test_expand_matches_unpackbits(isa):
    lib = _cuda.lib()
    have = lib.ht_host_expand_isa(-2)
    if isa > have:
        pytest.skip(f"this CPU offers vector form {have} at most")
    assert lib.ht_host_expand_isa(isa) == isa
    try:
        rng = np.random.default_rng(isa)
        for n_out in (1, 7, 8, 9, 63, 64, 65, 255, 256, 257, 1000, 4096 + 13, 200_000 // 8 + 5):
            for off in (0, 1, 3, 16, 63):
                width = (n_out + 7) // 8
                rows = 3
                bits = rng.integers(0, 256, size=(rows, width + 5), dtype=np.uint8)[:, :width]  # pitched source
                pitch = n_out + 11
                buf, out = _expand(bits, n_out, pitch, off)
                want = np.unpackbits(np.ascontiguousarray(bits), axis=1, bitorder="little")[:, :n_out]
                got = np.lib.stride_tricks.as_strided(out, shape=(rows, n_out), strides=(pitch, 1))
                np.testing.assert_array_equal(got, want)
                # nothing outside the rows was written
                assert (buf[:off] == 0xAB).all()
                gaps = np.lib.stride_tricks.as_strided(out[n_out:], shape=(rows - 1, pitch - n_out), strides=(pitch, 1))
                assert (gaps == 0xAB).all()
                assert (out[(rows - 1) * pitch + n_out:] == 0xAB).all()
    finally:
        lib.ht_host_expand_isa(-2)


def test_expand_rejects_bad_arguments():
    lib = _cuda.lib()
    a = np.zeros(16, dtype=np.uint8)
    assert lib.ht_expand_rowbits_host(a.ctypes.data, 1, 2, 64, a.ctypes.data, 64) == _cuda.HT_ERR_BAD_ARG  # bits pitch < 8
    assert lib.ht_expand_rowbits_host(None, 8, 2, 64, a.ctypes.data, 64) == _cuda.HT_ERR_BAD_ARG
    assert lib.ht_expand_rowbits_host(a.ctypes.data, 8, 0, 64, a.ctypes.data, 64) == 0


def test_rowbits_pitch():
    lib = _cuda.lib()
    for H in (1, 15, 16, 17, 100_000, 12_345):
        pitch = lib.ht_rowbits_pitch(H)
        assert pitch % 16 == 0 and pitch >= (H + 15) // 16 * 4 and pitch * 8 >= 2 * H
