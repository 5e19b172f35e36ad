"""The SWAR bit tricks of csrc/ht_common.cuh, compiled for the host with g++ and checked
exhaustively / on random words (they are HT_HD = __host__ __device__)."""
import shutil
import subprocess
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parent.parent

SRC = r"""
#include <cstdio>
#include <cstdlib>
#include "ht_common.cuh"
int main() {
    using namespace ht;
    for (uint32_t n = 0; n < 16; ++n) {
        uint32_t w = spread4(n);
        for (int i = 0; i < 4; ++i) if (((w >> (8 * i)) & 0xff) != ((n >> i) & 1)) { printf("spread4 %u\n", n); return 1; }
    }
    uint64_t st = 1;
    for (int it = 0; it < 200000; ++it) {
        st = splitmix64(st);
        uint32_t x = (uint32_t)st, pat = ((uint32_t)(st >> 32) & 0xff) * 0x01010101u;
        if (it & 1) x = (x & 0xffff0000u) | (pat & 0xffffu);
        uint32_t f = eq_bytes_msb(x, pat);
        for (int i = 0; i < 4; ++i) {
            bool eq = ((x >> (8 * i)) & 0xff) == ((pat >> (8 * i)) & 0xff);
            if ((((f >> (8 * i)) & 0xff) == 0x80) != eq || ((f >> (8 * i)) & 0x7f)) { printf("eq_bytes %08x %08x\n", x, pat); return 1; }
        }
    }
    uint32_t a[32], b[32];
    for (int it = 0; it < 200; ++it) {
        for (int i = 0; i < 32; ++i) { st = splitmix64(st); a[i] = b[i] = (uint32_t)st; }
        transpose32(a);
        for (int i = 0; i < 32; ++i) for (int j = 0; j < 32; ++j)
            if (((a[i] >> j) & 1) != ((b[j] >> i) & 1)) { printf("transpose32\n"); return 1; }
    }
    if (splitmix64(0) != 0xe220a8397b1dcdafull) { printf("splitmix64\n"); return 1; }
    printf("ok\n");
    return 0;
}
"""


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
def test_bit_tricks_on_host(tmp_path):
    cuda_inc = Path("/usr/local/cuda/include")
    if not (cuda_inc / "cuda_runtime.h").exists():
        pytest.skip("CUDA headers not found")
    src = tmp_path / "bits.cpp"
    src.write_text(SRC)
    exe = tmp_path / "bits"
    subprocess.run(
        ["g++", "-O1", "-std=c++17", f"-I{REPO / 'include'}", f"-I{REPO / 'haptools_b200' / 'csrc'}", f"-I{cuda_inc}",
         str(src), "-o", str(exe)], check=True)
    res = subprocess.run([str(exe)], capture_output=True, text=True)
    assert res.returncode == 0 and res.stdout.strip() == "ok", res.stdout


def test_numpy_splitmix_matches_c_constant():
    import numpy as np

    from haptools_b200 import synth

    assert int(synth.splitmix64(np.uint64(0))) == 0xE220A8397B1DCDAF


def test_variant_major_multiply_gather_constants_exhaustive():
    """pack_vm_item's biallelic path (csrc/ht_pack.cu): four multiply-adds turn 16 genotype bytes of 0/1
    (8 samples x 2 interleaved strands) into the 8 strand-0 bits (byte 2 of the sum) and the 8 strand-1
    bits (byte 3).  The constants are read out of the kernel source and checked on all 2^16 inputs."""
    import re

    import numpy as np

    text = (REPO / "haptools_b200" / "csrc" / "ht_pack.cu").read_text()
    m = re.search(r"r\[j\] = it\.x\[j\]\.x \* (0x[0-9a-fA-F]+)u \+ it\.x\[j\]\.y \* (0x[0-9a-fA-F]+)u \+ "
                  r"it\.x\[j\]\.z \* (0x[0-9a-fA-F]+)u \+ it\.x\[j\]\.w \* (0x[0-9a-fA-F]+)u;", text)
    assert m, "the multiply-gather statement of pack_vm_item was not found"
    consts = [int(c, 16) for c in m.groups()]
    codes = np.arange(1 << 16, dtype=np.uint32)
    bits = ((codes[:, None] >> np.arange(16, dtype=np.uint32)) & 1).astype(np.uint8)  # byte i of the chunk
    words = np.ascontiguousarray(bits).view("<u4").astype(np.uint64)                  # (65536, 4)
    r = np.zeros(len(codes), dtype=np.uint64)
    for i, c in enumerate(consts):
        r = (r + words[:, i] * np.uint64(c)) & np.uint64(0xFFFFFFFF)
    strand0 = sum(bits[:, 2 * s].astype(np.uint64) << np.uint64(s) for s in range(8))
    strand1 = sum(bits[:, 2 * s + 1].astype(np.uint64) << np.uint64(s) for s in range(8))
    np.testing.assert_array_equal((r >> np.uint64(16)) & np.uint64(0xFF), strand0)
    np.testing.assert_array_equal((r >> np.uint64(24)) & np.uint64(0xFF), strand1)
