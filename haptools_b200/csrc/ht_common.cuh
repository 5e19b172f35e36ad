// Shared helpers for libhaptools_cuda (sm_100a).  See include/haptools_cuda.h for the ABI.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "haptools_cuda.h"

#if defined(__CUDACC__)
#define HT_HD __host__ __device__ __forceinline__
#else
#define HT_HD inline
#endif

namespace ht {

constexpr int kTileSamples = HT_TILE_SAMPLES;  // 1024
constexpr int kTileWords = HT_TILE_WORDS;      // 32
constexpr int kRecordWords = HT_RECORD_WORDS;  // 64 uint32 = 256 B per (tile, key)

// thread-local error string (ht_capi.cu)
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t err, const char* what);
int bad_arg(const char* what);

#define HT_CUDA_TRY(expr)                                        \
    do {                                                         \
        cudaError_t _e = (expr);                                 \
        if (_e != cudaSuccess) return ::ht::cuda_fail(_e, #expr); \
    } while (0)

// ---------------------------------------------------------------------------------------
// bit tricks (host+device so that tests/test_bits.cpp can check them with g++)
// ---------------------------------------------------------------------------------------

// 0x80 in every byte of x that equals the corresponding byte of `pat`, 0x00 elsewhere.
HT_HD uint32_t eq_bytes_msb(uint32_t x, uint32_t pat) {
    const uint32_t z = x ^ pat;  // zero byte <=> match
    // classic zero-byte detector; the per-byte add cannot carry across bytes
    const uint32_t t = ((z & 0x7f7f7f7fu) + 0x7f7f7f7fu) | z;
    return ~t & 0x80808080u;
}

// 4 low bits of nib -> 4 bytes of 0x00/0x01 (byte i = bit i).  Requires nib < 16.
HT_HD uint32_t spread4(uint32_t nib) {
    // nib | nib<<7 | nib<<14 | nib<<21 : the four copies do not overlap, so the multiply
    // is carry-free; bit 8*i then holds bit i of nib.
    return (nib * 0x00204081u) & 0x01010101u;
}

// In-place transpose of a 32x32 bit matrix held as 32 words, LSB-first convention:
// after the call, bit j of a[i] == bit i of (old) a[j].
HT_HD void transpose32(uint32_t (&a)[32]) {
    uint32_t m = 0x0000ffffu;
#pragma unroll
    for (int j = 16; j != 0; j >>= 1, m ^= (m << j)) {
#pragma unroll
        for (int k = 0; k < 32; k = (k + j + 1) & ~j) {
            // swap the high-j bits of a[k] with the low-j bits of a[k + j]
            const uint32_t t = ((a[k] >> j) ^ a[k + j]) & m;
            a[k] ^= (t << j);
            a[k + j] ^= t;
        }
    }
}

HT_HD uint64_t splitmix64(uint64_t x) {
    x += 0x9e3779b97f4a7c15ull;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}

}  // namespace ht
