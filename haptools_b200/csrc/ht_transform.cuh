// Device helpers shared by the transform kernels (ht_transform.cu, ht_transform_local.cu).
#pragma once

#include "ht_common.cuh"

namespace ht {

constexpr int kHapBlock = 128;      // haplotypes per CTA
constexpr int kHapsPerWarp = 16;    // -> 32 result bits (16 haps x 2 strands) per sample
constexpr int kWarps = kHapBlock / kHapsPerWarp;  // 8
constexpr int kThreads = kWarps * 32;             // 256
constexpr int kStageKeys = 3072;    // key indices staged per CTA (12 KB, each haplotype's list
                                    // padded to a multiple of 4); longer blocks read their
                                    // lists straight from global memory
constexpr int kTilesPerCta = 1;     // consecutive tiles per CTA (key lists staged once)
constexpr int kUPitch = kTileSamples + 4;  // +4 words: the 8 warps' rows land in distinct banks

__device__ __forceinline__ void st_stream_v4(uint8_t* p, uint4 v) {
#ifdef HT_EXP_NO_OUT_STORE  // experiment only (profiles/): keep the math alive, drop the stores
    if ((v.x ^ v.y ^ v.z ^ v.w) != 0x9e3779b9u) return;
#endif
    // the 1-byte-per-call result is written once and never re-read by this kernel
#if defined(HT_EXP_STORE_MODE) && HT_EXP_STORE_MODE == 1  // experiments (profiles/): plain store
    *reinterpret_cast<uint4*>(p) = v;
#elif defined(HT_EXP_STORE_MODE) && HT_EXP_STORE_MODE == 2
    asm volatile("st.global.wt.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
#elif defined(HT_EXP_STORE_MODE) && HT_EXP_STORE_MODE == 3
    asm volatile("st.global.cg.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
#elif defined(HT_EXP_STORE_MODE) && HT_EXP_STORE_MODE == 4
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol) : "memory");
#else
    asm volatile("st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y),
                 "r"(v.z), "r"(v.w)
                 : "memory");
#endif
}

// transpose32 (ht_common.cuh) with the 16- and 8-bit stages done by PRMT (2 instructions per
// word pair instead of 5)
__device__ __forceinline__ void transpose32_dev(uint32_t (&a)[32]) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const uint32_t lo = a[k], hi = a[k + 16];
        a[k] = __byte_perm(lo, hi, 0x5410);
        a[k + 16] = __byte_perm(lo, hi, 0x7632);
    }
#pragma unroll
    for (int k = 0; k < 32; k = (k + 8 + 1) & ~8) {
        const uint32_t lo = a[k], hi = a[k + 8];
        a[k] = __byte_perm(lo, hi, 0x6240);
        a[k + 8] = __byte_perm(lo, hi, 0x7351);
    }
    uint32_t m = 0x0f0f0f0fu;
#pragma unroll
    for (int j = 4; j != 0; j >>= 1, m ^= (m << j)) {
#pragma unroll
        for (int k = 0; k < 32; k = (k + j + 1) & ~j) {
            const uint32_t t = ((a[k] >> j) ^ a[k + j]) & m;
            a[k] ^= (t << j);
            a[k + j] ^= t;
        }
    }
}

// half-tile form of the transform (ht_transform_half.cu); needs the planner's quads and 16-byte rows
bool transform_half_enabled();
// blk_key_lo / blk_key_cnt / stage_lines: optional key runs of the 128-haplotype blocks (staged form)
int launch_transform_u8_half(const uint32_t* planes, int64_t n_samples, int64_t n_keys,
                             const uint32_t* quad_off16, const uint4* quads, int64_t n_haps, uint8_t* out,
                             int64_t out_s_samp, unsigned long long* counts, cudaStream_t st,
                             const uint32_t* blk_key_lo = nullptr, const uint32_t* blk_key_cnt = nullptr,
                             int64_t stage_lines = 0);

__device__ __forceinline__ uint4 expand16(uint32_t bits) {
    uint4 o;
    o.x = spread4(bits & 15u);
    o.y = spread4((bits >> 4) & 15u);
    o.z = spread4((bits >> 8) & 15u);
    o.w = spread4((bits >> 12) & 15u);
    return o;
}

}  // namespace ht
