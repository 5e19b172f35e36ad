// Pieces shared by the pack kernels (ht_pack.cu, ht_pack_anc_tma.cu).  See ht_pack.cu for the plane layout.
#pragma once

#include "ht_common.cuh"

namespace ht {

constexpr int kPackWarps = 8;
constexpr int kPackThreads = kPackWarps * 32;

struct PackArgs {
    const uint8_t* gt;
    int64_t gs, gv, gc;  // byte strides: sample, variant, strand
    const uint8_t* anc;
    int64_t as, av, ac;
    int64_t n;  // samples
    const uint32_t* var_col;
    const uint32_t* var_key_off;
    const uint8_t* key_allele;
    const uint8_t* key_label;
    const int32_t* col_var;  // dense column -> variant map (sample-major kernels)
    const int32_t* col_key01;  // dense column -> {key of allele 0, key of allele 1} (-1 = none)
    const int32_t* col_key_range;  // dense column -> {first key, key count} (ancestry fast path)
    const uint32_t* var_sel;   // optional indirection: generic kernel packs variants var_sel[i]
    int64_t n_cols;
    int64_t n_vars;
    int64_t n_keys;     // record pitch of the plane workspace (total keys)
    int64_t key_base;   // added to var_key_off-relative key ids (chunked int32 packing)
    uint32_t* planes;
    uint32_t* flags;
};

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    return __byte_perm(a, b, sel);
}

// flags at bit 7 of each byte -> the 4 flags as a nibble (byte 0 -> bit 0)
__device__ __forceinline__ uint32_t nib(uint32_t f) { return (f * 0x00204081u) >> 28; }

// 0x80 in every byte that is non-zero
__device__ __forceinline__ uint32_t nz_bytes_msb(uint32_t x) {
    return (((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x) & 0x80808080u;
}

// QC side channel: HT_FLAG_MISSING if some byte >= 254, HT_FLAG_NON_BIALLELIC if some byte in
// 2..253 (Genotypes.check_missing / check_biallelic, haptools/data/genotypes.py:444-528)
__device__ __forceinline__ uint32_t qc_flags4(uint32_t x) {
    const uint32_t miss = eq_bytes_msb(x | 0x01010101u, 0xffffffffu);
    const uint32_t hi = nz_bytes_msb(x & 0xfefefefeu) & ~miss;
    return (miss ? HT_FLAG_MISSING : 0u) | (hi ? HT_FLAG_NON_BIALLELIC : 0u);
}

__device__ __forceinline__ void publish_flags(uint32_t* flags, uint32_t f) {
    if (!flags) return;
    // lanes may have exited or diverged: reduce over whoever is here
    const unsigned m = __activemask();
    f = __reduce_or_sync(m, f);
    if ((threadIdx.x & 31) == __ffs(m) - 1 && f) {
        // one atomic per warp at most, and none once the bits are already set
        if (f & ~*reinterpret_cast<volatile uint32_t*>(flags)) atomicOr(flags, f);
    }
}

// mask of the samples of word `word` of `tile` that exist
__device__ __forceinline__ uint32_t valid_mask(int64_t n, int64_t tile, int word) {
    const int64_t rem = n - tile * kTileSamples - 32 * (int64_t)word;
    return rem >= 32 ? 0xffffffffu : (rem <= 0 ? 0u : ((1u << (int)rem) - 1u));
}

__device__ __forceinline__ uint2* record(const PackArgs& a, int64_t tile, uint32_t key) {
    return reinterpret_cast<uint2*>(a.planes) +
           ((size_t)tile * (size_t)a.n_keys + (size_t)a.key_base + key) * kTileWords;
}

// bytes b of G[0..3] -> word b (G[q] holds rows 8q..8q+7 of the 4 byte columns)
__device__ __forceinline__ void byte_transpose4(const uint32_t (&G)[4], uint32_t (&w)[4]) {
    const uint32_t t0 = prmt(G[0], G[1], 0x5140), t1 = prmt(G[2], G[3], 0x5140);
    const uint32_t t2 = prmt(G[0], G[1], 0x7362), t3 = prmt(G[2], G[3], 0x7362);
    w[0] = prmt(t0, t1, 0x5410);
    w[1] = prmt(t0, t1, 0x7632);
    w[2] = prmt(t2, t3, 0x5410);
    w[3] = prmt(t2, t3, 0x7632);
}

// ht_pack_anc_tma.cu: the TMA-fed ancestry pack (sample-major, keys with allele <= 1 and label < 8).
// HT_ERR_UNSUPPORTED (with the reason as the error string) when the driver cannot encode tensor maps here.
bool pack_anc_tma_available();
int launch_pack_sm2_anc_tma(const PackArgs& a, cudaStream_t st);

}  // namespace ht
