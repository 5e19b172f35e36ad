// K2 -- the haplotype transform proper (sm_100a).
//
// Replaces the reference's per-haplotype loop
//     hap_gts.data[:, i] = np.all(equality_arr[:, idxs[i]], axis=1)
// (haptools/data/haplotypes.py:1411-1412; haptools/transform.py:144-148 with the ancestry
// comparison folded into the keys by the pack kernel).
//
// Work decomposition
//   CTA  = one (sample tile of 1024 samples) x (block of 128 haplotypes); 8 warps.
//   warp = 16 haplotypes of the block; lane = one 32-sample word of the tile.
//   1. AND: for each of its 16 haplotypes the warp walks the haplotype's key list (staged in
//      shared memory for the whole block) and ANDs the 256-byte (tile, key) records --
//      lane L reads the uint2 {strand0, strand1} of word L, so each key is one coalesced
//      256 B read that hits L2 (all CTAs in flight work on the same tile: blockIdx is
//      tile-major).  32 accumulators per lane: a[2*i + c] = haplotype i, strand c.
//   2. 32x32 bit transpose in registers: a[j] now holds, for sample 32*L + j, the 32 result
//      bits in output order (hap-major, strand-minor = the byte order of a row of `out`).
//   3. Exchange through shared memory (swizzled, conflict-free both ways) so that 16
//      consecutive lanes cover the 256 contiguous output bytes of one sample row; each lane
//      expands 16 bits to 16 bytes (carry-free multiply) and issues one 16-byte streaming
//      store.  Every store instruction writes 2 x 256 contiguous bytes.
//
// Roofline: HBM write of n*H*2 bytes; plane reads are L2 hits (A*256 B per tile).
#include <cstring>

#include "ht_transform.cuh"

namespace ht {

__device__ __forceinline__ uint2 ld_plane(const uint2* p) {
    // Plane records are re-read by every haplotype that uses the key (about 11 times per tile on the
    // UKB-scale plan) while 8x as many result bytes stream through L2: ask L2 to evict them last
    // (createpolicy is pure -- not volatile -- so the compiler keeps one copy per kernel).
#if defined(HT_EXP_LD_MODE) && HT_EXP_LD_MODE == 1  // experiments (profiles/exp_variants.py)
    uint2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
#elif defined(HT_EXP_LD_MODE) && HT_EXP_LD_MODE == 2
    uint2 v;
    asm volatile("ld.global.cg.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
#elif defined(HT_EXP_LD_MODE) && HT_EXP_LD_MODE == 3
    uint2 v;
    asm volatile("ld.global.nc.L1::evict_last.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
#elif defined(HT_EXP_LD_MODE) && HT_EXP_LD_MODE == 5  // plain read-only load, no L2 hint (the v5 default)
    return __ldg(p);
#else
    uint2 v;
    uint64_t pol;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("ld.global.nc.L2::cache_hint.v2.u32 {%0, %1}, [%2], %3;" : "=r"(v.x), "=r"(v.y) : "l"(p), "l"(pol));
    return v;
#endif
}

// AND of the (tile, key) records of one haplotype, for this lane's word; both strands.
// Unstaged variant (key list read from global memory): 4 keys per round trip, the last batch
// repeats the final key (AND is idempotent) instead of a scalar tail loop.  Loops are
// deliberately NOT unrolled: the whole kernel has to stay inside the instruction cache.
__device__ __forceinline__ uint2 and_keys_global(const uint2* __restrict__ tile_rec,  // + lane
                                                 const uint32_t* __restrict__ keys, uint32_t kb,
                                                 uint32_t ke) {
    uint32_t r0 = 0xffffffffu, r1 = 0xffffffffu;
    const uint32_t last = ke - 1;
#pragma unroll 1
    for (uint32_t j = kb; j < ke; j += 4) {
        const uint32_t k0 = __ldg(keys + j), k1 = __ldg(keys + min(j + 1, last));
        const uint32_t k2 = __ldg(keys + min(j + 2, last)), k3 = __ldg(keys + min(j + 3, last));
        const uint2 v0 = ld_plane(tile_rec + (size_t)k0 * kTileWords);
        const uint2 v1 = ld_plane(tile_rec + (size_t)k1 * kTileWords);
        const uint2 v2 = ld_plane(tile_rec + (size_t)k2 * kTileWords);
        const uint2 v3 = ld_plane(tile_rec + (size_t)k3 * kTileWords);
        r0 &= (v0.x & v1.x) & (v2.x & v3.x);
        r1 &= (v0.y & v1.y) & (v2.y & v3.y);
    }
    return make_uint2(r0, r1);
}

// Staged key quads hold BYTE offsets of the keys' records inside a tile (key * 256); the low
// bits of the FIRST entry of a quad carry flags.  (Needs n_keys < 2^24; else unstaged path.)
constexpr uint32_t kQuadLast = 1u;  // last quad of its haplotype: emit the result
constexpr uint32_t kQuadSkip = 2u;  // stands for a haplotype without alleles
constexpr uint32_t kOffMask = 0xffffff00u;
constexpr int64_t kMaxStagedKeys = 1 << 24;

#ifndef HT_BATCH_QUADS
#define HT_BATCH_QUADS 2
#endif
#ifndef HT_MIN_CTAS
#define HT_MIN_CTAS 4
#endif
constexpr int kBQ = HT_BATCH_QUADS;  // quads per batch: 4 * kBQ plane records in flight per batch

struct Batch {
    uint2 v[4 * kBQ];
    uint32_t f[kBQ];  // flag words of the quads
};

__device__ __forceinline__ uint2 ld_rec(const char* tile_bytes, uint32_t off) {
    return ld_plane(reinterpret_cast<const uint2*>(tile_bytes + off));
}

__device__ __forceinline__ void load_batch(const char* __restrict__ tile_bytes,
                                           const uint4* __restrict__ quads, uint32_t q,
                                           uint32_t q_last, Batch& b) {
    uint4 k[kBQ];
#pragma unroll
    for (int i = 0; i < kBQ; ++i) k[i] = quads[min(q + i, q_last)];
#pragma unroll
    for (int i = 0; i < kBQ; ++i) {
        b.f[i] = k[i].x;
        b.v[4 * i + 0] = ld_rec(tile_bytes, k[i].x & kOffMask);
        b.v[4 * i + 1] = ld_rec(tile_bytes, k[i].y);
        b.v[4 * i + 2] = ld_rec(tile_bytes, k[i].z);
        b.v[4 * i + 3] = ld_rec(tile_bytes, k[i].w);
    }
}

__device__ __forceinline__ void consume_quad(const uint2* v, uint32_t flags, uint32_t& r0,
                                             uint32_t& r1, uint2*& park) {
    if (!(flags & kQuadSkip)) {
        r0 &= (v[0].x & v[1].x) & (v[2].x & v[3].x);
        r1 &= (v[0].y & v[1].y) & (v[2].y & v[3].y);
    }
    if (flags & kQuadLast) {  // warp-uniform
        *park = make_uint2(r0, r1);
        park += 32;
        r0 = r1 = 0xffffffffu;
    }
}

// Staged AND phase of one warp.  The key lists of its 16 haplotypes are consecutive quads in
// shared memory (each list padded to a multiple of 4 by repeating its last key -- AND is
// idempotent; a haplotype without alleles gets one "skip" quad), so the warp walks them as
// ONE stream, two quads (8 records of 256 B) per step, always with the next step's 8 reads
// already in flight; a haplotype's result is emitted when its last quad is consumed.
__device__ __forceinline__ void and_phase_staged(const char* __restrict__ tile_rec,
                                                 const uint4* __restrict__ quads, uint32_t q,
                                                 uint32_t q_end, uint2* park) {
    if (q >= q_end) return;
    uint32_t r0 = 0xffffffffu, r1 = 0xffffffffu;
    const uint32_t q_last = q_end - 1;
    Batch A, B;
    load_batch(tile_rec, quads, q, q_last, A);
#pragma unroll 1
    while (true) {
        if (q + kBQ < q_end) load_batch(tile_rec, quads, q + kBQ, q_last, B);
#pragma unroll
        for (int i = 0; i < kBQ; ++i)
            if (i == 0 || q + i < q_end) consume_quad(A.v + 4 * i, A.f[i], r0, r1, park);
        q += kBQ;
        if (q >= q_end) break;
        if (q + kBQ < q_end) load_batch(tile_rec, quads, q + kBQ, q_last, A);
#pragma unroll
        for (int i = 0; i < kBQ; ++i)
            if (i == 0 || q + i < q_end) consume_quad(B.v + 4 * i, B.f[i], r0, r1, park);
        q += kBQ;
        if (q >= q_end) break;
    }
}

// Stage the CSR slice of one haplotype block (offsets, then every warp the lists of its own
// 16 haplotypes as padded, flagged quads of record byte offsets).  Called by all threads of the
// CTA; `my_row` (the warp's exchange row) is scratch.  Returns whether the lists were staged.
__device__ __forceinline__ bool stage_block(const uint32_t* __restrict__ hap_off,
                                            const uint32_t* __restrict__ hap_key, int64_t h0,
                                            int n_blk_haps, int64_t n_keys, uint32_t* s_off,
                                            uint32_t* s_quad, uint32_t* s_wsum, uint32_t* s_keys,
                                            uint32_t* my_row, int hl_begin, int hl_end) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // ---- stage the block's CSR slice: offsets, then each key list padded to quads ------
    for (int i = tid; i <= n_blk_haps; i += kThreads) s_off[i] = __ldg(hap_off + h0 + i);
    __syncthreads();
    if (tid < kHapBlock) {  // exclusive prefix sum of the quad counts of the block's haplotypes
        const uint32_t nq = tid < n_blk_haps ? max(1u, (s_off[tid + 1] - s_off[tid] + 3u) >> 2) : 0u;
        uint32_t inc = nq;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += t;
        }
        if (lane == 31) s_wsum[warp] = inc;
        s_quad[tid] = inc - nq;  // within-warp exclusive
    }
    __syncthreads();
    if (tid < kHapBlock) {
        uint32_t base = 0;
        for (int w = 0; w < warp; ++w) base += s_wsum[w];
        s_quad[tid] += base;
    }
    if (tid == 0) s_quad[kHapBlock] = s_wsum[0] + s_wsum[1] + s_wsum[2] + s_wsum[3];
    __syncthreads();
    // n_keys == 0 means every haplotype is empty and `planes` may be NULL: never load
    // every warp stages (and later consumes) the lists of its own 16 haplotypes: a contiguous
    // slice of hap_key.  Raw copy first (independent coalesced loads, one round trip), into
    // the warp's exchange row, which is free until the AND phase parks results in it.
    const uint32_t w_kb = hl_begin < hl_end ? s_off[hl_begin] : 0u;
    const uint32_t w_cnt = hl_begin < hl_end ? s_off[hl_end] - w_kb : 0u;
    const bool staged = s_quad[kHapBlock] * 4u <= (uint32_t)kStageKeys && n_keys > 0 &&
                        n_keys < kMaxStagedKeys && w_cnt <= (uint32_t)kTileSamples;
    if (staged) {
#pragma unroll 1
        for (uint32_t j = lane; j < w_cnt; j += 128) {
            const uint32_t* src = hap_key + w_kb + j;
            const uint32_t k0 = __ldg(src);
            const uint32_t k1 = j + 32 < w_cnt ? __ldg(src + 32) : 0u;
            const uint32_t k2 = j + 64 < w_cnt ? __ldg(src + 64) : 0u;
            const uint32_t k3 = j + 96 < w_cnt ? __ldg(src + 96) : 0u;
            my_row[j] = k0;
            if (j + 32 < w_cnt) my_row[j + 32] = k1;
            if (j + 64 < w_cnt) my_row[j + 64] = k2;
            if (j + 96 < w_cnt) my_row[j + 96] = k3;
        }
        __syncwarp();
        // pad each list to whole quads (repeat its last key) and flag the quads
#pragma unroll 1
        for (int hl = hl_begin; hl < hl_end; ++hl) {
            const uint32_t kb = s_off[hl] - w_kb, len = s_off[hl + 1] - s_off[hl];
            const uint32_t base = 4u * s_quad[hl], padded = max(4u, (len + 3u) & ~3u);
#pragma unroll 1
            for (uint32_t j = lane; j < padded; j += 32) {
                uint32_t off = len ? my_row[kb + min(j, len - 1)] << 8 : 0u;  // byte offset of the record
                if ((j & 3u) == 0u) off |= (j + 4 >= padded ? kQuadLast : 0u) | (len ? 0u : kQuadSkip);
                s_keys[base + j] = off;
            }
        }
    }
    __syncwarp();

    return staged;
}

// out_vec: 16 -> every full 8-haplotype group is stored with one 16-byte store (requires
//          out and out_s_samp to be multiples of 16); 2 -> 2-byte stores only (any layout
//          with even addresses); 1 -> byte stores.
template <int kOutVec>
__global__ void __launch_bounds__(kThreads, HT_MIN_CTAS)
transform_u8_kernel(const uint32_t* __restrict__ planes, int64_t n_samples, int64_t n_keys,
                    const uint32_t* __restrict__ hap_off, const uint32_t* __restrict__ hap_key,
                    const uint32_t* __restrict__ quad_off16, const uint4* __restrict__ quads,
                    int64_t n_haps, uint32_t n_hap_blocks, int64_t tile0,
                    uint8_t* __restrict__ out, int64_t out_s_samp,
                    unsigned long long* __restrict__ counts) {
    __shared__ __align__(16) uint32_t s_u[kWarps * kUPitch];  // 32.9 KB
    __shared__ __align__(16) uint32_t s_keys[kStageKeys];      // 12 KB, as uint4 quads
    __shared__ uint32_t s_off[kHapBlock + 1];                  // CSR offsets of the block
    __shared__ uint32_t s_quad[kHapBlock + 1];                 // first quad of each haplotype
    __shared__ uint32_t s_wsum[kHapBlock / 32];

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    // a CTA handles kTilesPerCta consecutive tiles of one haplotype block: the key lists are
    // staged once and reused
    const int64_t tile_first = tile0 + (int64_t)(blockIdx.x / n_hap_blocks) * kTilesPerCta;
    const int64_t n_tiles_all = (n_samples + kTileSamples - 1) / kTileSamples;
    const int64_t h0 = (int64_t)(blockIdx.x % n_hap_blocks) * kHapBlock;
    const int n_blk_haps = (int)min((int64_t)kHapBlock, n_haps - h0);

    const int hl_begin = warp * kHapsPerWarp;
    const int hl_end = min(hl_begin + kHapsPerWarp, n_blk_haps);
    uint32_t* my_row = s_u + warp * kUPitch;
    // n_keys == 0 means every haplotype is empty and `planes` may be NULL: never load
    bool staged = false;
    uint32_t q_begin = 0, q_end = 0;
    if (quads) {
        // the planner already laid the lists out as padded, flagged quads (plan.py quad_tables):
        // every warp copies the quads of its own 16 haplotypes, no CTA-wide pass
        const int64_t g0 = (h0 >> 4), n16 = (n_haps + kHapsPerWarp - 1) >> 4;
        const uint32_t blk_q0 = __ldg(quad_off16 + g0);
        const uint32_t blk_q1 = __ldg(quad_off16 + min(g0 + kWarps, n16));
        if ((blk_q1 - blk_q0) * 4u <= (uint32_t)kStageKeys) {
            staged = true;
            if (hl_begin < hl_end) {
                const uint32_t w_q0 = __ldg(quad_off16 + g0 + warp), w_q1 = __ldg(quad_off16 + g0 + warp + 1);
                q_begin = w_q0 - blk_q0;
                q_end = w_q1 - blk_q0;
                uint4* dst = reinterpret_cast<uint4*>(s_keys);
                for (uint32_t i = q_begin + lane; i < q_end; i += 32) dst[i] = __ldg(quads + blk_q0 + i);
            }
            __syncwarp();
        }
    }
    if (!staged) {  // block-uniform
        staged = stage_block(hap_off, hap_key, h0, n_blk_haps, n_keys, s_off, s_quad, s_wsum, s_keys,
                             my_row, hl_begin, hl_end);
        q_begin = s_quad[hl_begin];
        q_end = s_quad[hl_end];
    }

    uint2* park = reinterpret_cast<uint2*>(my_row) + lane;

#pragma unroll 1
    for (int ti = 0; ti < kTilesPerCta; ++ti) {
        const int64_t tile = tile_first + ti;
        if (tile >= n_tiles_all) break;  // uniform over the CTA
        if (ti) __syncthreads();         // everybody is done reading the exchange buffer

        // ---- 1. AND ------------------------------------------------------------------
        // results are parked in the warp's own slice of s_u (which is only reused for the
        // exchange after the warp has read everything back): keeps the haplotype loop rolled
        const char* tile_bytes = reinterpret_cast<const char*>(planes) +
                                 ((size_t)tile * (size_t)n_keys) * (kRecordWords * 4) + 8 * lane;
        if (staged) {
            if (hl_begin < hl_end)
                and_phase_staged(tile_bytes, reinterpret_cast<const uint4*>(s_keys), q_begin, q_end, park);
        } else {
            const uint2* tile_rec = reinterpret_cast<const uint2*>(tile_bytes);
#pragma unroll 1
            for (int hl = hl_begin; hl < hl_end; ++hl) {
                uint2 r = make_uint2(0xffffffffu, 0xffffffffu);
                if (n_keys > 0) r = and_keys_global(tile_rec, hap_key, s_off[hl], s_off[hl + 1]);
                park[(hl - hl_begin) * 32] = r;
            }
        }
        uint32_t a[32];  // a[2*i + c] = haplotype i of this warp, strand c
#pragma unroll
        for (int i = 0; i < kHapsPerWarp; ++i) {
            const uint2 r = park[i * 32];
            a[2 * i] = r.x;
            a[2 * i + 1] = r.y;
        }
        if (counts) {
            // allele counts for check_maf (haptools/data/genotypes.py:559-625), fused: popcount
            // of my word of every result, one warp reduction and one atomic per haplotype and tile
            const int64_t rem = n_samples - tile * kTileSamples - 32 * (int64_t)lane;
            const uint32_t vm = rem >= 32 ? 0xffffffffu : (rem <= 0 ? 0u : ((1u << (int)rem) - 1u));
#pragma unroll
            for (int i = 0; i < kHapsPerWarp; ++i) {
                const uint32_t c = __reduce_add_sync(0xffffffffu, __popc(a[2 * i] & vm) + __popc(a[2 * i + 1] & vm));
                if (lane == 0 && hl_begin + i < hl_end && c) atomicAdd(counts + h0 + hl_begin + i, (unsigned long long)c);
            }
        }
        __syncwarp();

        // ---- 2. bits over samples -> bits over (haplotype, strand) --------------------
        transpose32_dev(a);

        // ---- 3. exchange + expand + store ----------------------------------------------
        // sample 32*L + j of the tile is kept at slot 32*L + ((j + L) & 31): lanes write
        // distinct banks here, and the 16 words one warp reads below fall into 16 distinct banks.
        uint32_t* my_u = my_row + 32 * lane;
#pragma unroll
        for (int j = 0; j < 32; ++j) my_u[(j + lane) & 31] = a[j];
        __syncthreads();

        // (cheap to recompute per tile; keeping them live across the AND phase costs registers)
        const int half = threadIdx.x & 1;               // low / high 16 bits of the word = haps 0-7 / 8-15
        const int b = (threadIdx.x >> 1) & 7;           // which warp's 16 haplotypes
        const int c0 = (threadIdx.x >> 5) * 2 + ((threadIdx.x >> 4) & 1);  // sample row within each group of 16
        const int hl0 = b * kHapsPerWarp + half * 8;    // first haplotype (in block) of my 16 bytes
        const int n_my_haps = min(8, n_blk_haps - hl0);  // may be <= 0
        const uint32_t* u_col = s_u + b * kUPitch;
        const int shift = 16 * half;
        if (n_my_haps <= 0) continue;  // the barriers above are reached by everybody
        const int64_t s_base = tile * kTileSamples;
        const int rows = (int)min((int64_t)kTileSamples, n_samples - s_base);
        uint8_t* p = out + 2 * (h0 + hl0) + (s_base + c0) * out_s_samp;
        const int64_t p_step = 16 * out_s_samp;

        if (kOutVec == 16 && n_my_haps == 8 && rows == kTileSamples) {
            // full tile, full group: two rows (32*L + c0 and 32*L + 16 + c0) per iteration
#pragma unroll 4
            for (int L = 0; L < 32; ++L) {
                const int t = (c0 + L) & 31;
                const uint32_t w0 = u_col[32 * L + t], w1 = u_col[32 * L + (t ^ 16)];
                st_stream_v4(p, expand16((w0 >> shift) & 0xffffu));
                st_stream_v4(p + p_step, expand16((w1 >> shift) & 0xffffu));
                p += 2 * p_step;
            }
            continue;
        }
#pragma unroll 1
        for (int it = 0; it < kTileSamples / 16; ++it, p += p_step) {
            const int sl = it * 16 + c0;
            if (sl >= rows) break;  // rows only grow with `it`
            const int L = it >> 1, j = ((it & 1) << 4) + c0;
            const uint32_t bits = (u_col[32 * L + ((j + L) & 31)] >> shift) & 0xffffu;
            if (kOutVec == 16 && n_my_haps == 8) {
                st_stream_v4(p, expand16(bits));
            } else if (kOutVec >= 2) {
                for (int q = 0; q < n_my_haps; ++q) {
                    const uint32_t two = (bits >> (2 * q)) & 3u;
                    *reinterpret_cast<uint16_t*>(p + 2 * q) = (uint16_t)((two & 1u) | ((two & 2u) << 7));
                }
            } else {
                for (int q = 0; q < 2 * n_my_haps; ++q) p[q] = (uint8_t)((bits >> q) & 1u);
            }
        }
    }
}

// Bit-packed result: out_bits[tile][hap][word][strand]; warp per haplotype.
__global__ void __launch_bounds__(kThreads)
transform_bits_kernel(const uint32_t* __restrict__ planes, int64_t n_samples, int64_t n_keys,
                      const uint32_t* __restrict__ hap_off, const uint32_t* __restrict__ hap_key,
                      int64_t n_haps, uint32_t n_hap_blocks, int64_t tile0,
                      uint32_t* __restrict__ out_bits) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t tile = tile0 + blockIdx.x / n_hap_blocks;
    const int64_t h = (int64_t)(blockIdx.x % n_hap_blocks) * kWarps + warp;
    if (h >= n_haps) return;
    const uint2* tile_rec =
        reinterpret_cast<const uint2*>(planes) + ((size_t)tile * (size_t)n_keys) * kTileWords + lane;
    uint2 r = and_keys_global(tile_rec, hap_key, __ldg(hap_off + h), __ldg(hap_off + h + 1));
    // samples beyond n_samples read as 0, like the planes
    const int64_t rem = n_samples - tile * kTileSamples - 32 * (int64_t)lane;
    const uint32_t vm = rem >= 32 ? 0xffffffffu : (rem <= 0 ? 0u : ((1u << rem) - 1u));
    r.x &= vm;
    r.y &= vm;
    reinterpret_cast<uint2*>(out_bits)[((size_t)tile * (size_t)n_haps + (size_t)h) * kTileWords + lane] = r;
}

// Bit-packed result + all-gather: the record of (local tile, haplotype) goes into every rank's gathered buffer at
// tile dst_tile0 + local tile.  The 8 warps of a CTA compute 8 consecutive haplotypes of one tile, i.e. 2 KB that are
// contiguous in every destination: the records are staged in shared memory and ONE thread hands them to the TMA as
// a bulk store per destination (cp.async.bulk.global.shared::cta) -- full-width NVLink writes issued by the copy
// unit while the warps are already gone, instead of 8-byte stores per lane (390 GB/s against 770 for a DMA between
// two GPUs, profiles/r2/p2p.json).
struct PeerDsts {
    uint32_t* p[HT_MAX_PEERS];
};

// kOrdered: warp i of the grid computes haplotype hap_order[i] (the planner's key order: neighbouring warps share
// plane lines, which the L1 / L2 then serve -- 3 x fewer L2->SM reads than in an arbitrary caller order) and every
// warp hands its own 256-byte record to the TMA; records scatter at no cost, they are whole and aligned.
template <bool kOrdered>
__global__ void __launch_bounds__(kThreads)
transform_bits_gather_kernel(const uint32_t* __restrict__ planes, int64_t n_samples, int64_t n_keys,
                             const uint32_t* __restrict__ hap_off, const uint32_t* __restrict__ hap_key,
                             const uint32_t* __restrict__ hap_order, int64_t n_haps, uint32_t n_hap_blocks, int64_t tile0,
                             const PeerDsts dsts, int n_dsts, int64_t dst_tile0) {
    __shared__ __align__(128) uint2 s_rec[kWarps][kTileWords];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t tile = tile0 + blockIdx.x / n_hap_blocks;
    const int64_t h0 = (int64_t)(blockIdx.x % n_hap_blocks) * kWarps;
    const int64_t i = h0 + warp;
    int64_t h = i;
    if (kOrdered && i < n_haps) h = __ldg(hap_order + i);
    if (i < n_haps) {
        const uint2* tile_rec =
            reinterpret_cast<const uint2*>(planes) + ((size_t)tile * (size_t)n_keys) * kTileWords + lane;
        uint2 r = and_keys_global(tile_rec, hap_key, __ldg(hap_off + h), __ldg(hap_off + h + 1));
        const int64_t rem = n_samples - tile * kTileSamples - 32 * (int64_t)lane;
        const uint32_t vm = rem >= 32 ? 0xffffffffu : (rem <= 0 ? 0u : ((1u << rem) - 1u));
        r.x &= vm;
        r.y &= vm;
        s_rec[warp][lane] = r;
    }
    if (kOrdered) {
        __syncwarp();
        if (lane == 0 && i < n_haps) {
            const size_t at = ((size_t)(dst_tile0 + tile) * (size_t)n_haps + (size_t)h) * kTileWords;  // in uint2
            const uint32_t src = (uint32_t)__cvta_generic_to_shared(&s_rec[warp][0]);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#pragma unroll
            for (int d = 0; d < HT_MAX_PEERS; ++d)
                if (d < n_dsts)
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(reinterpret_cast<uint2*>(dsts.p[d]) + at),
                                 "r"(src), "n"(kTileWords * 8)
                                 : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        return;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t bytes = (uint32_t)min((int64_t)kWarps, n_haps - h0) * (kTileWords * 8);
        const size_t at = ((size_t)(dst_tile0 + tile) * (size_t)n_haps + (size_t)h0) * kTileWords;  // in uint2
        const uint32_t src = (uint32_t)__cvta_generic_to_shared(&s_rec[0][0]);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the warps' writes, visible to the copy unit
#pragma unroll
        for (int d = 0; d < HT_MAX_PEERS; ++d)  // unrolled: the pointers stay in the parameter bank
            if (d < n_dsts)
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(reinterpret_cast<uint2*>(dsts.p[d]) + at),
                             "r"(src), "r"(bytes)
                             : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // shared memory must outlive the reads
    }
}

// counts[h] = number of set bits of haplotype h over all tiles, both strands.
__global__ void __launch_bounds__(kThreads)
count_bits_kernel(const uint32_t* __restrict__ bits, int64_t n_tiles, int64_t n_haps,
                  unsigned long long* __restrict__ counts) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t h = (int64_t)blockIdx.x * kWarps + warp;
    if (h >= n_haps) return;
    unsigned long long c = 0;
    for (int64_t t = blockIdx.y; t < n_tiles; t += gridDim.y) {
        const uint2 v = __ldg(reinterpret_cast<const uint2*>(bits) + ((size_t)t * n_haps + h) * kTileWords + lane);
        c += __popc(v.x) + __popc(v.y);
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
    if (lane == 0 && c) atomicAdd(counts + h, c);
}

static int check_common(const uint32_t* planes, int64_t n_samples, int64_t n_keys,
                        const uint32_t* hap_off, const uint32_t* hap_key, int64_t n_haps) {
    if (n_samples < 0 || n_keys < 0 || n_haps < 0) return bad_arg("negative size");
    if (n_haps > 0 && !hap_off) return bad_arg("hap_off is NULL");
    if (n_keys > 0 && !planes) return bad_arg("planes is NULL");
    if (n_keys > 0 && !hap_key) return bad_arg("hap_key is NULL");
    if (n_keys > 0 && (reinterpret_cast<uintptr_t>(planes) & 15))
        return bad_arg("planes must be 16-byte aligned");
    if (n_keys > 0xffffffffLL) {
        set_error("too many keys (%lld)", (long long)n_keys);
        return HT_ERR_UNSUPPORTED;
    }
    return HT_OK;
}

}  // namespace ht

extern "C" {

static int transform_u8_impl(const uint32_t* planes, int64_t n_samples, int64_t n_keys,
                             const uint32_t* hap_off, const uint32_t* hap_key, const uint32_t* quad_off16,
                             const uint32_t* quads, int64_t n_haps, uint8_t* out, int64_t out_s_samp,
                             uint64_t* counts, void* stream, const uint32_t* blk_key_lo,
                             const uint32_t* blk_key_cnt, int64_t stage_lines) {
    using namespace ht;
    if (int rc = check_common(planes, n_samples, n_keys, hap_off, hap_key, n_haps)) return rc;
    if (counts && n_haps > 0)
        HT_CUDA_TRY(cudaMemsetAsync(counts, 0, sizeof(uint64_t) * (size_t)n_haps, (cudaStream_t)stream));
    if (n_samples == 0 || n_haps == 0) return HT_OK;
    unsigned long long* cnt = reinterpret_cast<unsigned long long*>(counts);
    if (!out) return bad_arg("out is NULL");
    if (out_s_samp < 2 * n_haps) return bad_arg("out_s_samp < 2 * n_haps");
    if ((quad_off16 == nullptr) != (quads == nullptr)) return bad_arg("quad_off16 and quads go together");
    if (quads && (reinterpret_cast<uintptr_t>(quads) & 15)) return bad_arg("quads must be 16-byte aligned");
    if (n_keys == 0 || n_keys >= kMaxStagedKeys) quads = nullptr;  // record offsets need key * 256 < 2^32
    const uint4* quads4 = reinterpret_cast<const uint4*>(quads);
    if (!quads4) quad_off16 = nullptr;
    // with no keys at all every haplotype is empty; the kernel then never touches `planes`
    const int64_t n_tiles = ht_num_tiles(n_samples);
    const int64_t n_hb = (n_haps + kHapBlock - 1) / kHapBlock;
    if (n_hb > 0x7fffffff) {
        set_error("too many haplotypes (%lld)", (long long)n_haps);
        return HT_ERR_UNSUPPORTED;
    }
    const uintptr_t addr = reinterpret_cast<uintptr_t>(out);
    const int vec = ((addr | (uintptr_t)out_s_samp) & 15) == 0 ? 16
                    : ((addr | (uintptr_t)out_s_samp) & 1) == 0 ? 2 : 1;
    if (quads4 && vec == 16 && transform_half_enabled())
        return launch_transform_u8_half(planes, n_samples, n_keys, quad_off16, quads4, n_haps, out, out_s_samp, cnt,
                                        (cudaStream_t)stream, blk_key_lo, blk_key_cnt, stage_lines);
    // a grid holds at most 2^31-1 blocks: launch ranges of tile groups
    const int64_t groups_per_launch = max((int64_t)1, (int64_t)0x7fffffff / n_hb);
    const int64_t tiles_per_launch = groups_per_launch * kTilesPerCta;
    for (int64_t t0 = 0; t0 < n_tiles; t0 += tiles_per_launch) {
        const int64_t nt = min(tiles_per_launch, n_tiles - t0);
        const dim3 grid((unsigned)(((nt + kTilesPerCta - 1) / kTilesPerCta) * n_hb));
        cudaStream_t st = (cudaStream_t)stream;
        if (vec == 16)
            transform_u8_kernel<16><<<grid, kThreads, 0, st>>>(planes, n_samples, n_keys, hap_off, hap_key, quad_off16, quads4,
                                                             n_haps, (uint32_t)n_hb, t0, out, out_s_samp, cnt);
        else if (vec == 2)
            transform_u8_kernel<2><<<grid, kThreads, 0, st>>>(planes, n_samples, n_keys, hap_off, hap_key, quad_off16, quads4,
                                                            n_haps, (uint32_t)n_hb, t0, out, out_s_samp, cnt);
        else
            transform_u8_kernel<1><<<grid, kThreads, 0, st>>>(planes, n_samples, n_keys, hap_off, hap_key, quad_off16, quads4,
                                                            n_haps, (uint32_t)n_hb, t0, out, out_s_samp, cnt);
        HT_CUDA_TRY(cudaGetLastError());
    }
    return HT_OK;
}

int ht_transform_u8(const uint32_t* planes, int64_t n_samples, int64_t n_keys,
                    const uint32_t* hap_off, const uint32_t* hap_key, const uint32_t* quad_off16,
                    const uint32_t* quads, int64_t n_haps, uint8_t* out, int64_t out_s_samp,
                    uint64_t* counts, void* stream) {
    return transform_u8_impl(planes, n_samples, n_keys, hap_off, hap_key, quad_off16, quads, n_haps, out, out_s_samp,
                             counts, stream, nullptr, nullptr, 0);
}

int ht_transform_u8_staged(const uint32_t* planes, int64_t n_samples, int64_t n_keys,
                           const uint32_t* hap_off, const uint32_t* hap_key, const uint32_t* quad_off16,
                           const uint32_t* quads, const uint32_t* blk_key_lo, const uint32_t* blk_key_cnt,
                           int64_t stage_lines, int64_t n_haps, uint8_t* out, int64_t out_s_samp,
                           uint64_t* counts, void* stream) {
    if (stage_lines < 0 || stage_lines > HT_STAGE_LINES_MAX) return ht::bad_arg("stage_lines out of range (0..HT_STAGE_LINES_MAX)");
    if (stage_lines > 0 && (!blk_key_lo || !blk_key_cnt)) return ht::bad_arg("blk_key_lo / blk_key_cnt are NULL");
    return transform_u8_impl(planes, n_samples, n_keys, hap_off, hap_key, quad_off16, quads, n_haps, out, out_s_samp,
                             counts, stream, blk_key_lo, blk_key_cnt, stage_lines);
}

int ht_transform_bits(const uint32_t* planes, int64_t n_samples, int64_t n_keys,
                      const uint32_t* hap_off, const uint32_t* hap_key, int64_t n_haps,
                      uint32_t* out_bits, void* stream) {
    using namespace ht;
    if (int rc = check_common(planes, n_samples, n_keys, hap_off, hap_key, n_haps)) return rc;
    if (n_samples == 0 || n_haps == 0) return HT_OK;
    if (!out_bits) return bad_arg("out_bits is NULL");
    const int64_t n_tiles = ht_num_tiles(n_samples);
    const int64_t n_hb = (n_haps + kWarps - 1) / kWarps;
    if (n_hb > 0x7fffffff) {
        set_error("too many haplotypes (%lld)", (long long)n_haps);
        return HT_ERR_UNSUPPORTED;
    }
    const int64_t tiles_per_launch = max((int64_t)1, (int64_t)0x7fffffff / n_hb);
    for (int64_t t0 = 0; t0 < n_tiles; t0 += tiles_per_launch) {
        const int64_t nt = min(tiles_per_launch, n_tiles - t0);
        transform_bits_kernel<<<dim3((unsigned)(nt * n_hb)), kThreads, 0, (cudaStream_t)stream>>>(
            planes, n_samples, n_keys, hap_off, hap_key, n_haps, (uint32_t)n_hb, t0, out_bits);
        HT_CUDA_TRY(cudaGetLastError());
    }
    return HT_OK;
}

int ht_transform_bits_gather(const uint32_t* planes, int64_t n_samples, int64_t n_keys,
                             const uint32_t* hap_off, const uint32_t* hap_key, const uint32_t* hap_order, int64_t n_haps,
                             uint32_t* const* dsts, int32_t n_dsts, int64_t dst_tile0, void* stream) {
    using namespace ht;
    if (int rc = check_common(planes, n_samples, n_keys, hap_off, hap_key, n_haps)) return rc;
    if (n_dsts < 1 || n_dsts > HT_MAX_PEERS) return bad_arg("n_dsts must be 1..HT_MAX_PEERS");
    if (dst_tile0 < 0) return bad_arg("negative dst_tile0");
    if (!dsts) return bad_arg("dsts is NULL");
    PeerDsts pd;
    for (int d = 0; d < HT_MAX_PEERS; ++d) pd.p[d] = d < n_dsts ? dsts[d] : nullptr;
    for (int d = 0; d < n_dsts; ++d)
        if (!pd.p[d] || (reinterpret_cast<uintptr_t>(pd.p[d]) & 15)) return bad_arg("a destination is NULL or not 16-byte aligned");
    if (n_samples == 0 || n_haps == 0) return HT_OK;
    const int64_t n_tiles = ht_num_tiles(n_samples);
    const int64_t n_hb = (n_haps + kWarps - 1) / kWarps;
    if (n_hb > 0x7fffffff) {
        set_error("too many haplotypes (%lld)", (long long)n_haps);
        return HT_ERR_UNSUPPORTED;
    }
    const int64_t tiles_per_launch = max((int64_t)1, (int64_t)0x7fffffff / n_hb);
    for (int64_t t0 = 0; t0 < n_tiles; t0 += tiles_per_launch) {
        const int64_t nt = min(tiles_per_launch, n_tiles - t0);
        if (hap_order)
            transform_bits_gather_kernel<true><<<dim3((unsigned)(nt * n_hb)), kThreads, 0, (cudaStream_t)stream>>>(
                planes, n_samples, n_keys, hap_off, hap_key, hap_order, n_haps, (uint32_t)n_hb, t0, pd, n_dsts, dst_tile0);
        else
            transform_bits_gather_kernel<false><<<dim3((unsigned)(nt * n_hb)), kThreads, 0, (cudaStream_t)stream>>>(
                planes, n_samples, n_keys, hap_off, hap_key, nullptr, n_haps, (uint32_t)n_hb, t0, pd, n_dsts, dst_tile0);
        HT_CUDA_TRY(cudaGetLastError());
    }
    return HT_OK;
}

int ht_peer_alloc(size_t bytes, void** ptr, void* handle) {
    using namespace ht;
    if (!ptr || !handle || bytes == 0) return bad_arg("ht_peer_alloc: NULL pointer or empty buffer");
    static_assert(sizeof(cudaIpcMemHandle_t) == HT_PEER_HANDLE_BYTES, "handle size");
    void* p = nullptr;
    HT_CUDA_TRY(cudaMalloc(&p, bytes));
    cudaError_t e = cudaMemset(p, 0, bytes);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return cuda_fail(e, "ht_peer_alloc");
    }
    memcpy(handle, &h, sizeof(h));
    *ptr = p;
    return HT_OK;
}

int ht_peer_open(const void* handle, void** ptr) {
    using namespace ht;
    if (!ptr || !handle) return bad_arg("ht_peer_open: NULL pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    HT_CUDA_TRY(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return HT_OK;
}

int ht_peer_close(void* ptr) {
    using namespace ht;
    if (!ptr) return bad_arg("ht_peer_close: NULL");
    HT_CUDA_TRY(cudaIpcCloseMemHandle(ptr));
    return HT_OK;
}

int ht_peer_free(void* ptr) {
    using namespace ht;
    if (!ptr) return bad_arg("ht_peer_free: NULL");
    HT_CUDA_TRY(cudaFree(ptr));
    return HT_OK;
}

int ht_count_bits(const uint32_t* bits, int64_t n_samples, int64_t n_haps, uint64_t* counts,
                  void* stream) {
    using namespace ht;
    if (n_samples < 0 || n_haps < 0) return bad_arg("negative size");
    if (n_haps == 0) return HT_OK;
    if (!counts) return bad_arg("counts is NULL");
    HT_CUDA_TRY(cudaMemsetAsync(counts, 0, sizeof(uint64_t) * (size_t)n_haps, (cudaStream_t)stream));
    if (n_samples == 0) return HT_OK;
    if (!bits) return bad_arg("bits is NULL");
    const int64_t n_tiles = ht_num_tiles(n_samples);
    const int64_t gx = (n_haps + kWarps - 1) / kWarps;
    if (gx > 0x7fffffff) {
        set_error("too many haplotypes (%lld)", (long long)n_haps);
        return HT_ERR_UNSUPPORTED;
    }
    const unsigned gy = (unsigned)min((int64_t)64, n_tiles);
    count_bits_kernel<<<dim3((unsigned)gx, gy), kThreads, 0, (cudaStream_t)stream>>>(
        bits, n_tiles, n_haps, reinterpret_cast<unsigned long long*>(counts));
    HT_CUDA_TRY(cudaGetLastError());
    return HT_OK;
}

}  // extern "C"
