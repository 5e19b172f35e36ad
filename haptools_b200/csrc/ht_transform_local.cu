// K2, two-phase ("local") form of the haplotype transform (sm_100a).
//
// Same result as ht_transform_u8 -- the reference's loop
//     hap_gts.data[:, i] = np.all(equality_arr[:, idxs[i]], axis=1)
// (haptools/data/haplotypes.py:1411-1412; haptools/transform.py:144-148) -- with fewer L2
// accesses when haplotypes share alleles (K = sum of list lengths >> A = distinct keys):
//
// The one-kernel transform reads one 256-byte (tile, key) record from L2 per haplotype
// allele, and on B200 L2 slice throughput, not HBM, is what that kernel saturates (DESIGN.md
// section 5).  A haplotype is a set of NEARBY variants, and keys are numbered by variant
// column, so haplotypes ordered by their first key and cut into blocks touch a short
// contiguous run of records per block (planner: haptools_b200/plan.py build_local_tables).
//
//   phase 1  and_local_bits_kernel: CTA = (tile, block of <= 128 neighbouring haplotypes).
//            One TMA bulk copy (cp.async.bulk + mbarrier) brings the block's run of records
//            (<= dmax * 256 B) into shared memory; the key lists of the block are staged next
//            to it; every warp ANDs its haplotypes out of shared memory (two records per
//            LDS.128, half-warp each) and writes the bit-packed 256-byte result record at the
//            haplotype's position in the CALLER's order -- the public packed format of
//            ht_transform_bits.
//   phase 2  unpack_bits_u8_kernel: bit-packed records -> (n, H, 2) bytes: the transpose /
//            exchange / expand / 16-byte streaming store path of transform_u8_kernel, fed by
//            16 independent 256-byte record reads per warp.
//
// Per tile, in 256-byte records moved through L2: K + A (one kernel) vs
// sum(block runs) + A + 4 H (two phases).  The planner picks (LocalTables.worthwhile).
#include "ht_transform.cuh"

namespace ht {

constexpr int kLocalMaxBlock = 128;    // haplotypes per block at most (planner: LOCAL_MAX_BLOCK)
constexpr int kLocalStageKeys = 4096;  // relative keys (uint16) staged per block: 8 KB
constexpr int kLocalChunkRecs = 16;    // records per bulk copy (4 KB), one copy per lane of warp 0

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "HT_MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra HT_MBAR_DONE;\n"
        "bra HT_MBAR_WAIT;\n"
        "HT_MBAR_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// TMA bulk copy global -> shared, completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

__device__ __forceinline__ uint4 and4(uint4 a, uint4 b) {
    return make_uint4(a.x & b.x, a.y & b.y, a.z & b.z, a.w & b.w);
}

// ---- phase 1 ---------------------------------------------------------------------------
// CTA = one block of neighbouring haplotypes x kAndTiles consecutive tiles: the block's key lists
// are staged once, then per tile one TMA bulk copy brings the block's run of records in, the warps
// AND their haplotypes out of shared memory, and the next tile's copy is issued.  Three CTAs per
// SM interleave copy waits with AND phases.
// bits: record (tl * n_haps + h) for group-local tile tl and haplotype h (caller's order).
constexpr int kAndTiles = 8;

__global__ void __launch_bounds__(kThreads, 3)
and_local_bits_kernel(const uint32_t* __restrict__ planes, int64_t n_samples, int64_t n_keys,
                      const uint32_t* __restrict__ blk_hap_off, const uint32_t* __restrict__ blk_key_lo,
                      const uint32_t* __restrict__ blk_key_cnt, const uint32_t* __restrict__ s_hap_off,
                      const uint32_t* __restrict__ s_hap_key, const uint32_t* __restrict__ hap_order,
                      int64_t n_haps, uint32_t n_blocks, int64_t tile0, int64_t n_tiles, uint32_t dmax,
                      uint32_t* __restrict__ bits, unsigned long long* __restrict__ counts) {
    extern __shared__ __align__(128) uint8_t s_dyn[];
    uint8_t* s_planes = s_dyn;                                                   // dmax * 256 B
    uint16_t* s_keys = reinterpret_cast<uint16_t*>(s_dyn + (size_t)dmax * 256);  // kLocalStageKeys
    __shared__ uint32_t s_off[kLocalMaxBlock + 1];
    __shared__ uint32_t s_dst[kLocalMaxBlock];
    __shared__ __align__(8) uint64_t s_bar;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t b = blockIdx.x % n_blocks;
    const int64_t tl_first = (int64_t)(blockIdx.x / n_blocks) * kAndTiles;
    const int n_my_tiles = (int)min((int64_t)kAndTiles, n_tiles - tl_first);
    const uint32_t j0 = __ldg(blk_hap_off + b), nh = __ldg(blk_hap_off + b + 1) - j0;
    const uint32_t lo = __ldg(blk_key_lo + b), cnt = __ldg(blk_key_cnt + b);
    const size_t tile_stride = (size_t)n_keys * (kRecordWords * 4);
    const char* run0 = reinterpret_cast<const char*>(planes) + (size_t)(tile0 + tl_first) * tile_stride;

    // the run of tile ti -> shared memory: one bulk copy of 16 records per lane of warp 0
    auto issue_copy = [&](int ti) {
        if (lane == 0) mbar_expect_tx(&s_bar, cnt * 256u);
        __syncwarp();
        const char* src = run0 + (size_t)ti * tile_stride + (size_t)lo * 256;
        for (uint32_t r = lane * kLocalChunkRecs; r < cnt; r += 32 * kLocalChunkRecs) {
            const uint32_t nrec = min((uint32_t)kLocalChunkRecs, cnt - r);
            bulk_g2s(s_planes + (size_t)r * 256, src + (size_t)r * 256, nrec * 256u, &s_bar);
        }
    };
    if (warp == 0) {
        if (lane == 0) mbar_init(&s_bar, 1);
        __syncwarp();
        if (cnt) issue_copy(0);
    }
    // ---- stage the block's lists once ------------------------------------------------------
    for (uint32_t i = tid; i <= nh; i += kThreads) s_off[i] = __ldg(s_hap_off + j0 + i);
    for (uint32_t i = tid; i < nh; i += kThreads) s_dst[i] = __ldg(hap_order + j0 + i);
    __syncthreads();
    const uint32_t kb0 = s_off[0], nk = s_off[nh] - kb0;
    const bool keys_staged = cnt > 0 && nk <= (uint32_t)kLocalStageKeys;
    if (keys_staged)
        for (uint32_t i = tid; i < nk; i += kThreads) s_keys[i] = (uint16_t)__ldg(s_hap_key + kb0 + i);
    __syncthreads();

    // AND: warp per haplotype, half-warp per record (lane & 15 = 16 bytes of it)
    const int half = lane >> 4;
    const uint32_t my16 = 16u * (lane & 15);
    const uint8_t* my_planes = s_planes + my16;

#pragma unroll 1
    for (int ti = 0; ti < n_my_tiles; ++ti) {
        const int64_t tile = tile0 + tl_first + ti;
        if (cnt) mbar_wait(&s_bar, ti & 1);
        // validity mask of my two words (samples >= n_samples read as 0, like the planes)
        const int64_t rem = n_samples - tile * kTileSamples - 64 * (int64_t)(lane & 15);
        const uint32_t vm0 = rem >= 32 ? 0xffffffffu : (rem <= 0 ? 0u : ((1u << (int)rem) - 1u));
        const uint32_t vm1 = rem >= 64 ? 0xffffffffu : (rem <= 32 ? 0u : ((1u << (int)(rem - 32)) - 1u));
        char* out_bytes = reinterpret_cast<char*>(bits) + ((size_t)(tl_first + ti) * (size_t)n_haps) * (kRecordWords * 4) + my16;
        const char* tile_bytes = run0 + (size_t)ti * tile_stride;

#pragma unroll 1
        for (uint32_t hl = warp; hl < nh; hl += kWarps) {
            const uint32_t kb = s_off[hl] - kb0, len = s_off[hl + 1] - s_off[hl];
            uint4 r = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
            if (len) {
                if (keys_staged && len <= 32u) {
                    // the list sits in one register per lane; every record address comes over a
                    // shuffle, so the LDS.128 of a haplotype are independent of each other
                    const uint32_t mine = (uint32_t)s_keys[kb + min((uint32_t)lane, len - 1)] * 256u;
#pragma unroll
                    for (uint32_t t = 0; t < 16; t += 4) {
                        if (2 * t >= len) break;  // warp-uniform
                        uint4 v[4];
#pragma unroll
                        for (uint32_t u = 0; u < 4; ++u) {
                            const uint32_t off = __shfl_sync(0xffffffffu, mine, min(2 * (t + u) + half, len - 1));
                            v[u] = *reinterpret_cast<const uint4*>(my_planes + off);
                        }
                        r = and4(r, and4(and4(v[0], v[1]), and4(v[2], v[3])));
                    }
                } else if (cnt) {  // long lists (or not staged): keys one by one, records from shared memory
                    const uint32_t last = kb + len - 1;
#pragma unroll 2
                    for (uint32_t t = kb + half; t < kb + len + half; t += 2) {
                        const uint32_t k = keys_staged ? (uint32_t)s_keys[min(t, last)] : __ldg(s_hap_key + kb0 + min(t, last));
                        r = and4(r, *reinterpret_cast<const uint4*>(my_planes + k * 256u));
                    }
                } else {  // block whose run does not fit: global keys, records straight from L2
                    const uint32_t last = kb + len - 1;
#pragma unroll 2
                    for (uint32_t t = kb + half; t < kb + len + half; t += 2) {
                        const uint32_t k = __ldg(s_hap_key + kb0 + min(t, last));
                        r = and4(r, __ldg(reinterpret_cast<const uint4*>(tile_bytes + (size_t)k * 256 + my16)));
                    }
                }
                r.x &= __shfl_xor_sync(0xffffffffu, r.x, 16);
                r.y &= __shfl_xor_sync(0xffffffffu, r.y, 16);
                r.z &= __shfl_xor_sync(0xffffffffu, r.z, 16);
                r.w &= __shfl_xor_sync(0xffffffffu, r.w, 16);
            }
            r.x &= vm0;
            r.y &= vm0;
            r.z &= vm1;
            r.w &= vm1;
            const uint32_t h = s_dst[hl];
            if (counts) {
                // allele counts for check_maf (haptools/data/genotypes.py:559-625): both half-warps
                // hold the same words, so count lanes 0-15 only
                const uint32_t pc = half ? 0u : __popc(r.x) + __popc(r.y) + __popc(r.z) + __popc(r.w);
                const uint32_t c = __reduce_add_sync(0xffffffffu, pc);
                if (lane == 0 && c) atomicAdd(counts + h, (unsigned long long)c);
            }
            if (!half) *reinterpret_cast<uint4*>(out_bytes + (size_t)h * (kRecordWords * 4)) = r;
        }
        if (cnt && ti + 1 < n_my_tiles) {
            __syncthreads();  // everybody is done with this tile's records
            if (warp == 0) issue_copy(ti + 1);
        }
    }
}

// ---- phase 2 ---------------------------------------------------------------------------
// CTA = (tile, 128 haplotypes); warp = 16 haplotypes; lane = one 32-sample word, both strands.
// One tile per CTA on purpose: CTAs that loop over several tiles drift apart and the write
// stream loses its DRAM locality (8 tiles per CTA: 10.0 ms instead of 8.5 ms at n = 200 k, with
// register prefetch of the next tile's records or with a TMA prefetch alike; an L2 bulk prefetch
// of the next tile's records from a one-tile CTA changes nothing either: 8.57 vs 8.36 ms).
template <int kOutVec>
__global__ void __launch_bounds__(kThreads, 4)
unpack_bits_u8_kernel(const uint32_t* __restrict__ bits, int64_t n_samples, int64_t n_haps,
                      uint32_t n_hap_blocks, int64_t tile0, int64_t n_tiles, uint8_t* __restrict__ out,
                      int64_t out_s_samp) {
    __shared__ __align__(16) uint32_t s_u[kWarps * kUPitch];  // 32.9 KB
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t tl = blockIdx.x / n_hap_blocks, tile = tile0 + tl;
    const int64_t h0 = (int64_t)(blockIdx.x % n_hap_blocks) * kHapBlock;
    const int n_blk_haps = (int)min((int64_t)kHapBlock, n_haps - h0);
    const int hl_begin = warp * kHapsPerWarp;

    // 16 independent coalesced 256-byte reads per warp; records are read exactly once
    const uint2* rec = reinterpret_cast<const uint2*>(bits) +
                       ((size_t)tl * (size_t)n_haps + (size_t)(h0 + hl_begin)) * kTileWords + lane;
    uint32_t a[32];  // a[2*i + c] = haplotype i of this warp, strand c
#pragma unroll
    for (int i = 0; i < kHapsPerWarp; ++i) {
        uint2 r = make_uint2(0u, 0u);
        if (hl_begin + i < n_blk_haps)
            asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(rec + (size_t)i * kTileWords));
        a[2 * i] = r.x;
        a[2 * i + 1] = r.y;
    }
    transpose32_dev(a);

    // exchange + expand + store: identical to transform_u8_kernel (ht_transform.cu) step 3
    uint32_t* my_u = s_u + warp * kUPitch + 32 * lane;
#pragma unroll
    for (int j = 0; j < 32; ++j) my_u[(j + lane) & 31] = a[j];
    __syncthreads();

    const int half = threadIdx.x & 1;
    const int bw = (threadIdx.x >> 1) & 7;
    const int c0 = (threadIdx.x >> 5) * 2 + ((threadIdx.x >> 4) & 1);
    const int hl0 = bw * kHapsPerWarp + half * 8;
    const int n_my_haps = min(8, n_blk_haps - hl0);
    const uint32_t* u_col = s_u + bw * kUPitch;
    const int shift = 16 * half;
    if (n_my_haps <= 0) return;
    const int64_t s_base = tile * kTileSamples;
    const int rows = (int)min((int64_t)kTileSamples, n_samples - s_base);
    uint8_t* p = out + 2 * (h0 + hl0) + (s_base + c0) * out_s_samp;
    const int64_t p_step = 16 * out_s_samp;

    if (kOutVec == 16 && n_my_haps == 8 && rows == kTileSamples) {
#pragma unroll 4
        for (int L = 0; L < 32; ++L) {
            const int t = (c0 + L) & 31;
            const uint32_t w0 = u_col[32 * L + t], w1 = u_col[32 * L + (t ^ 16)];
            st_stream_v4(p, expand16((w0 >> shift) & 0xffffu));
            st_stream_v4(p + p_step, expand16((w1 >> shift) & 0xffffu));
            p += 2 * p_step;
        }
        return;
    }
#pragma unroll 1
    for (int it = 0; it < kTileSamples / 16; ++it, p += p_step) {
        const int sl = it * 16 + c0;
        if (sl >= rows) break;
        const int L = it >> 1, j = ((it & 1) << 4) + c0;
        const uint32_t wbits = (u_col[32 * L + ((j + L) & 31)] >> shift) & 0xffffu;
        if (kOutVec == 16 && n_my_haps == 8) {
            st_stream_v4(p, expand16(wbits));
        } else if (kOutVec >= 2) {
            for (int q = 0; q < n_my_haps; ++q) {
                const uint32_t two = (wbits >> (2 * q)) & 3u;
                *reinterpret_cast<uint16_t*>(p + 2 * q) = (uint16_t)((two & 1u) | ((two & 2u) << 7));
            }
        } else {
            for (int q = 0; q < 2 * n_my_haps; ++q) p[q] = (uint8_t)((wbits >> q) & 1u);
        }
    }
}

constexpr int kUnpackTiles = 1;

static int launch_unpack(const uint32_t* bits, int64_t n_samples, int64_t n_haps, int64_t tile0,
                         int64_t n_tiles, uint8_t* out, int64_t out_s_samp, cudaStream_t st) {
    const int64_t n_hb = (n_haps + kHapBlock - 1) / kHapBlock;
    const uintptr_t addr = reinterpret_cast<uintptr_t>(out);
    const int vec = ((addr | (uintptr_t)out_s_samp) & 15) == 0 ? 16
                    : ((addr | (uintptr_t)out_s_samp) & 1) == 0 ? 2 : 1;
    const int64_t groups_per_launch = max((int64_t)1, (int64_t)0x7fffffff / n_hb);
    const int64_t tiles_per_launch = groups_per_launch * kUnpackTiles;
    for (int64_t t0 = 0; t0 < n_tiles; t0 += tiles_per_launch) {
        const int64_t nt = min(tiles_per_launch, n_tiles - t0);
        const dim3 grid((unsigned)(((nt + kUnpackTiles - 1) / kUnpackTiles) * n_hb));
        const uint32_t* src = bits + (size_t)t0 * (size_t)n_haps * kRecordWords;
        if (vec == 16)
            unpack_bits_u8_kernel<16><<<grid, kThreads, 0, st>>>(src, n_samples, n_haps, (uint32_t)n_hb, tile0 + t0, nt, out, out_s_samp);
        else if (vec == 2)
            unpack_bits_u8_kernel<2><<<grid, kThreads, 0, st>>>(src, n_samples, n_haps, (uint32_t)n_hb, tile0 + t0, nt, out, out_s_samp);
        else
            unpack_bits_u8_kernel<1><<<grid, kThreads, 0, st>>>(src, n_samples, n_haps, (uint32_t)n_hb, tile0 + t0, nt, out, out_s_samp);
        HT_CUDA_TRY(cudaGetLastError());
    }
    return HT_OK;
}

static size_t local_smem_bytes(int64_t dmax) { return (size_t)dmax * 256 + kLocalStageKeys * sizeof(uint16_t); }

static int check_local(const uint32_t* planes, int64_t n_samples, int64_t n_keys,
                       const ht_local_tables* t, int64_t n_haps) {
    if (n_samples < 0 || n_keys < 0 || n_haps < 0) return bad_arg("negative size");
    if (!t) return bad_arg("local tables are NULL");
    if (n_haps == 0) return HT_OK;
    if (t->n_blocks <= 0 || t->n_blocks > 0x7fffffff) return bad_arg("local tables: bad n_blocks");
    if (t->dmax < 1 || t->dmax > 800) return bad_arg("local tables: dmax out of range (1..800)");
    if (!t->blk_hap_off || !t->blk_key_lo || !t->blk_key_cnt || !t->s_hap_off || !t->hap_order)
        return bad_arg("local tables: NULL table");
    if (n_keys > 0 && (!planes || !t->s_hap_key)) return bad_arg("planes / s_hap_key is NULL");
    if (n_keys > 0 && (reinterpret_cast<uintptr_t>(planes) & 15)) return bad_arg("planes must be 16-byte aligned");
    if (n_keys > 0xffffffffLL) {
        set_error("too many keys (%lld)", (long long)n_keys);
        return HT_ERR_UNSUPPORTED;
    }
    return HT_OK;
}

static int launch_and_local(const uint32_t* planes, int64_t n_samples, int64_t n_keys,
                            const ht_local_tables* t, int64_t n_haps, int64_t tile0, int64_t n_tiles,
                            uint32_t* bits, unsigned long long* counts, cudaStream_t st) {
    const size_t smem = local_smem_bytes(t->dmax);
    // per launch: the attribute is per device, and a caller may drive several devices from one thread
    HT_CUDA_TRY(cudaFuncSetAttribute(and_local_bits_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t groups_per_launch = max((int64_t)1, (int64_t)0x7fffffff / t->n_blocks);
    const int64_t tiles_per_launch = groups_per_launch * kAndTiles;
    for (int64_t t0 = 0; t0 < n_tiles; t0 += tiles_per_launch) {
        const int64_t nt = min(tiles_per_launch, n_tiles - t0);
        const unsigned grid = (unsigned)(((nt + kAndTiles - 1) / kAndTiles) * t->n_blocks);
        and_local_bits_kernel<<<grid, kThreads, smem, st>>>(
            planes, n_samples, n_keys, t->blk_hap_off, t->blk_key_lo, t->blk_key_cnt, t->s_hap_off,
            t->s_hap_key, t->hap_order, n_haps, (uint32_t)t->n_blocks, tile0 + t0, nt, (uint32_t)t->dmax,
            bits + (size_t)t0 * (size_t)n_haps * kRecordWords, counts);
        HT_CUDA_TRY(cudaGetLastError());
    }
    return HT_OK;
}

}  // namespace ht

extern "C" {

size_t ht_local_workspace_bytes(int64_t n_samples, int64_t n_haps) {
    if (n_samples <= 0 || n_haps <= 0) return 0;
    const int64_t tiles = ht_num_tiles(n_samples);
    // 64 tiles per group keep both launches long (>= 10 waves of CTAs) and the workspace small
    return (size_t)(tiles < 64 ? tiles : 64) * (size_t)n_haps * HT_RECORD_WORDS * sizeof(uint32_t);
}

int ht_transform_bits_local(const uint32_t* planes, int64_t n_samples, int64_t n_keys,
                            const ht_local_tables* tables, int64_t n_haps, uint32_t* out_bits,
                            uint64_t* counts, void* stream) {
    using namespace ht;
    if (int rc = check_local(planes, n_samples, n_keys, tables, n_haps)) return rc;
    if (counts && n_haps > 0)
        HT_CUDA_TRY(cudaMemsetAsync(counts, 0, sizeof(uint64_t) * (size_t)n_haps, (cudaStream_t)stream));
    if (n_samples == 0 || n_haps == 0) return HT_OK;
    if (!out_bits) return bad_arg("out_bits is NULL");
    return launch_and_local(planes, n_samples, n_keys, tables, n_haps, 0, ht_num_tiles(n_samples), out_bits,
                            reinterpret_cast<unsigned long long*>(counts), (cudaStream_t)stream);
}

int ht_unpack_bits_u8(const uint32_t* bits, int64_t n_samples, int64_t n_haps, uint8_t* out,
                      int64_t out_s_samp, void* stream) {
    using namespace ht;
    if (n_samples < 0 || n_haps < 0) return bad_arg("negative size");
    if (n_samples == 0 || n_haps == 0) return HT_OK;
    if (!bits || !out) return bad_arg("bits / out is NULL");
    if (reinterpret_cast<uintptr_t>(bits) & 7) return bad_arg("bits must be 8-byte aligned");
    if (out_s_samp < 2 * n_haps) return bad_arg("out_s_samp < 2 * n_haps");
    if ((n_haps + kHapBlock - 1) / kHapBlock > 0x7fffffff) return bad_arg("too many haplotypes");
    return launch_unpack(bits, n_samples, n_haps, 0, ht_num_tiles(n_samples), out, out_s_samp, (cudaStream_t)stream);
}

int ht_transform_u8_local(const uint32_t* planes, int64_t n_samples, int64_t n_keys,
                          const ht_local_tables* tables, int64_t n_haps, uint8_t* out,
                          int64_t out_s_samp, uint64_t* counts, void* workspace,
                          size_t workspace_bytes, void* stream) {
    using namespace ht;
    if (int rc = check_local(planes, n_samples, n_keys, tables, n_haps)) return rc;
    if (counts && n_haps > 0)
        HT_CUDA_TRY(cudaMemsetAsync(counts, 0, sizeof(uint64_t) * (size_t)n_haps, (cudaStream_t)stream));
    if (n_samples == 0 || n_haps == 0) return HT_OK;
    if (!out) return bad_arg("out is NULL");
    if (out_s_samp < 2 * n_haps) return bad_arg("out_s_samp < 2 * n_haps");
    const size_t tile_bytes = (size_t)n_haps * kRecordWords * sizeof(uint32_t);
    if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 15) || workspace_bytes < tile_bytes)
        return bad_arg("workspace must be 16-byte aligned and hold at least one tile (n_haps * 256 bytes)");
    const int64_t n_tiles = ht_num_tiles(n_samples);
    const int64_t group = (int64_t)min((size_t)n_tiles, workspace_bytes / tile_bytes);
    uint32_t* bits = reinterpret_cast<uint32_t*>(workspace);
    for (int64_t t0 = 0; t0 < n_tiles; t0 += group) {
        const int64_t nt = min(group, n_tiles - t0);
        if (int rc = launch_and_local(planes, n_samples, n_keys, tables, n_haps, t0, nt, bits,
                                      reinterpret_cast<unsigned long long*>(counts), (cudaStream_t)stream))
            return rc;
        if (int rc = launch_unpack(bits, n_samples, n_haps, t0, nt, out, out_s_samp, (cudaStream_t)stream)) return rc;
    }
    return HT_OK;
}

}  // extern "C"
