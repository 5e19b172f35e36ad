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
#include "ht_common.cuh"

namespace ht {

constexpr int kHapBlock = 128;      // haplotypes per CTA
constexpr int kHapsPerWarp = 16;    // -> 32 result bits (16 haps x 2 strands) per sample
constexpr int kWarps = kHapBlock / kHapsPerWarp;  // 8
constexpr int kThreads = kWarps * 32;             // 256
constexpr int kStageKeys = 3072;    // key indices staged per CTA (12 KB); longer lists are
                                    // read straight from global memory
constexpr int kUPitch = kTileSamples + 4;  // +4 words: the 8 warps' rows land in distinct banks

__device__ __forceinline__ uint2 ld_plane(const uint2* p) {
    // planes are re-read by every haplotype that uses the key: keep them in L1/L2
    return __ldg(p);
}

__device__ __forceinline__ void st_stream_v4(uint8_t* p, uint4 v) {
    // the 1-byte-per-call result is written once and never re-read by this kernel
    asm volatile("st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y),
                 "r"(v.z), "r"(v.w)
                 : "memory");
}

// AND of the (tile, key) records of one haplotype, for this lane's word; both strands.
template <bool kStaged>
__device__ __forceinline__ uint2 and_keys(const uint2* __restrict__ tile_rec,  // + lane
                                          const uint32_t* __restrict__ keys,   // smem or global
                                          uint32_t kb, uint32_t ke) {
    uint32_t r0 = 0xffffffffu, r1 = 0xffffffffu;
    uint32_t j = kb;
    for (; j + 4 <= ke; j += 4) {
        const uint32_t k0 = kStaged ? keys[j] : __ldg(keys + j);
        const uint32_t k1 = kStaged ? keys[j + 1] : __ldg(keys + j + 1);
        const uint32_t k2 = kStaged ? keys[j + 2] : __ldg(keys + j + 2);
        const uint32_t k3 = kStaged ? keys[j + 3] : __ldg(keys + j + 3);
        const uint2 v0 = ld_plane(tile_rec + (size_t)k0 * kTileWords);
        const uint2 v1 = ld_plane(tile_rec + (size_t)k1 * kTileWords);
        const uint2 v2 = ld_plane(tile_rec + (size_t)k2 * kTileWords);
        const uint2 v3 = ld_plane(tile_rec + (size_t)k3 * kTileWords);
        r0 &= (v0.x & v1.x) & (v2.x & v3.x);
        r1 &= (v0.y & v1.y) & (v2.y & v3.y);
    }
    for (; j < ke; ++j) {
        const uint32_t k0 = kStaged ? keys[j] : __ldg(keys + j);
        const uint2 v0 = ld_plane(tile_rec + (size_t)k0 * kTileWords);
        r0 &= v0.x;
        r1 &= v0.y;
    }
    return make_uint2(r0, r1);
}

// out_vec: 16 -> every full 8-haplotype group is stored with one 16-byte store (requires
//          out and out_s_samp to be multiples of 16); 2 -> 2-byte stores only (any layout
//          with even addresses); 1 -> byte stores.
template <int kOutVec>
__global__ void __launch_bounds__(kThreads)
transform_u8_kernel(const uint32_t* __restrict__ planes, int64_t n_samples, int64_t n_keys,
                    const uint32_t* __restrict__ hap_off, const uint32_t* __restrict__ hap_key,
                    int64_t n_haps, uint32_t n_hap_blocks, int64_t tile0,
                    uint8_t* __restrict__ out, int64_t out_s_samp) {
    __shared__ uint32_t s_u[kWarps * kUPitch];  // 32.9 KB
    __shared__ uint32_t s_keys[kStageKeys];     // 12 KB
    __shared__ uint32_t s_off[kHapBlock + 1];

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int64_t tile = tile0 + blockIdx.x / n_hap_blocks;
    const int64_t h0 = (int64_t)(blockIdx.x % n_hap_blocks) * kHapBlock;
    const int n_blk_haps = (int)min((int64_t)kHapBlock, n_haps - h0);

    // ---- stage the block's CSR slice -------------------------------------------------
    for (int i = tid; i <= n_blk_haps; i += kThreads) s_off[i] = __ldg(hap_off + h0 + i);
    __syncthreads();
    const uint32_t k_begin = s_off[0];
    const uint32_t k_count = s_off[n_blk_haps] - k_begin;
    const bool staged = k_count <= (uint32_t)kStageKeys;
    if (staged) {
        for (uint32_t i = tid; i < k_count; i += kThreads) s_keys[i] = __ldg(hap_key + k_begin + i);
    }
    __syncthreads();

    // ---- 1. AND ----------------------------------------------------------------------
    const uint2* tile_rec =
        reinterpret_cast<const uint2*>(planes) + ((size_t)tile * (size_t)n_keys) * kTileWords + lane;
    uint32_t a[32];
#pragma unroll
    for (int i = 0; i < kHapsPerWarp; ++i) {
        const int hl = warp * kHapsPerWarp + i;
        uint2 r = make_uint2(0u, 0u);
        if (hl < n_blk_haps) {
            const uint32_t kb = s_off[hl], ke = s_off[hl + 1];
            r = staged ? and_keys<true>(tile_rec, s_keys, kb - k_begin, ke - k_begin)
                       : and_keys<false>(tile_rec, hap_key, kb, ke);
        }
        a[2 * i] = r.x;
        a[2 * i + 1] = r.y;
    }

    // ---- 2. bits over samples -> bits over (haplotype, strand) ------------------------
    transpose32(a);

    // ---- 3. exchange + expand + store --------------------------------------------------
    // sample 32*L + j of the tile is kept at slot 32*L + ((j + L) & 31): lanes write distinct
    // banks here, and the 16 words one warp reads below fall into 16 distinct banks.
    uint32_t* my_u = s_u + warp * kUPitch + 32 * lane;
#pragma unroll
    for (int j = 0; j < 32; ++j) my_u[(j + lane) & 31] = a[j];
    __syncthreads();

    const int half = lane & 1;        // low / high 16 bits of the word = haps 0-7 / 8-15
    const int b = (lane >> 1) & 7;    // which warp's 16 haplotypes
    const int rsub = lane >> 4;       // 2 sample rows per warp-wide store
    const int hl0 = b * kHapsPerWarp + half * 8;  // first haplotype (in block) of my 16 bytes
    const int n_my_haps = min(8, n_blk_haps - hl0);  // may be <= 0
    if (n_my_haps <= 0) return;  // no barrier below
    const int64_t s_base = tile * kTileSamples;
    const int rows = (int)min((int64_t)kTileSamples, n_samples - s_base);
    uint8_t* out_col = out + 2 * (h0 + hl0);

#pragma unroll 4
    for (int it = 0; it < kTileSamples / (2 * kWarps); ++it) {
        const int sl = it * (2 * kWarps) + warp * 2 + rsub;
        if (sl >= rows) break;  // rows only grow with `it`
        const int L = sl >> 5, j = sl & 31;
        const uint32_t w = s_u[b * kUPitch + 32 * L + ((j + L) & 31)];
        const uint32_t bits = half ? (w >> 16) : (w & 0xffffu);
        uint8_t* p = out_col + (s_base + sl) * out_s_samp;
        if (kOutVec == 16 && n_my_haps == 8) {
            uint4 o;
            o.x = spread4(bits & 15u);
            o.y = spread4((bits >> 4) & 15u);
            o.z = spread4((bits >> 8) & 15u);
            o.w = spread4(bits >> 12);
            st_stream_v4(p, o);
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
    uint2 r = and_keys<false>(tile_rec, hap_key, __ldg(hap_off + h), __ldg(hap_off + h + 1));
    // samples beyond n_samples read as 0, like the planes
    const int64_t rem = n_samples - tile * kTileSamples - 32 * (int64_t)lane;
    const uint32_t vm = rem >= 32 ? 0xffffffffu : (rem <= 0 ? 0u : ((1u << rem) - 1u));
    r.x &= vm;
    r.y &= vm;
    reinterpret_cast<uint2*>(out_bits)[((size_t)tile * (size_t)n_haps + (size_t)h) * kTileWords + lane] = r;
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
    return HT_OK;
}

}  // namespace ht

extern "C" {

int ht_transform_u8(const uint32_t* planes, int64_t n_samples, int64_t n_keys,
                    const uint32_t* hap_off, const uint32_t* hap_key, int64_t n_haps,
                    uint8_t* out, int64_t out_s_samp, void* stream) {
    using namespace ht;
    if (int rc = check_common(planes, n_samples, n_keys, hap_off, hap_key, n_haps)) return rc;
    if (n_samples == 0 || n_haps == 0) return HT_OK;
    if (!out) return bad_arg("out is NULL");
    if (out_s_samp < 2 * n_haps) return bad_arg("out_s_samp < 2 * n_haps");
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
    // a grid holds at most 2^31-1 blocks: launch ranges of tiles
    const int64_t tiles_per_launch = max((int64_t)1, (int64_t)0x7fffffff / n_hb);
    for (int64_t t0 = 0; t0 < n_tiles; t0 += tiles_per_launch) {
        const int64_t nt = min(tiles_per_launch, n_tiles - t0);
        const dim3 grid((unsigned)(nt * n_hb));
        cudaStream_t st = (cudaStream_t)stream;
        if (vec == 16)
            transform_u8_kernel<16><<<grid, kThreads, 0, st>>>(planes, n_samples, n_keys, hap_off, hap_key,
                                                             n_haps, (uint32_t)n_hb, t0, out, out_s_samp);
        else if (vec == 2)
            transform_u8_kernel<2><<<grid, kThreads, 0, st>>>(planes, n_samples, n_keys, hap_off, hap_key,
                                                            n_haps, (uint32_t)n_hb, t0, out, out_s_samp);
        else
            transform_u8_kernel<1><<<grid, kThreads, 0, st>>>(planes, n_samples, n_keys, hap_off, hap_key,
                                                            n_haps, (uint32_t)n_hb, t0, out, out_s_samp);
        HT_CUDA_TRY(cudaGetLastError());
    }
    return HT_OK;
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
