// K1 with local ancestry, TMA-fed (sm_100a) -- `pack_sm2_anc_tma_kernel`.
//
// Same arithmetic and same result as pack_sm2_anc_line_kernel (ht_pack.cu): for every key (allele a <= 1,
// label l < 8) of a genotype column, bit s of its plane = (GT[s] == a) && (ANC[s] == l), i.e.
//   np.equal(allele_arr, gts.data)          haptools/transform.py:137
//   gts.ancestry[:, idxs] == ancestries[i]  haptools/transform.py:146
// What differs is how the two byte matrices reach the arithmetic.  The register-load kernels keep 16 warps per
// SM (the accumulators of two matrices) and every warp alternates between "8 loads in flight" and "nothing in
// flight while it folds them": 62 % of the DRAM bandwidth, stalled on the row loads (profiles/r1/
// stalls_ancestry_line_pack.txt).  Here ONE thread per CTA keeps a ring of shared-memory stages filled with 2-D
// tensor copies (cp.async.bulk.tensor + mbarrier complete_tx), so the bytes in flight per SM are the ring
// (6 x 32 KB), whatever the warps are doing, and the accumulators live in registers.
//
//   item     = (tile, slab of 128 samples = 4 words, block of 64 columns): one box of 128 rows x 128 bytes of
//              each matrix = one 32 KB stage
//   CTA      = persistent, 1 per SM: warp 0 = producer (one elected lane), 2 consumer groups of 4 warps;
//              stage k of the CTA goes to group k % 2
//   phase A  (warp = one word, lane = 4 bytes = 2 columns x 2 strands of each of the word's 32 rows): rows come
//              out of shared memory with conflict-free 128-byte LDS.32; per 8 rows the GT bit and the three label
//              bits are merged into nibbles (two rows per register) and a 4 x 4 bit transpose across 4 registers
//              leaves one byte of each of the 4 planes per byte column -- 6 ALU operations per row and word
//              against 12 for shift-accumulating every plane; PRMT byte transposes make the 32-bit words.
//              Lanes that saw a byte outside {0, 1} / {0..7} re-read their rows from the stage and build the
//              "invalid" plane, so the result is exact for any input.  The stage is released to the producer.
//   phase B  (thread = column x half of the slab's words): plane words change hands through a 10 KB buffer;
//              for every key of its column the thread combines the planes and stores 16 bytes; the two
//              threads of a column write one whole 32-byte sector of the key's record in one instruction.
#include <cuda.h>

#include <cstdlib>

#include "ht_pack.cuh"

namespace ht {

constexpr int kTmaRows = 128;                                  // samples of a slab (4 words)
constexpr int kTmaBoxBytes = 128;                              // 64 columns x 2 strands
constexpr int kTmaCols = kTmaBoxBytes / 2;
constexpr int kTmaMatBytes = kTmaRows * kTmaBoxBytes;          // one matrix's box: 16 KB
constexpr int kTmaStageBytes = 2 * kTmaMatBytes;               // GT box + ANC box
#ifndef HT_TMA_GROUPS
#define HT_TMA_GROUPS 2
#endif
constexpr int kTmaGroups = HT_TMA_GROUPS, kTmaGroupWarps = 4, kTmaGroupThreads = 32 * kTmaGroupWarps;
constexpr int kTmaThreads = 32 * (1 + kTmaGroups * kTmaGroupWarps);  // 288
constexpr int kTmaSlabsPerTile = kTileSamples / kTmaRows;      // 8
constexpr int kPlanePitch = 136;                               // words per (plane, word) row of the hand-over buffer:
                                                               // 128 + 8, so that the two halves of a column (two rows
                                                               // apart) fall into different banks
constexpr int kPlaneBufWords = 5 * 4 * kPlanePitch;            // planes: allele 1, label bits 0..2, invalid
#ifndef HT_TMA_STAGES
#define HT_TMA_STAGES 6
#endif
constexpr int kTmaStages = HT_TMA_STAGES;
constexpr size_t kTmaSmem = (size_t)kTmaStages * kTmaStageBytes + sizeof(uint32_t) * kTmaGroups * kPlaneBufWords +
                            sizeof(uint64_t) * 2 * kTmaStages;

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "HT_TMA_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra HT_TMA_DONE;\n"
        "bra HT_TMA_WAIT;\n"
        "HT_TMA_DONE:\n"
        "}\n" ::"r"(smem_addr(bar)),
        "r"(parity)
        : "memory");
}

// box (c0 = byte column, c1 = row) of a 2-D uint8 tensor -> shared memory, bytes counted on `bar`
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;" ::
            "r"(smem_addr(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_addr(bar)), "l"(policy)
        : "memory");
}

__device__ __forceinline__ void group_barrier(int id) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(kTmaGroupThreads) : "memory");
}

// bits of `b` where `mask` is 0, bits of `a` where it is 1
__device__ __forceinline__ uint32_t bitsel(uint32_t mask, uint32_t a, uint32_t b) { return (a & mask) | (b & ~mask); }

// 4 x 4 bit transpose inside every nibble, across four registers: afterwards bit (4 h + k) of z[b] is what bit
// (4 h + b) of z[k] was (h = 0, 1: the two nibbles of a byte are independent)
__device__ __forceinline__ void nibble_transpose4(uint32_t (&z)[4]) {
    uint32_t t;
    t = ((z[0] >> 1) ^ z[1]) & 0x55555555u; z[1] ^= t; z[0] ^= t << 1;
    t = ((z[2] >> 1) ^ z[3]) & 0x55555555u; z[3] ^= t; z[2] ^= t << 1;
    t = ((z[0] >> 2) ^ z[2]) & 0x33333333u; z[2] ^= t; z[0] ^= t << 2;
    t = ((z[1] >> 2) ^ z[3]) & 0x33333333u; z[3] ^= t; z[1] ^= t << 2;
}

struct TmaItems {
    int64_t n_tiles, total;  // items = n_tiles * 8 * n_cb
    uint32_t n_cb;           // 64-column blocks of a row
    uint32_t group;          // column blocks per group: items run (tile, group, slab, block in group)
    uint32_t chunk_log2;     // 2^chunk_log2 consecutive items go to one CTA
    uint32_t stream_hint;    // 1: L2 evict_first on the matrix loads (measurements only)
};

struct TmaItem {
    uint32_t tile, slab, cb;
    bool valid;
};

// k-th item of this CTA.  Idx = uint32_t when the launch has fewer than 2^32 items (the usual case: 32-bit divisions).
template <class Idx>
__device__ __forceinline__ TmaItem tma_item(const TmaItems& w, uint32_t k) {
    TmaItem out;
    const Idx idx = (((Idx)(k >> w.chunk_log2) * gridDim.x + blockIdx.x) << w.chunk_log2) + (k & ((1u << w.chunk_log2) - 1u));
    out.valid = idx < (Idx)w.total;
    const uint32_t per_tile = kTmaSlabsPerTile * w.n_cb;
    out.tile = (uint32_t)(idx / per_tile);
    const uint32_t r = (uint32_t)(idx - (Idx)out.tile * per_tile);
    const uint32_t per_group = kTmaSlabsPerTile * w.group;
    const uint32_t g = r / per_group, rr = r - g * per_group;
    const uint32_t gsize = min(w.group, w.n_cb - g * w.group);
    out.slab = rr / gsize;
    out.cb = g * w.group + (rr - out.slab * gsize);
    return out;
}

constexpr int kInfoAhead = 4;  // keys of a column whose (allele, label) bytes are fetched one item ahead

template <class Idx>
__global__ void __launch_bounds__(kTmaThreads, 1)
pack_sm2_anc_tma_kernel(const __grid_constant__ CUtensorMap map_gt, const __grid_constant__ CUtensorMap map_anc,
                        const PackArgs a, const TmaItems w) {
    extern __shared__ __align__(1024) uint8_t s_stage[];  // ring of stages, then the hand-over buffers, then the barriers
    uint32_t* s_planes = reinterpret_cast<uint32_t*>(s_stage + (size_t)kTmaStages * kTmaStageBytes);
    uint64_t* s_full = reinterpret_cast<uint64_t*>(s_planes + kTmaGroups * kPlaneBufWords);
    uint64_t* s_empty = s_full + kTmaStages;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        if (smem_addr(s_stage) & 127u) __trap();  // tensor copies need 128-byte aligned destinations
        for (int s = 0; s < kTmaStages; ++s) {
            mbar_init(&s_full[s], 1);
            mbar_init(&s_empty[s], kTmaGroupWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == 0) {
        // ---- producer: one lane keeps the ring full
        if (lane == 0) {
            uint64_t policy;  // see launch_pack_sm2_anc_tma
            if (w.stream_hint) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
            else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(policy));
            uint32_t s = 0, round = 0;
            for (uint32_t k = 0;; ++k) {
                const TmaItem it = tma_item<Idx>(w, k);
                if (!it.valid) break;
                if (round > 0) mbar_wait(&s_empty[s], (round - 1) & 1u);
                uint8_t* dst = s_stage + (size_t)s * kTmaStageBytes;
                const int c0 = (int)(it.cb * kTmaBoxBytes), c1 = (int)(it.tile * kTileSamples + it.slab * kTmaRows);
                mbar_expect_tx(&s_full[s], kTmaStageBytes);
                tma_load_2d(dst, &map_gt, c0, c1, &s_full[s], policy);
                tma_load_2d(dst + kTmaMatBytes, &map_anc, c0, c1, &s_full[s], policy);
                if (++s == kTmaStages) {
                    s = 0;
                    ++round;
                }
            }
        }
        return;
    }

    // ---- consumers
    const int group = (warp - 1) / kTmaGroupWarps, gw = (warp - 1) % kTmaGroupWarps;  // gw = my word of the slab (phase A)
    const int gtid = gw * 32 + lane;
    const int bcol = gtid >> 1, bhalf = gtid & 1;  // phase B: my column of the block, my half of the slab's words
    uint32_t* planes = s_planes + group * kPlaneBufWords;
    const int bar_reads = 1 + 2 * group, bar_writes = 2 + 2 * group;
    uint32_t fl = 0u;

    // software pipeline over my group's items: the key range of phase B's column is fetched one item ahead, the
    // (allele, label) bytes of its first keys at the end of the previous item -- no load of phase B waits
    auto key_range = [&](const TmaItem& it) {
        const int64_t col = (int64_t)it.cb * kTmaCols + bcol;
        int2 kr = make_int2(0, 0);
        if (it.valid && col < a.n_cols) kr = __ldg(reinterpret_cast<const int2*>(a.col_key_range) + col);
        return kr;
    };
    auto key_info = [&](int key) {
        return (uint32_t)__ldg(a.key_allele + key) | ((uint32_t)__ldg(a.key_label + key) << 8);
    };
    // the prefetched bytes stay in the registers the loads fill (no arithmetic on them before phase B: the first
    // instruction that touches a loaded byte waits for it)
    auto fetch_info = [&](const int2& r, uint32_t (&al)[kInfoAhead], uint32_t (&lb)[kInfoAhead]) {
#pragma unroll
        for (int t = 0; t < kInfoAhead; ++t) {
            al[t] = lb[t] = 0u;
            if (t < r.y) {
                al[t] = __ldg(a.key_allele + r.x + t);
                lb[t] = __ldg(a.key_label + r.x + t);
            }
        }
    };
    uint32_t k = group;
    TmaItem it = tma_item<Idx>(w, k);
    int2 kr = key_range(it);
    uint32_t info_a[kInfoAhead], info_l[kInfoAhead];
    fetch_info(kr, info_a, info_l);
    uint32_t s = group % kTmaStages, round = group / kTmaStages;

    while (it.valid) {
        const TmaItem it_next = tma_item<Idx>(w, k + kTmaGroups);
        const int2 kr_next = key_range(it_next);

        mbar_wait(&s_full[s], round & 1u);
        const uint8_t* g_rows = s_stage + (size_t)s * kTmaStageBytes + (gw * 32) * kTmaBoxBytes + 4 * lane;
        const uint8_t* l_rows = g_rows + kTmaMatBytes;

        // ---- phase A: 32 rows x 4 bytes of both matrices -> 4 plane words per byte column
        uint32_t blk[4][4];  // [8-row block][plane]: byte = byte column, bit = row of the block
        uint32_t d = 0u;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint32_t x[8], c[8];
#pragma unroll
            for (int rr = 0; rr < 8; ++rr) {
                x[rr] = *reinterpret_cast<const uint32_t*>(g_rows + (8 * q + rr) * kTmaBoxBytes);
                c[rr] = *reinterpret_cast<const uint32_t*>(l_rows + (8 * q + rr) * kTmaBoxBytes);
            }
            uint32_t z[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                // nibble of a row: bit 0 = GT bit 0, bits 1..3 = label bits 0..2; rows i (low) and i + 4 (high)
                const uint32_t lo = bitsel(0x01010101u, x[i], c[i] << 1);
                const uint32_t hi = bitsel(0x01010101u, x[i + 4], c[i + 4] << 1);
                z[i] = bitsel(0x0f0f0f0fu, lo, hi << 4);
                d |= ((x[i] | x[i + 4]) & 0xfefefefeu) | ((c[i] | c[i + 4]) & 0xf8f8f8f8u);
            }
            nibble_transpose4(z);
#pragma unroll
            for (int b = 0; b < 4; ++b) blk[q][b] = z[b];
        }
        uint32_t P[5][4];  // [plane][byte column]: bit = row of the word
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const uint32_t G[4] = {blk[0][b], blk[1][b], blk[2][b], blk[3][b]};
            byte_transpose4(G, P[b]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) P[4][i] = 0u;
        if (d) {
            // some byte of mine is a missing / multi-allelic call or a label >= 8: the invalid plane, from a
            // second pass over my rows in the stage
            const int64_t col0 = (int64_t)it.cb * kTmaCols + 2 * lane;
            uint32_t refmask = 0u;
            if (a.flags) {
                if (col0 < a.n_cols && __ldg(a.col_key_range + 2 * col0 + 1) > 0) refmask |= 0x0000ffffu;
                if (col0 + 1 < a.n_cols && __ldg(a.col_key_range + 2 * col0 + 3) > 0) refmask |= 0xffff0000u;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint32_t z = 0u;
#pragma unroll 1
                for (int rr = 0; rr < 8; ++rr) {  // row rr enters at bit 7, ends at bit rr
                    const uint32_t x = *reinterpret_cast<const uint32_t*>(g_rows + (8 * q + rr) * kTmaBoxBytes);
                    const uint32_t c = *reinterpret_cast<const uint32_t*>(l_rows + (8 * q + rr) * kTmaBoxBytes);
                    z = (z >> 1) | nz_bytes_msb((x & 0xfefefefeu) | (c & 0xf8f8f8f8u));
                    fl |= qc_flags4(x & refmask);
                }
                blk[q][0] = z;
            }
            const uint32_t G[4] = {blk[0][0], blk[1][0], blk[2][0], blk[3][0]};
            byte_transpose4(G, P[4]);
        }
        // the stage has been read: hand it back
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[s]);

        // ---- hand-over: planes[plane][word][byte column]
        group_barrier(bar_reads);  // phase B of the previous item has read the buffer
#pragma unroll
        for (int p = 0; p < 5; ++p)
            *reinterpret_cast<uint4*>(planes + (p * 4 + gw) * kPlanePitch + 4 * lane) = make_uint4(P[p][0], P[p][1], P[p][2], P[p][3]);
        group_barrier(bar_writes);

        // ---- phase B: every key of my column, words 2 * bhalf, 2 * bhalf + 1 of the slab, both strands
        if (kr.y > 0) {
            uint2 Q[5][2];  // [plane][word of my half]: {strand 0, strand 1}
#pragma unroll
            for (int p = 0; p < 5; ++p)
#pragma unroll
                for (int ww = 0; ww < 2; ++ww)
                    Q[p][ww] = *reinterpret_cast<const uint2*>(planes + (p * 4 + 2 * bhalf + ww) * kPlanePitch + 2 * bcol);
            const int word0 = (int)it.slab * 4 + 2 * bhalf;
            const uint32_t vm0 = valid_mask(a.n, it.tile, word0), vm1 = valid_mask(a.n, it.tile, word0 + 1);
            const uint32_t ok00 = ~Q[4][0].x & vm0, ok01 = ~Q[4][0].y & vm0, ok10 = ~Q[4][1].x & vm1, ok11 = ~Q[4][1].y & vm1;
            uint4* rec = reinterpret_cast<uint4*>(record(a, it.tile, (uint32_t)kr.x) + word0);
            auto emit = [&](uint32_t inf, int t) {
                const uint32_t mA = (inf & 1u) ? 0u : 0xffffffffu, m0 = (inf & 0x100u) ? 0u : 0xffffffffu;
                const uint32_t m1 = (inf & 0x200u) ? 0u : 0xffffffffu, m2 = (inf & 0x400u) ? 0u : 0xffffffffu;
                uint4 o;
                o.x = (Q[0][0].x ^ mA) & (Q[1][0].x ^ m0) & (Q[2][0].x ^ m1) & (Q[3][0].x ^ m2) & ok00;
                o.y = (Q[0][0].y ^ mA) & (Q[1][0].y ^ m0) & (Q[2][0].y ^ m1) & (Q[3][0].y ^ m2) & ok01;
                o.z = (Q[0][1].x ^ mA) & (Q[1][1].x ^ m0) & (Q[2][1].x ^ m1) & (Q[3][1].x ^ m2) & ok10;
                o.w = (Q[0][1].y ^ mA) & (Q[1][1].y ^ m0) & (Q[2][1].y ^ m1) & (Q[3][1].y ^ m2) & ok11;
                rec[(size_t)t * (kRecordWords / 4)] = o;  // records of consecutive keys are 256 bytes apart
            };
#pragma unroll
            for (int t = 0; t < kInfoAhead; ++t)
                if (t < kr.y) emit(info_a[t] | (info_l[t] << 8), t);
#pragma unroll 1
            for (int t = kInfoAhead; t < kr.y; ++t) emit(key_info(kr.x + t), t);
        }

        // next item of my group
        fetch_info(kr_next, info_a, info_l);
        it = it_next;
        kr = kr_next;
        k += kTmaGroups;
        s += kTmaGroups;
        if (s >= kTmaStages) {
            s -= kTmaStages;
            ++round;
        }
    }
    if (a.flags && fl) {
        if (fl & ~*reinterpret_cast<volatile uint32_t*>(a.flags)) atomicOr(a.flags, fl);
    }
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// cuTensorMapEncodeTiled through the runtime (the library does not link libcuda: it has to load on machines
// without a driver, tests/test_abi.py)
static EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

bool pack_anc_tma_available() { return encode_tiled() != nullptr; }

static int make_map(CUtensorMap* m, const uint8_t* base, int64_t row_pitch, int64_t n_rows, int64_t row_bytes) {
    const cuuint64_t dims[2] = {(cuuint64_t)row_bytes, (cuuint64_t)n_rows};
    const cuuint64_t strides[1] = {(cuuint64_t)row_pitch};
    const cuuint32_t box[2] = {kTmaBoxBytes, kTmaRows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode_tiled()(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint8_t*>(base), dims, strides, box, estr,
                                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) for a %lld x %lld byte matrix, pitch %lld", (int)r, (long long)n_rows,
                  (long long)row_bytes, (long long)row_pitch);
        return HT_ERR_UNSUPPORTED;
    }
    return HT_OK;
}

// Preconditions (checked by the caller, ht_pack_u8): strand stride 1, variant stride 2, rows and bases 16-byte
// aligned, col_key_range present.
int launch_pack_sm2_anc_tma(const PackArgs& a, cudaStream_t st) {
    if (!encode_tiled()) {
        set_error("cuTensorMapEncodeTiled is not available from this driver");
        return HT_ERR_UNSUPPORTED;
    }
    const int64_t row_bytes = 2 * a.n_cols;
    if (a.n >= ((int64_t)1 << 31) || row_bytes >= ((int64_t)1 << 31)) {
        set_error("TMA pack: matrix too large for 32-bit tensor coordinates");
        return HT_ERR_UNSUPPORTED;
    }
    CUtensorMap map_gt, map_anc;
    if (int rc = make_map(&map_gt, a.gt, a.gs, a.n, row_bytes)) return rc;
    if (int rc = make_map(&map_anc, a.anc, a.as, a.n, row_bytes)) return rc;

    TmaItems w;
    w.n_tiles = ht_num_tiles(a.n);
    w.n_cb = (uint32_t)((row_bytes + kTmaBoxBytes - 1) / kTmaBoxBytes);
    w.total = w.n_tiles * kTmaSlabsPerTile * (int64_t)w.n_cb;
    // item order: HT_TMA_GROUP column blocks per group (default 64: the 8 sectors of a key's record are written
    // within ~500 items of each other and the CTAs in flight read one or two slabs; 0 = the whole row, i.e.
    // slab-major inside a tile: 15.3 against 14.5 ms), HT_TMA_CHUNK consecutive items per CTA (default 1)
    const char* g_env = getenv("HT_TMA_GROUP");
    const char* c_env = getenv("HT_TMA_CHUNK");
    const int64_t g = g_env ? atoll(g_env) : 64, c = c_env ? atoll(c_env) : 1;
    w.group = (uint32_t)((g > 0 && g < w.n_cb) ? g : w.n_cb);
    // L2 evict_first on the matrix loads is SLOWER: 15.2 against 14.5 ms on the ancestry config (rows of whole
    // 128-byte lines), 30 against 14.5 ms with 100,000-byte rows, where neighbouring boxes share the lines at
    // either end (profiles/r2/pack_anc_tma.json).  HT_TMA_HINT=1 switches it on for measurements.
    const char* h_env = getenv("HT_TMA_HINT");
    w.stream_hint = h_env ? (uint32_t)(atoi(h_env) != 0) : 0u;
    w.chunk_log2 = 0;
    while ((2ll << w.chunk_log2) <= c && w.chunk_log2 < 6) ++w.chunk_log2;

    int dev = 0, sms = 0;
    HT_CUDA_TRY(cudaGetDevice(&dev));
    HT_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int64_t want = (w.total + (1ll << w.chunk_log2) - 1) >> w.chunk_log2;
    const unsigned grid = (unsigned)(want < sms ? want : sms);
    if (grid == 0) return HT_OK;
    // 32-bit item arithmetic unless the launch has 2^32 items or more (the index of the first item past the end,
    // computed by every CTA, must fit too)
    const char* wide_env = getenv("HT_TMA_WIDE");  // tests: the 64-bit instance on small inputs
    const bool wide = w.total + ((int64_t)4 * (grid + 1) << w.chunk_log2) >= ((int64_t)1 << 32) || (wide_env && atoi(wide_env));
    auto kernel = wide ? pack_sm2_anc_tma_kernel<uint64_t> : pack_sm2_anc_tma_kernel<uint32_t>;
    HT_CUDA_TRY(cudaFuncSetAttribute((const void*)kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTmaSmem));
    kernel<<<grid, kTmaThreads, kTmaSmem, st>>>(map_gt, map_anc, a, w);
    HT_CUDA_TRY(cudaGetLastError());
    return HT_OK;
}

}  // namespace ht
