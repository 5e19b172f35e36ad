// K2, half-tile form (sm_100a): the transform kernel ht_transform_u8 launches whenever the planner's
// key quads are available and the result rows take 16-byte stores.
//
// Same result as transform_u8_kernel (ht_transform.cu) -- the reference's
//     hap_gts.data[:, i] = np.all(equality_arr[:, idxs[i]], axis=1)
// (haptools/data/haplotypes.py:1411-1412; haptools/transform.py:144-148) -- with a different work
// decomposition, chosen from what ncu and the key-space experiment showed about the tile form
// (DESIGN.md section 5): its AND phase waits for plane records that miss the near L2 partition, and
// how many do depends on the bytes of records the resident CTAs share (25.6 MB per 1024-sample tile
// at 100 k keys).
//
//   CTA  = one HALF tile (512 samples = one 128-byte cache line of every 256-byte record) x 128
//          haplotypes; blockIdx is half-tile-major, so the CTAs in flight share 12.8 MB of lines.
//   warp = 16 haplotypes; lane = (word w = lane & 15, half = lane >> 4).
//   1. AND: one load instruction reads TWO records' lines: lanes 0-15 the 16 words of one key,
//      lanes 16-31 those of the next key of the same haplotype (a quad of 4 keys = 2 loads).  When a
//      haplotype's last quad is consumed the two halves are combined with one shuffle per strand.
//      Lanes of half g keep the results of haplotypes 8g..8g+7: 16 words (8 haplotypes x 2 strands),
//      each over 32 samples.
//   2. 16x32 bit transpose in registers (2 PRMT + 2 shift/select stages).  The order of the stages
//      is chosen so that output word W_m holds TWO samples (s and s+4), interleaved such that the
//      four 0/1 bytes of result word k are (W >> k) & 0x01010101 (sample s) and
//      (W >> (4 + k)) & 0x01010101 (sample s + 4): bits -> bytes costs 2 instructions per 4 bytes.
//   3. Exchange through shared memory (pitch 514 words per warp: conflict-free both ways): 16
//      consecutive threads cover the 256 contiguous result bytes of one sample row; every store
//      instruction writes 2 x 256 contiguous bytes (st.global.cs.v4).
//
// Algorithmic bytes are those of the tile form: planes read once, n*H*2 result bytes written.
#include <atomic>
#include <cstdlib>

#include "ht_transform.cuh"

namespace ht {

constexpr int kHalfSamples = kTileSamples / 2;  // 512
constexpr int kXPitch = 16 * 32 + 2;            // exchange words per warp (+2: rows of the 8 warps in distinct banks)

constexpr uint32_t kHQuadLast = 1u, kHQuadSkip = 2u, kHOffMask = 0xffffff00u;

#ifndef HT_HALF_BQ
#define HT_HALF_BQ 4
#endif
#ifndef HT_HALF_MIN_CTAS
#define HT_HALF_MIN_CTAS 4
#endif
#ifndef HT_STAGE_MIN_CTAS
#define HT_STAGE_MIN_CTAS 3
#endif
#ifndef HT_HALF_PER_CTA
#define HT_HALF_PER_CTA 1
#endif
constexpr int kHalfPerCta = HT_HALF_PER_CTA;  // consecutive half tiles one CTA handles
constexpr int kHBQ = HT_HALF_BQ;  // quads per batch: 2 * kHBQ loads (of 2 x 128 B) in flight per batch

__device__ __forceinline__ uint2 ld_line(const char* lane_base, uint32_t off, uint64_t pol) {
    // Lines are re-read by every haplotype that uses the key while 8x as many result bytes stream
    // through L2: evict_last.  The policy arrives as a kernel parameter (made once per device by
    // make_policy_kernel) so that it lives in a uniform register; a createpolicy inside this kernel
    // costs two R2UR per load.  Address = 64-bit lane base + 32-bit record offset in ONE instruction
    // (IMAD.WIDE.U32) instead of an add-with-carry pair.
    uint64_t addr;
    asm("mad.wide.u32 %0, %1, 1, %2;" : "=l"(addr) : "r"(off), "l"(lane_base));
    uint2 v;
    asm volatile("ld.global.nc.L2::cache_hint.v2.u32 {%0, %1}, [%2], %3;" : "=r"(v.x), "=r"(v.y) : "l"(addr), "l"(pol));
    return v;
}

// staged form: the line sits in shared memory at the (transformed) offset; `lane_base` is then a shared-window
// address (stage buffer + 8 * word)
__device__ __forceinline__ uint2 ld_line_smem(uint32_t lane_base_s, uint32_t off) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(lane_base_s + off));
    return v;
}

struct HBatch {
    uint2 v[2 * kHBQ];
    uint32_t f[kHBQ];
};

// One haplotype done: combine the two halves' partial ANDs, the half that keeps this haplotype parks it.
__device__ __forceinline__ void hemit(uint32_t flags, uint32_t& r0, uint32_t& r1, uint2*& park, uint32_t& emitted,
                                      int half) {
    if (flags & kHQuadSkip) r0 = r1 = 0xffffffffu;  // haplotype without alleles: its one quad read record 0
    r0 &= __shfl_xor_sync(0xffffffffu, r0, 16);
    r1 &= __shfl_xor_sync(0xffffffffu, r1, 16);
    if ((int)(emitted >> 3) == half) {
        *park = make_uint2(r0, r1);
        park += 32;
    }
    ++emitted;
    r0 = r1 = 0xffffffffu;
}

// Staged form: `qp` points at the lane's pair (entries 0-1 or 2-3, flags replicated into entry 2) of
// the warp's first quad in shared memory; the warp's stream was padded to a multiple of kHBQ quads
// (copies of its last quad without flags), so there is no clamping and no tail.
template <bool kSmem>
__device__ __forceinline__ void hload_batch_s(const char* __restrict__ lane_base, uint32_t lane_base_s,
                                              const uint2* __restrict__ qp, uint64_t pol, HBatch& b) {
    uint2 k[kHBQ];
#pragma unroll
    for (int i = 0; i < kHBQ; ++i) k[i] = qp[2 * i];
#pragma unroll
    for (int i = 0; i < kHBQ; ++i) {
        b.f[i] = k[i].x;
        if (kSmem) {  // offsets are line * 128 here: the low 7 bits carry the flags
            b.v[2 * i + 0] = ld_line_smem(lane_base_s, k[i].x & ~127u);
            b.v[2 * i + 1] = ld_line_smem(lane_base_s, k[i].y);
        } else {
            b.v[2 * i + 0] = ld_line(lane_base, k[i].x & kHOffMask, pol);
            b.v[2 * i + 1] = ld_line(lane_base, k[i].y, pol);
        }
    }
}

__device__ __forceinline__ void hconsume_batch(const HBatch& b, uint32_t& r0, uint32_t& r1, uint2*& park,
                                               uint32_t& emitted, int half) {
#pragma unroll
    for (int i = 0; i < kHBQ; ++i) {
        r0 &= b.v[2 * i].x & b.v[2 * i + 1].x;
        r1 &= b.v[2 * i].y & b.v[2 * i + 1].y;
        if (b.f[i] & kHQuadLast) hemit(b.f[i], r0, r1, park, emitted, half);  // warp-uniform
    }
}

template <bool kSmem>
__device__ __forceinline__ void hand_phase_staged(const char* __restrict__ lane_base, uint32_t lane_base_s,
                                                  const uint2* __restrict__ qp, uint32_t n_batches, int half,
                                                  uint64_t pol, uint2* park) {
    if (n_batches == 0) return;
    uint32_t r0 = 0xffffffffu, r1 = 0xffffffffu, emitted = 0;
    HBatch A, B;
    hload_batch_s<kSmem>(lane_base, lane_base_s, qp, pol, A);
#pragma unroll 1
    while (true) {
        if (n_batches > 1) hload_batch_s<kSmem>(lane_base, lane_base_s, qp + 2 * kHBQ, pol, B);
        hconsume_batch(A, r0, r1, park, emitted, half);
        if (n_batches <= 1) break;
        if (n_batches > 2) hload_batch_s<kSmem>(lane_base, lane_base_s, qp + 4 * kHBQ, pol, A);
        hconsume_batch(B, r0, r1, park, emitted, half);
        if (n_batches <= 2) break;
        n_batches -= 2;
        qp += 4 * kHBQ;
    }
}

// Unstaged form (a block whose quads do not fit the staging buffer): quads straight from the planner's
// table in global memory, one quad at a time.
__device__ __forceinline__ void hand_phase_global(const char* __restrict__ lane_base, const uint4* __restrict__ quads,
                                                  uint32_t q, uint32_t q_end, int half, uint64_t pol, uint2* park) {
    uint32_t r0 = 0xffffffffu, r1 = 0xffffffffu, emitted = 0;
#pragma unroll 1
    for (; q < q_end; ++q) {
        const uint4 t = __ldg(quads + q);
        const uint2 v0 = ld_line(lane_base, (half ? t.z : t.x) & kHOffMask, pol);
        const uint2 v1 = ld_line(lane_base, half ? t.w : t.y, pol);
        r0 &= v0.x & v1.x;
        r1 &= v0.y & v1.y;
        if (t.x & kHQuadLast) hemit(t.x, r0, r1, park, emitted, half);
    }
}

// 16 words (slot i = result byte 2*hap + strand of a 16-byte piece, bits over 32 samples) -> 16 words
// (slot m, holding samples j(m) and j(m) + 4 interleaved as described at the top of the file).
// Every stage swaps one bit of the slot index with one bit of the bit position:
//   position 4 <-> slot bit 1, position 3 <-> slot bit 0, position 1 <-> slot bit 3, position 0 <-> slot bit 2
// so that afterwards slot m = (j1 j0 j4 j3), position = (i1 i0 j2 i3 i2).
__device__ __forceinline__ void transpose16x32_dev(uint32_t (&a)[16]) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        if (k & 2) continue;
        const uint32_t lo = a[k], hi = a[k + 2];
        a[k] = __byte_perm(lo, hi, 0x5410);
        a[k + 2] = __byte_perm(lo, hi, 0x7632);
    }
#pragma unroll
    for (int k = 0; k < 16; k += 2) {
        const uint32_t lo = a[k], hi = a[k + 1];
        a[k] = __byte_perm(lo, hi, 0x6240);
        a[k + 1] = __byte_perm(lo, hi, 0x7351);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {  // shift 2, partner k + 8
        const uint32_t x = a[k], y = a[k + 8];
        a[k] = (x & 0x33333333u) | ((y << 2) & 0xccccccccu);
        a[k + 8] = ((x >> 2) & 0x33333333u) | (y & 0xccccccccu);
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) {  // shift 1, partner k + 4
        if (k & 4) continue;
        const uint32_t x = a[k], y = a[k + 4];
        a[k] = (x & 0x55555555u) | ((y << 1) & 0xaaaaaaaau);
        a[k + 4] = ((x >> 1) & 0x55555555u) | (y & 0xaaaaaaaau);
    }
}

// first of the two samples (the other is + 4) of a word held by slot m after transpose16x32_dev
__device__ __forceinline__ int slot_sample(int m) {
    return 8 * (m & 3) + (m >> 2);  // = 16 m1 + 8 m0 + 2 m3 + m2
}

__global__ void make_policy_kernel(uint64_t* out) {
    uint64_t pol;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    *out = pol;
}

// Step 3b of the kernel: every thread expands words of the exchange buffer to bytes and stores its 16-byte pieces
// of the (half tile x 128 haplotypes) result band.
__device__ __forceinline__ void hstore_phase(const uint32_t* __restrict__ s_x, int tid, int n_blk_haps, int ht,
                                             int64_t h0, int64_t n_samples, uint8_t* __restrict__ out, int64_t out_s_samp) {
    const int cb = tid & 15;  // 16-byte piece of the 256-byte row band: haplotypes h0 + 8 cb .. + 7
    const int g = tid >> 4;   // word of the half tile: samples 32 g .. 32 g + 31
    const int n_my_haps = min(8, n_blk_haps - 8 * cb);
    if (n_my_haps <= 0) return;
    const int64_t s_base = (int64_t)ht * kHalfSamples;
    const int rows = (int)min((int64_t)kHalfSamples, n_samples - s_base);
    const uint32_t* src = s_x + (cb >> 1) * kXPitch + (cb & 1) * 16 + g;
    uint8_t* p0 = out + 2 * (h0 + 8 * cb) + (s_base + 32 * g) * out_s_samp;

    if (n_my_haps == 8 && rows == kHalfSamples) {
#pragma unroll 4
        for (int m = 0; m < 16; ++m) {
            const uint32_t W = src[m * 32];
            uint8_t* p = p0 + (int64_t)slot_sample(m) * out_s_samp;
            uint4 lo, hi;
            lo.x = W & 0x01010101u;
            lo.y = (W >> 1) & 0x01010101u;
            lo.z = (W >> 2) & 0x01010101u;
            lo.w = (W >> 3) & 0x01010101u;
            hi.x = (W >> 4) & 0x01010101u;
            hi.y = (W >> 5) & 0x01010101u;
            hi.z = (W >> 6) & 0x01010101u;
            hi.w = (W >> 7) & 0x01010101u;
            st_stream_v4(p, lo);
            st_stream_v4(p + 4 * out_s_samp, hi);
        }
        return;
    }
    // ragged edge: last rows of the matrix and / or a last group of fewer than 8 haplotypes
#pragma unroll 1
    for (int m = 0; m < 16; ++m) {
        const uint32_t W = src[m * 32];
#pragma unroll 1
        for (int u = 0; u < 2; ++u) {
            const int sl = 32 * g + slot_sample(m) + 4 * u;
            if (sl >= rows) continue;
            uint8_t* p = p0 + (int64_t)(slot_sample(m) + 4 * u) * out_s_samp;
            const uint32_t V = W >> (4 * u);
            if (n_my_haps == 8) {  // last rows of the matrix, full group: still one 16-byte store per row
                st_stream_v4(p, make_uint4(V & 0x01010101u, (V >> 1) & 0x01010101u, (V >> 2) & 0x01010101u,
                                           (V >> 3) & 0x01010101u));
                continue;
            }
            for (int q = 0; q < n_my_haps; ++q) {
                // bytes 2q, 2q+1 of the piece: byte i sits at bit 8 * (i & 3) + (i >> 2)
                const int i0 = 2 * q, i1 = 2 * q + 1;
                const uint32_t b0 = (V >> (8 * (i0 & 3) + (i0 >> 2))) & 1u;
                const uint32_t b1 = (V >> (8 * (i1 & 3) + (i1 >> 2))) & 1u;
                *reinterpret_cast<uint16_t*>(p + 2 * q) = (uint16_t)(b0 | (b1 << 8));
            }
        }
    }
}

// kStage: blocks whose keys fall into a run of at most `stage_lines` consecutive keys (blk_key_lo / blk_key_cnt,
// planner: TransformPlan.block_spans) copy that run of 128-byte lines into shared memory ONCE with cp.async and
// AND out of shared memory; what crosses L2 -> SM per block is then the run (219 lines on average for the
// position-sorted UKB-scale plan) instead of one line per haplotype allele (1,406).  Other blocks of the same
// launch take the L2 path below.  Position-sorted haplotype sets (what `haptools index` produces,
// haptools/index.py:31-94) are where blocks are key-local.
template <bool kStage, int kHpc>
__global__ void __launch_bounds__(kThreads, kStage ? HT_STAGE_MIN_CTAS : HT_HALF_MIN_CTAS)
transform_u8_half_kernel(const uint32_t* __restrict__ planes, int64_t n_samples, int64_t n_keys,
                         const uint32_t* __restrict__ quad_off16, const uint4* __restrict__ quads,
                         const uint32_t* __restrict__ blk_key_lo, const uint32_t* __restrict__ blk_key_cnt,
                         uint32_t stage_lines, int64_t n_haps, uint32_t n_hap_blocks, int64_t group0,
                         uint8_t* __restrict__ out, int64_t out_s_samp, unsigned long long* __restrict__ counts,
                         uint64_t pol) {
    __shared__ __align__(16) uint32_t s_x[kWarps * kXPitch];  // 16.4 KB: parked results, then the exchange
    __shared__ __align__(16) uint32_t s_keys[kStageKeys];      // 12 KB, as uint4 quads
    extern __shared__ __align__(128) uint8_t s_stage[];        // kStage: stage_lines * 128 B

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int w = lane & 15, half = lane >> 4;
    // the CTA's `hpc` consecutive half tiles (samples [512 ht, 512 ht + 512)): the key lists are staged once for
    // all of them, and the stores of one half drain while the next one's records are being ANDed
    // (32-bit: 2^31 half tiles are 10^12 samples; n_haps < 2^31 * 128 is checked by the launcher's grid limit)
    const int ht_first = (int)((group0 + (int64_t)(blockIdx.x / n_hap_blocks)) * kHpc);
    const int n_half = (int)((n_samples + kHalfSamples - 1) / kHalfSamples);
    const int64_t h0 = (int64_t)(blockIdx.x % n_hap_blocks) * kHapBlock;
    const int n_blk_haps = (int)min((int64_t)kHapBlock, n_haps - h0);
    const int hl_begin = warp * kHapsPerWarp;
    const int hl_end = min(hl_begin + kHapsPerWarp, n_blk_haps);
    uint32_t* my_row = s_x + warp * kXPitch;

    // ---- the quads of my 16 haplotypes ---------------------------------------------------------
    // staged at quad index q_begin + kHBQ * warp of s_keys, padded to a multiple of kHBQ quads
    const int64_t g0 = (h0 >> 4), n16 = (n_haps + kHapsPerWarp - 1) >> 4;
    const uint32_t blk_q0 = __ldg(quad_off16 + g0);
    const uint32_t blk_q1 = __ldg(quad_off16 + min(g0 + kWarps, n16));
    const bool staged = (blk_q1 - blk_q0 + kHBQ * kWarps) * 4u <= (uint32_t)kStageKeys;  // block-uniform
    uint32_t q_begin = 0, q_end = 0;
    if (hl_begin < hl_end) {
        q_begin = __ldg(quad_off16 + g0 + warp) - blk_q0;
        q_end = __ldg(quad_off16 + g0 + warp + 1) - blk_q0;
    }
    const uint32_t n_batches = (q_end - q_begin + kHBQ - 1) / kHBQ;
    uint4* my_quads = reinterpret_cast<uint4*>(s_keys) + q_begin + kHBQ * warp;
    // ---- the block's run of lines -> shared memory (asynchronously; waited for after the quads are staged) ------
    uint32_t run_lo = 0, run_cnt = 0;
    if (kStage && staged) {
        const uint32_t hb = blockIdx.x % n_hap_blocks;
        run_cnt = __ldg(blk_key_cnt + hb);
        if (run_cnt > stage_lines) run_cnt = 0;  // block-uniform: 0 = this block reads its lines from L2
        if (run_cnt) run_lo = __ldg(blk_key_lo + hb);
    }
    auto stage_run = [&](int ht) {  // the run's line of half tile `ht` -> shared memory, asynchronously
        const char* src = reinterpret_cast<const char*>(planes) +
                          ((size_t)(ht >> 1) * (size_t)n_keys + run_lo) * (kRecordWords * 4) + (ht & 1) * 128 + 16 * (tid & 7);
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(s_stage) + 16 * (tid & 7);
        for (uint32_t line = tid >> 3; line < run_cnt; line += kThreads / 8)
            asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(dst + line * 128),
                         "l"(src + (size_t)line * (kRecordWords * 4)), "l"(pol)
                         : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (kStage && run_cnt) stage_run(ht_first);
    if (staged && q_begin < q_end) {
        const uint4* srcq = quads + blk_q0 + q_begin;
        const uint32_t nq = q_end - q_begin;
        const uint32_t lo256 = run_lo << 8;
        for (uint32_t i = lane; i < n_batches * kHBQ; i += 32) {
            uint4 t = __ldg(srcq + min(i, nq - 1));
            uint32_t f = t.x & ~kHOffMask;
            const bool skip = (f & kHQuadSkip) != 0u;  // the one quad of a haplotype without alleles
            if (i >= nq) f = 0u;                       // padding: the warp's last quad again, without flags
            t.x &= kHOffMask;
            if (kStage && run_cnt) {
                // record offsets (key * 256) -> offsets into the staged run (line * 128: bit 7 now belongs to the
                // offset, the flags keep bits 0-1).  The quad of a haplotype without alleles points at record 0,
                // which may lie outside the run (and so does a padding copy of it): any staged line will do
                if (skip) {
                    t = make_uint4(0u, 0u, 0u, 0u);
                } else {
                    t.x = (t.x - lo256) >> 1;
                    t.y = (t.y - lo256) >> 1;
                    t.z = (t.z - lo256) >> 1;
                    t.w = (t.w - lo256) >> 1;
                }
            }
            t.x |= f;
            t.z |= f;  // both halves of the warp see the flags in their own pair
            my_quads[i] = t;
        }
    }
    __syncwarp();

#pragma unroll 1
    for (int it = 0; it < kHpc; ++it) {
        const int ht = ht_first + it;
        if (ht >= n_half) break;  // CTA-uniform
        if (kStage && run_cnt) {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncthreads();  // everybody's pieces of this half tile's run have landed
        }

        // ---- 1. AND ------------------------------------------------------------------------------
        // results are parked in the warp's own exchange row: lane (w, g) keeps haplotypes 8g..8g+7
        uint2* park = reinterpret_cast<uint2*>(my_row) + lane;
        if (q_begin < q_end) {
            const char* lane_base = reinterpret_cast<const char*>(planes) +
                                    ((size_t)(ht >> 1) * (size_t)n_keys) * (kRecordWords * 4) + (ht & 1) * 128 + 8 * w;
            if (kStage && run_cnt)
                hand_phase_staged<true>(lane_base, (uint32_t)__cvta_generic_to_shared(s_stage) + 8 * w,
                                        reinterpret_cast<const uint2*>(my_quads) + half, n_batches, half, pol, park);
            else if (staged)
                hand_phase_staged<false>(lane_base, 0u, reinterpret_cast<const uint2*>(my_quads) + half, n_batches, half, pol, park);
            else
                hand_phase_global(lane_base, quads + blk_q0, q_begin, q_end, half, pol, park);
        }
        __syncwarp();
        uint32_t a[16];  // a[2*j + c] = haplotype 8*half + j of this warp, strand c; bits over my 32 samples
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint2 r = park[j * 32];
            a[2 * j] = r.x;
            a[2 * j + 1] = r.y;
        }
        if (counts) {
            // allele counts for check_maf (haptools/data/genotypes.py:559-625), fused: popcount of my word
            // of every result, one reduction per half warp and one atomic per haplotype and half tile
            const int64_t rem = n_samples - (int64_t)ht * kHalfSamples - 32 * (int64_t)w;
            const uint32_t vm = rem >= 32 ? 0xffffffffu : (rem <= 0 ? 0u : ((1u << (int)rem) - 1u));
            const uint32_t members = half ? 0xffff0000u : 0x0000ffffu;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint32_t c = __reduce_add_sync(members, __popc(a[2 * j] & vm) + __popc(a[2 * j + 1] & vm));
                const int hl = hl_begin + 8 * half + j;
                if (w == 0 && hl < hl_end && c) atomicAdd(counts + h0 + hl, (unsigned long long)c);
            }
        }
        __syncwarp();  // everybody has read its parked results: the row is free for the exchange

        // ---- 2. bits over samples -> interleaved bits over (haplotype, strand) --------------------
        transpose16x32_dev(a);

        // ---- 3. exchange + expand + store ------------------------------------------------------------
        {
            uint32_t* dst = my_row + half * 16 + w;
#pragma unroll
            for (int m = 0; m < 16; ++m) dst[m * 32] = a[m];
        }
        __syncthreads();  // the exchange is complete, and every warp is done with this half tile's records
        const bool more = it + 1 < kHpc && ht + 1 < n_half;  // CTA-uniform
        if (kStage && run_cnt && more) stage_run(ht + 1);  // the next run flies while this half is stored
        hstore_phase(s_x, tid, n_blk_haps, ht, h0, n_samples, out, out_s_samp);
        if (more) __syncthreads();  // the exchange rows are free for the next half's parked results
    }
}

// evict_last policy word of the current device (made once per device and process; 0 = not made yet:
// createpolicy never returns 0 for evict_last -- and if it did the word would just be made again)
static int l2_policy(uint64_t* pol, cudaStream_t st) {
    constexpr int kMaxDev = 64;
    static std::atomic<uint64_t> cache[kMaxDev];
    int dev = 0;
    HT_CUDA_TRY(cudaGetDevice(&dev));
    if (dev >= 0 && dev < kMaxDev) {
        const uint64_t v = cache[dev].load(std::memory_order_acquire);
        if (v) {
            *pol = v;
            return HT_OK;
        }
    }
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cap) == cudaSuccess && cap != cudaStreamCaptureStatusNone) {
        set_error("ht_transform_u8: call once eagerly on this device before capturing it in a CUDA graph");
        return HT_ERR_UNSUPPORTED;
    }
    uint64_t* d = nullptr;
    HT_CUDA_TRY(cudaMalloc(&d, sizeof(uint64_t)));
    make_policy_kernel<<<1, 1, 0, st>>>(d);
    cudaError_t e = cudaMemcpyAsync(pol, d, sizeof(uint64_t), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d);
    if (e != cudaSuccess) return cuda_fail(e, "make_policy_kernel");
    if (dev >= 0 && dev < kMaxDev) cache[dev].store(*pol, std::memory_order_release);
    return HT_OK;
}

bool transform_half_enabled() {
    const char* e = getenv("HT_TRANSFORM_KERNEL");  // "tile" forces the 1024-sample form (A/B measurements)
    return !(e && e[0] == 't');
}

int launch_transform_u8_half(const uint32_t* planes, int64_t n_samples, int64_t n_keys,
                             const uint32_t* quad_off16, const uint4* quads, int64_t n_haps, uint8_t* out,
                             int64_t out_s_samp, unsigned long long* counts, cudaStream_t st,
                             const uint32_t* blk_key_lo, const uint32_t* blk_key_cnt, int64_t stage_lines) {
    uint64_t pol = 0;
    if (int rc = l2_policy(&pol, st)) return rc;
    const int64_t n_hb = (n_haps + kHapBlock - 1) / kHapBlock;
    const int64_t per_launch = max((int64_t)1, (int64_t)0x7fffffff / n_hb);
    const bool stage = blk_key_lo && blk_key_cnt && stage_lines > 0;
    // half tiles per CTA (HT_HALF_PER_CTA in the environment: measurements, tests)
    const char* hpc_e = getenv("HT_HALF_PER_CTA");
    const int hpc_env = hpc_e ? atoi(hpc_e) : 0;
    const int hpc = (hpc_env ? hpc_env : kHalfPerCta) >= 2 ? 2 : 1;
    const int64_t n_half = ((n_samples + kHalfSamples - 1) / kHalfSamples + hpc - 1) / hpc;  // groups of hpc half tiles
    const size_t smem = stage ? (size_t)stage_lines * 128 : 0;
    auto launch = [&](auto kernel) -> int {
        if (stage)  // per launch: the attribute is per device, and a caller may drive several devices from one thread
            HT_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        for (int64_t t0 = 0; t0 < n_half; t0 += per_launch) {
            const int64_t nt = min(per_launch, n_half - t0);
            kernel<<<dim3((unsigned)(nt * n_hb)), kThreads, smem, st>>>(
                planes, n_samples, n_keys, quad_off16, quads, stage ? blk_key_lo : nullptr, stage ? blk_key_cnt : nullptr,
                stage ? (uint32_t)stage_lines : 0u, n_haps, (uint32_t)n_hb, t0, out, out_s_samp, counts, pol);
            HT_CUDA_TRY(cudaGetLastError());
        }
        return HT_OK;
    };
    if (stage) return hpc == 2 ? launch(transform_u8_half_kernel<true, 2>) : launch(transform_u8_half_kernel<true, 1>);
    return hpc == 2 ? launch(transform_u8_half_kernel<false, 2>) : launch(transform_u8_half_kernel<false, 1>);
}

}  // namespace ht
