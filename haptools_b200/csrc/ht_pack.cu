// K1 -- pack (sm_100a): genotype bytes -> per-key bit-planes over samples.
//
// Replaces, without materialising either of them,
//   * the column gather `gts.subset(variants=...)`            haptools/data/haplotypes.py:1388
//                                                             -> haptools/data/genotypes.py:440
//   * `np.equal(allele_arr, gts.data[:, :, :2])`              haptools/data/haplotypes.py:1404
//                                                             haptools/transform.py:137
//   * `gts.ancestry[:, idxs[i]] == ancestries[i]`             haptools/transform.py:146
//     (folded into the key: bit = GT==allele && ANC==label)
//
// Plane layout (include/haptools_cuda.h): planes[tile][key][word 0..31][strand 0..1] uint32,
// tile = 1024 samples, bit b of word w = sample tile*1024 + 32*w + b, bits of samples >= n
// are 0.  A (tile, key) record is 256 contiguous bytes -- what one warp of the transform
// kernel reads per key.
//
// Three kernels, chosen from the strides of the genotype array:
//   generic        any strides, uint8 or pgenlib int32: byte loads + __ballot_sync
//   variant-major  samples contiguous (VCF reader layout): 128-bit coalesced loads, SWAR byte
//                  compare, carry-free multiply to gather match bits, 4-lane shuffle merge
//   sample-major   variants contiguous (PGEN reader layout / C-contiguous arrays): each thread
//                  streams a 4-byte column chunk down the sample rows, accumulating match
//                  bits bytewise, then byte-transposes with PRMT (details at the kernel)
//
// Roofline: HBM read of n*V*2 genotype bytes (+ ancestry) and write of A*2*ceil(n/32)*4.
#include <cstdlib>

#include "ht_pack.cuh"

namespace ht {

// -----------------------------------------------------------------------------------------
// generic kernel: warp per (tile, variant); lane = sample within a word; ballot per word
// -----------------------------------------------------------------------------------------
struct LoadU8 {
    // returns gt0 | gt1 << 8 | anc0 << 16 | anc1 << 24 of one sample
    template <bool kAnc>
    static __device__ __forceinline__ uint32_t load(const PackArgs& a, uint32_t col, int64_t s) {
        const uint8_t* g = a.gt + s * a.gs + (int64_t)col * a.gv;
        uint32_t x = (uint32_t)g[0] | ((uint32_t)g[a.gc] << 8);
        if (kAnc) {
            const uint8_t* l = a.anc + s * a.as + (int64_t)col * a.av;
            x |= ((uint32_t)l[0] << 16) | ((uint32_t)l[a.ac] << 24);
        }
        return x;
    }
};

struct LoadI32 {
    // pgenlib allele codes (PgenReader.read_alleles_list): int32 matrix (rows, 2*n), element
    // 2*s + c of row `col`; a.gv = row pitch in BYTES.  The reference turns -9 (missing) into -1
    // and stores the int32 into its uint8 matrix (haptools/data/genotypes.py:1388-1401), i.e.
    // -9 -> 255 and every other value modulo 256.
    template <bool kAnc>
    static __device__ __forceinline__ uint32_t load(const PackArgs& a, uint32_t col, int64_t s) {
        const int2 v = __ldg(reinterpret_cast<const int2*>(a.gt + (int64_t)col * a.gv) + s);
        const uint32_t b0 = v.x == -9 ? 255u : ((uint32_t)v.x & 0xffu);
        const uint32_t b1 = v.y == -9 ? 255u : ((uint32_t)v.y & 0xffu);
        uint32_t x = b0 | (b1 << 8);
        if (kAnc) {  // local-ancestry labels of the chunk: plain bytes with their own strides (row = `col`)
            const uint8_t* l = a.anc + s * a.as + (int64_t)col * a.av;
            x |= ((uint32_t)l[0] << 16) | ((uint32_t)l[a.ac] << 24);
        }
        return x;
    }
};

template <class Loader, bool kAnc>
__global__ void __launch_bounds__(kPackThreads)
pack_generic_kernel(const PackArgs a, uint32_t n_vblocks, int64_t tile0) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t tile = tile0 + blockIdx.x / n_vblocks;
    int64_t v = (int64_t)(blockIdx.x % n_vblocks) * kPackWarps + warp;
    if (v >= a.n_vars) return;
    if (a.var_sel) v = __ldg(a.var_sel + v);
    const uint32_t col = __ldg(a.var_col + v);
    const uint32_t kb = __ldg(a.var_key_off + v), ke = __ldg(a.var_key_off + v + 1);
    const int64_t s0 = tile * kTileSamples + lane;

    uint32_t x[32];
    uint32_t fl = 0;
#pragma unroll
    for (int w = 0; w < 32; ++w) {
        const int64_t s = s0 + 32 * w;
        x[w] = 0;
        if (s < a.n) {
            x[w] = Loader::template load<kAnc>(a, col, s);
            fl |= qc_flags4(x[w] & 0xffffu);
        }
    }
    publish_flags(a.flags, fl);

    constexpr uint32_t m0 = kAnc ? 0x00ff00ffu : 0x000000ffu;
    constexpr uint32_t m1 = kAnc ? 0xff00ff00u : 0x0000ff00u;
    for (uint32_t k = kb; k < ke; ++k) {
        uint32_t pat = (uint32_t)__ldg(a.key_allele + k) * 0x0101u;
        if (kAnc) pat |= (uint32_t)__ldg(a.key_label + k) * 0x01010000u;
        uint2 mine = make_uint2(0u, 0u);
#pragma unroll
        for (int w = 0; w < 32; ++w) {
            const bool valid = (s0 + 32 * w) < a.n;
            const uint32_t d = x[w] ^ pat;
            const uint32_t b0 = __ballot_sync(0xffffffffu, valid && (d & m0) == 0);
            const uint32_t b1 = __ballot_sync(0xffffffffu, valid && (d & m1) == 0);
            if (lane == w) mine = make_uint2(b0, b1);
        }
        record(a, tile, k)[lane] = mine;  // 256 B coalesced
    }
}

// -----------------------------------------------------------------------------------------
// variant-major kernel (s_strand == 1, s_samp == 2): warp per (tile, variant)
// -----------------------------------------------------------------------------------------
// 16 bytes = 8 samples x 2 strands of one variant; `nvalid` = how many of the 8 samples exist
__device__ __forceinline__ uint4 load_chunk16(const uint8_t* p, int nvalid) {
    if (nvalid >= 8) return __ldg(reinterpret_cast<const uint4*>(p));
    uint32_t w[4] = {0u, 0u, 0u, 0u};
    for (int i = 0; i < 2 * nvalid; ++i) w[i >> 2] |= (uint32_t)p[i] << (8 * (i & 3));
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// one (tile, variant) item of the variant-major kernel; executed by a whole warp.  The row loads of
// an item (vm_fetch) are issued one item AHEAD of its arithmetic (pack_vm_item), so a warp always
// has 2 KB (4 KB with ancestry) in flight instead of waiting out every item's DRAM latency.
template <bool kAnc>
struct VmItem {
    uint4 x[4];             // genotype bytes: 4 chunks of 8 samples x 2 strands
    uint4 q[kAnc ? 4 : 1];  // ancestry bytes, same chunks
    uint32_t kb, ke;        // the variant's keys
};

// column and key range of a variant: fetched TWO items ahead of the arithmetic, so that the row loads of
// the next item (one ahead) never wait for their own address (a warp issues in order: it would sit out
// an L2 round trip at the first use of `col` before it could even start on the current item)
struct VmIdx {
    uint32_t col, kb, ke;
};

__device__ __forceinline__ VmIdx vm_fetch_idx(const PackArgs& a, int64_t v) {
    VmIdx ix;
    ix.col = __ldg(a.var_col + v);
    ix.kb = __ldg(a.var_key_off + v);
    ix.ke = __ldg(a.var_key_off + v + 1);
    return ix;
}

template <bool kAnc>
__device__ __forceinline__ void vm_fetch(const PackArgs& a, int64_t tile, const VmIdx& ix, int lane, VmItem<kAnc>& it) {
    const uint32_t col = ix.col;
    it.kb = ix.kb;
    it.ke = ix.ke;
    const int rows = (int)min((int64_t)kTileSamples, a.n - tile * kTileSamples);
    const uint8_t* g = a.gt + (int64_t)col * a.gv + tile * (2 * kTileSamples);
    const uint8_t* l = kAnc ? a.anc + (int64_t)col * a.av + tile * (2 * kTileSamples) : nullptr;
    if (rows == kTileSamples) {  // every tile but the last: four unconditional 16-byte loads per lane
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            it.x[j] = __ldg(reinterpret_cast<const uint4*>(g) + j * 32 + lane);
            if (kAnc) it.q[j] = __ldg(reinterpret_cast<const uint4*>(l) + j * 32 + lane);
        }
        return;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int c = j * 32 + lane;
        const int nvalid = rows - 8 * c;
        it.x[j] = make_uint4(0u, 0u, 0u, 0u);
        if (nvalid > 0) it.x[j] = load_chunk16(g + 16 * c, nvalid);
        if (kAnc) {
            it.q[j] = make_uint4(0u, 0u, 0u, 0u);
            if (nvalid > 0) it.q[j] = load_chunk16(l + 16 * c, nvalid);
        }
    }
}

// Lane l holds, for j = 0..3, the byte M.byte[j] = 8 sample bits of chunk 32 j + l, i.e. byte (l & 3)
// of word 8 j + (l >> 2) of one strand's plane.  Returns word `lane` of the plane: a 4x4 byte transpose
// inside every group of 4 lanes (2 shuffles) leaves lane 4 g + j with the whole word 8 j + g, one more
// shuffle brings word L to lane L -- so that a warp stores a record with ONE coalesced instruction.
__device__ __forceinline__ uint32_t vm_gather_word(uint32_t M, int lane) {
    const uint32_t p = __shfl_xor_sync(0xffffffffu, M, 1);
    const uint32_t y = prmt(M, p, (lane & 1) ? 0x3715u : 0x6240u);  // even: [M0 p0 M2 p2]; odd: [p1 M1 p3 M3]
    const uint32_t q = __shfl_xor_sync(0xffffffffu, y, 2);
    const uint32_t w = prmt(y, q, (lane & 2) ? 0x3276u : 0x5410u);  // lane 4g+j: bytes j of lanes 4g..4g+3
    return __shfl_sync(0xffffffffu, w, 4 * (lane & 7) + (lane >> 3));
}

template <bool kAnc>
__device__ __forceinline__ void pack_vm_item(const PackArgs& a, int64_t tile, const VmItem<kAnc>& it, int lane) {
    const uint32_t kb = it.kb, ke = it.ke;

    if (a.flags) {
        uint32_t fl = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) fl |= qc_flags4(it.x[j].x) | qc_flags4(it.x[j].y) | qc_flags4(it.x[j].z) | qc_flags4(it.x[j].w);
        publish_flags(a.flags, fl);
    }
    const uint32_t vm = valid_mask(a.n, tile, lane);  // lane l stores word l of a record

    if (!kAnc) {
        // Biallelic single pass: if the variant's keys are just alleles {0, 1} and no byte of the
        // row is >= 2, the allele-1 plane is the bytes' low bit and the allele-0 plane its
        // complement.  A chunk is 4 words x = {s.c0, s.c1, (s+1).c0, (s+1).c1} of 0/1 bytes (samples
        // 2i, 2i+1 in word i); x * 0x10002 puts the word's two strand-0 bits next to each other at bits
        // 16, 17 and its two strand-1 bits at 24, 25 (the other partial products land in bits 1 and 9),
        // and the four words shifted by 2i tile bits 16..23 / 24..31 without a carry: FOUR multiply-adds
        // turn 16 genotype bytes into the 8 sample bits of both strands (bytes 2 and 3 of r).
        const uint32_t nk = ke - kb;
        const uint32_t al0 = nk >= 1 ? __ldg(a.key_allele + kb) : 0u;
        const uint32_t al1 = nk >= 2 ? __ldg(a.key_allele + kb + 1) : 1u;
        uint32_t d = 0u;
#pragma unroll
        for (int j = 0; j < 4; ++j) d |= (it.x[j].x | it.x[j].y) | (it.x[j].z | it.x[j].w);
        const bool dirty = __any_sync(0xffffffffu, (d & 0xfefefefeu) != 0u);
        if (nk <= 2 && al0 <= 1u && al1 <= 1u && !dirty) {
            // keys are sorted by allele: (0), (1) or (0, 1)
            const int k_ref = al0 == 0u ? (int)kb : -1;
            const int k_alt = al0 == 1u ? (int)kb : (nk == 2 ? (int)kb + 1 : -1);
            uint32_t r[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                r[j] = it.x[j].x * 0x00010002u + it.x[j].y * 0x00040008u + it.x[j].z * 0x00100020u + it.x[j].w * 0x00400080u;
            const uint32_t t0 = prmt(r[0], r[1], 0x7362), t1 = prmt(r[2], r[3], 0x7362);  // [r0.b2 r1.b2 r0.b3 r1.b3]
            const uint32_t M0 = prmt(t0, t1, 0x5410);  // byte j = 8 samples of chunk j, strand 0
            const uint32_t M1 = prmt(t0, t1, 0x7632);  // strand 1
            const uint32_t w0 = vm_gather_word(M0, lane), w1 = vm_gather_word(M1, lane);
            if (k_alt >= 0) record(a, tile, (uint32_t)k_alt)[lane] = make_uint2(w0 & vm, w1 & vm);
            if (k_ref >= 0) record(a, tile, (uint32_t)k_ref)[lane] = make_uint2(~w0 & vm, ~w1 & vm);
            return;
        }
    }

    // y = strand-0 bytes, z = strand-1 bytes of the lane's 4 chunks (8 samples each)
    uint32_t y[4][2], z[4][2], ay[4][2], az[4][2];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint4 x = it.x[j];
        y[j][0] = prmt(x.x, x.y, 0x6420);
        y[j][1] = prmt(x.z, x.w, 0x6420);
        z[j][0] = prmt(x.x, x.y, 0x7531);
        z[j][1] = prmt(x.z, x.w, 0x7531);
        if (kAnc) {
            const uint4 q = it.q[j];
            ay[j][0] = prmt(q.x, q.y, 0x6420);
            ay[j][1] = prmt(q.z, q.w, 0x6420);
            az[j][0] = prmt(q.x, q.y, 0x7531);
            az[j][1] = prmt(q.z, q.w, 0x7531);
        }
    }
#pragma unroll 1
    for (uint32_t k = kb; k < ke; ++k) {
        const uint32_t pat = (uint32_t)__ldg(a.key_allele + k) * 0x01010101u;
        const uint32_t lpat = kAnc ? (uint32_t)__ldg(a.key_label + k) * 0x01010101u : 0u;
        uint32_t M0 = 0u, M1 = 0u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint32_t f0a = eq_bytes_msb(y[j][0], pat), f0b = eq_bytes_msb(y[j][1], pat);
            uint32_t f1a = eq_bytes_msb(z[j][0], pat), f1b = eq_bytes_msb(z[j][1], pat);
            if (kAnc) {
                f0a &= eq_bytes_msb(ay[j][0], lpat);
                f0b &= eq_bytes_msb(ay[j][1], lpat);
                f1a &= eq_bytes_msb(az[j][0], lpat);
                f1b &= eq_bytes_msb(az[j][1], lpat);
            }
            M0 |= (nib(f0a) | (nib(f0b) << 4)) << (8 * j);  // 8 samples, strand 0
            M1 |= (nib(f1a) | (nib(f1b) << 4)) << (8 * j);  // 8 samples, strand 1
        }
        record(a, tile, k)[lane] = make_uint2(vm_gather_word(M0, lane) & vm, vm_gather_word(M1, lane) & vm);
    }
}
// loop state of a warp of the variant-major kernel.  Items are numbered hi * radix + lo and a warp's items
// advance by a constant step, so (hi, lo) are stepped without a 64-bit division.  Two orders:
//   tile-major     hi = tile, lo = variant (radix = n_vars): the 8 warps of a CTA write 8 neighbouring records
//   variant-major  hi = variant, lo = tile (radix = n_tiles): neighbouring warps read neighbouring 2 KB pieces
//                  of ONE variant row (a whole 5 KB row at 2,504 samples) -- for cohorts of a few tiles, where
//                  the tile-major order touches a row once per tile, a DRAM page apart each time
struct VmWalk {
    int64_t item, n_items, stride;
    int32_t radix, step_hi, step_lo, hi_n, lo_n, tile;  // tile: of the current item; *_n: next item
    bool vfirst;
    VmIdx ix_n;  // index entries of the next item
};

// one iteration: the row loads of the next item go out into `nxt`, the index entries of the item after it are
// asked for, then the current item (`cur`, loaded one iteration ago) is packed.  (Calling it with the two
// buffers in alternating roles instead of copying nxt to cur costs 268 bytes of spills at 64 registers:
// 0.189 against 0.125 ms on the 1000G-scale config.)
template <bool kAnc>
__device__ __forceinline__ void vm_step(const PackArgs& a, VmWalk& w, int lane, const VmItem<kAnc>& cur, VmItem<kAnc>& nxt) {
    int32_t hi_nn = w.hi_n + w.step_hi, lo_nn = w.lo_n + w.step_lo;
    if (lo_nn >= w.radix) {
        lo_nn -= w.radix;
        ++hi_nn;
    }
    const int32_t tile_n = w.vfirst ? w.lo_n : w.hi_n;
    VmIdx ix_nn = {0u, 0u, 0u};
    if (w.item + w.stride < w.n_items) vm_fetch<kAnc>(a, tile_n, w.ix_n, lane, nxt);
    if (w.item + 2 * w.stride < w.n_items) ix_nn = vm_fetch_idx(a, w.vfirst ? hi_nn : lo_nn);
    pack_vm_item<kAnc>(a, w.tile, cur, lane);
    w.tile = tile_n;
    w.hi_n = hi_nn;
    w.lo_n = lo_nn;
    w.ix_n = ix_nn;
    w.item += w.stride;
}

template <bool kAnc>
__global__ void __launch_bounds__(kPackThreads, kAnc ? 2 : 4)
pack_vm_kernel(const PackArgs a, int64_t n_tiles, bool vfirst) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    VmWalk w;
    w.vfirst = vfirst;
    w.n_items = n_tiles * a.n_vars;
    w.stride = (int64_t)gridDim.x * kPackWarps;
    w.item = (int64_t)blockIdx.x * kPackWarps + warp;
    if (w.item >= w.n_items) return;
    w.radix = (int32_t)(vfirst ? n_tiles : a.n_vars);
    w.step_hi = (int32_t)(w.stride / w.radix);
    w.step_lo = (int32_t)(w.stride % w.radix);
    const int32_t hi = (int32_t)(w.item / w.radix), lo = (int32_t)(w.item % w.radix);
    w.tile = vfirst ? lo : hi;
    VmItem<kAnc> b0, b1;
    vm_fetch<kAnc>(a, w.tile, vm_fetch_idx(a, vfirst ? hi : lo), lane, b0);
    w.hi_n = hi + w.step_hi;
    w.lo_n = lo + w.step_lo;
    if (w.lo_n >= w.radix) {
        w.lo_n -= w.radix;
        ++w.hi_n;
    }
    w.ix_n = VmIdx{0u, 0u, 0u};
    if (w.item + w.stride < w.n_items) w.ix_n = vm_fetch_idx(a, vfirst ? w.hi_n : w.lo_n);
#pragma unroll 1
    while (w.item < w.n_items) {
        vm_step<kAnc>(a, w, lane, b0, b1);
        b0 = b1;
    }
}

// -----------------------------------------------------------------------------------------
// sample-major kernel (s_strand == 1, s_var == 2)
//   CTA  = (tile of 1024 samples) x (512 consecutive columns = 1 KB of every sample row)
//   warp = 64 of those columns, all 1024 rows in 8 groups of 128 rows (4 words)
//   lane = 2 columns = 4 bytes of every row {col0.s0, col0.s1, col1.s0, col1.s1}
// Per row the lane folds its 4 bytes into 4 byte-wide shift registers (one per byte column);
// after 8 rows each byte holds 8 sample bits, after 32 rows a 4x4 byte transpose (PRMT) gives
// one 32-sample word per byte column.  Per 128 rows every (key, lane) pair stores one full
// 32-byte sector of the key's record.
//
// Columns whose keys are just alleles {0, 1} without ancestry take the single-pass path: the
// allele-1 plane is the bytes' low bit, the allele-0 plane its complement, both masked by an
// "invalid" plane (byte >= 2: missing or multi-allelic call) that is only built for 8-row
// blocks that contain such a byte -- exact for any input, 3 ALU ops per 4 bytes when clean.
// Other columns (multi-allelic keys, ancestry labels) take one SWAR-compare pass per key.
// -----------------------------------------------------------------------------------------
constexpr int kSmLaneCols = 2;
constexpr int kSmWarpCols = 32 * kSmLaneCols;            // 64
constexpr int kSmCtaCols = kPackWarps * kSmWarpCols;     // 512

__device__ __forceinline__ uint32_t load_cols4(const uint8_t* p, int ncols) {
    if (ncols >= 2) return __ldg(reinterpret_cast<const uint32_t*>(p));
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8);
}

// One 32-byte sector of a key's record: words 4*grp .. 4*grp+3, both strands, of byte-column
// pair i, read back from the thread's private staging column.  stage[w4*8 + k] = value word
// k of word w4, stage[w4*8 + 4 + k] = invalid word.
__device__ __forceinline__ void store_sector(uint2* rec4, const uint32_t* stage, int i, bool invert,
                                             const uint32_t (&vm)[4]) {
    uint32_t o[8];
#pragma unroll
    for (int w4 = 0; w4 < 4; ++w4) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const uint32_t v = stage[(w4 * 8 + 2 * i + c) * kPackThreads];
            const uint32_t z = stage[(w4 * 8 + 4 + 2 * i + c) * kPackThreads];
            o[2 * w4 + c] = (invert ? ~v : v) & ~z & vm[w4];
        }
    }
    uint4* dst = reinterpret_cast<uint4*>(rec4);
#ifdef HT_EXP_NO_PLANE_STORE  // experiment only (profiles/): keep the math alive, drop the stores
    if ((o[0] ^ o[1] ^ o[2] ^ o[3] ^ o[4] ^ o[5] ^ o[6] ^ o[7]) != 0x9e3779b9u) return;
#endif
    dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
    dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
}

// staging words per thread: plain = 4 words x {value[4], invalid[4]}; ancestry = 4 words x
// {alt[4], label bit 0..2 [4] each, invalid[4]}
constexpr int kSmStagePlain = 32, kSmStageAnc = 80;

template <bool kAnc>
__global__ void __launch_bounds__(kPackThreads, kAnc ? 2 : 4)
pack_sm_kernel(const PackArgs a, uint32_t n_cblocks, int64_t cblock0, int64_t tile0) {
    // thread-private staging (column = thread): keeps the per-word loop rolled, i.e. the code
    // small enough for the instruction cache
    extern __shared__ uint32_t s_stage[];  // (kAnc ? 80 : 32) * kPackThreads words
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t tile = tile0 + blockIdx.x / n_cblocks;
    const int64_t col0 =
        ((cblock0 + blockIdx.x % n_cblocks) * kPackWarps + warp) * kSmWarpCols + kSmLaneCols * lane;
    const int ncols = (int)min((int64_t)kSmLaneCols, a.n_cols - col0);  // may be <= 0
    if (ncols <= 0) return;

    // which of my columns are referenced, and by which keys
    uint32_t kb[2] = {0u, 0u}, nk[2] = {0u, 0u};
    int k_ref[2] = {-1, -1}, k_alt[2] = {-1, -1};  // key of allele 0 / 1 (single-pass path)
    bool simple = !kAnc;
    bool simple_anc = kAnc;  // every key: allele <= 1 and label < 8 -> single-pass ancestry path
    uint32_t refmask = 0u, rounds = 0u;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        if (i < ncols) {
            const int32_t v = __ldg(a.col_var + col0 + i);
            if (v >= 0) {
                kb[i] = __ldg(a.var_key_off + v);
                nk[i] = __ldg(a.var_key_off + v + 1) - kb[i];
                refmask |= 0xffffu << (16 * i);
                rounds = max(rounds, nk[i]);
                if (!kAnc) {
                    // keys of a variant are sorted by allele and distinct
                    for (uint32_t k = 0; k < min(nk[i], 2u); ++k) {
                        const uint32_t al = __ldg(a.key_allele + kb[i] + k);
                        if (al == 0) k_ref[i] = (int)(kb[i] + k);
                        else if (al == 1) k_alt[i] = (int)(kb[i] + k);
                        else simple = false;
                    }
                    simple = simple && nk[i] <= 2;
                } else {
                    for (uint32_t k = 0; k < nk[i]; ++k)
                        simple_anc = simple_anc && __ldg(a.key_allele + kb[i] + k) <= 1 &&
                                     __ldg(a.key_label + kb[i] + k) < 8;
                }
            }
        }
    }
    if (rounds == 0) return;

    const uint8_t* g = a.gt + col0 * 2;
    const uint8_t* l = kAnc ? a.anc + col0 * 2 : nullptr;
    uint32_t* stage = s_stage + threadIdx.x;
    uint32_t fl = 0;

#pragma unroll 1
    for (int grp = 0; grp < 8; ++grp) {
        const int64_t s_grp = tile * kTileSamples + 128 * grp;
        if (s_grp >= a.n) break;
        uint32_t vm[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) vm[k] = valid_mask(a.n, tile, 4 * grp + k);

        if (kAnc && simple_anc) {
            // Single pass for (allele, label) keys: accumulate the allele-1 plane and the three
            // low bit-planes of the label bytes, plus an "invalid" plane (GT byte >= 2 or label
            // byte >= 8) for 8-row blocks that contain such a byte.  Any key's plane is then
            // alt^ma & L0^m0 & L1^m1 & L2^m2 & ~invalid.
            const bool full = (s_grp + 128 <= a.n) && ncols == kSmLaneCols;
            const uint8_t* p = g + s_grp * a.gs;
            const uint8_t* pl = l + s_grp * a.as;
#pragma unroll 1
            for (int w4 = 0; w4 < 4; ++w4) {
                uint32_t GA[4], G0[4], G1[4], G2[4], GZ[4];
                bool any_dirty = false;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint32_t x[8], c[8];
                    if (full) {
#pragma unroll
                        for (int rr = 0; rr < 8; ++rr, p += a.gs, pl += a.as) {
                            x[rr] = __ldg(reinterpret_cast<const uint32_t*>(p));
                            c[rr] = __ldg(reinterpret_cast<const uint32_t*>(pl));
                        }
                    } else {
#pragma unroll
                        for (int rr = 0; rr < 8; ++rr, p += a.gs, pl += a.as) {
                            x[rr] = c[rr] = 0u;
                            if (s_grp + 32 * w4 + 8 * q + rr < a.n) {
                                x[rr] = load_cols4(p, ncols);
                                c[rr] = load_cols4(pl, ncols);
                            }
                        }
                    }
                    uint32_t aA = 0u, a0 = 0u, a1 = 0u, a2 = 0u, d = 0u;
#pragma unroll
                    for (int rr = 7; rr >= 0; --rr) {  // row rr ends at bit rr of its byte
                        aA = (aA << 1) + (x[rr] & 0x01010101u);
                        a0 = (a0 << 1) + (c[rr] & 0x01010101u);
                        a1 = (a1 << 1) + ((c[rr] >> 1) & 0x01010101u);
                        a2 = (a2 << 1) + ((c[rr] >> 2) & 0x01010101u);
                        d |= (x[rr] & 0xfefefefeu) | (c[rr] & 0xf8f8f8f8u);
                    }
                    GA[q] = aA;
                    G0[q] = a0;
                    G1[q] = a1;
                    G2[q] = a2;
                    GZ[q] = 0u;
                    if (d) {
                        any_dirty = true;
                        uint32_t z = 0u;
#pragma unroll
                        for (int rr = 7; rr >= 0; --rr) {
                            const uint32_t y = (x[rr] & 0xfefefefeu) | (c[rr] & 0xf8f8f8f8u);
                            z = (z << 1) + (nz_bytes_msb(y) >> 7);
                        }
                        GZ[q] = z;
                        if (a.flags) {
#pragma unroll
                            for (int rr = 0; rr < 8; ++rr) fl |= qc_flags4(x[rr] & refmask);
                        }
                    }
                }
                uint32_t w[4], z[4] = {0u, 0u, 0u, 0u};
                uint32_t* st = stage + (w4 * 20) * kPackThreads;
                byte_transpose4(GA, w);
#pragma unroll
                for (int k = 0; k < 4; ++k) st[k * kPackThreads] = w[k];
                byte_transpose4(G0, w);
#pragma unroll
                for (int k = 0; k < 4; ++k) st[(4 + k) * kPackThreads] = w[k];
                byte_transpose4(G1, w);
#pragma unroll
                for (int k = 0; k < 4; ++k) st[(8 + k) * kPackThreads] = w[k];
                byte_transpose4(G2, w);
#pragma unroll
                for (int k = 0; k < 4; ++k) st[(12 + k) * kPackThreads] = w[k];
                if (any_dirty) byte_transpose4(GZ, z);
#pragma unroll
                for (int k = 0; k < 4; ++k) st[(16 + k) * kPackThreads] = z[k];
            }
#pragma unroll
            for (int i = 0; i < 2; ++i) {
#pragma unroll 1
                for (uint32_t r = 0; r < nk[i]; ++r) {
                    const uint32_t al = __ldg(a.key_allele + kb[i] + r), lb = __ldg(a.key_label + kb[i] + r);
                    const uint32_t mA = al ? 0u : 0xffffffffu, m0 = (lb & 1u) ? 0u : 0xffffffffu;
                    const uint32_t m1 = (lb & 2u) ? 0u : 0xffffffffu, m2 = (lb & 4u) ? 0u : 0xffffffffu;
                    uint32_t o[8];
#pragma unroll
                    for (int w4 = 0; w4 < 4; ++w4) {
#pragma unroll
                        for (int c = 0; c < 2; ++c) {
                            const uint32_t* st = stage + (w4 * 20 + 2 * i + c) * kPackThreads;
                            const uint32_t v = (st[0] ^ mA) & (st[4 * kPackThreads] ^ m0) &
                                               (st[8 * kPackThreads] ^ m1) & (st[12 * kPackThreads] ^ m2);
                            o[2 * w4 + c] = v & ~st[16 * kPackThreads] & vm[w4];
                        }
                    }
                    uint4* dst = reinterpret_cast<uint4*>(record(a, tile, kb[i] + r) + 4 * grp);
                    dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
                    dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
                }
            }
        } else if (simple) {
            // all 128 rows exist and both of my columns do: no per-row checks, and the 32 row
            // loads of a word are issued back to back (32 independent loads in flight)
            const bool full = (s_grp + 128 <= a.n) && ncols == kSmLaneCols;
            const uint8_t* p = g + s_grp * a.gs;
#pragma unroll 1
            for (int w4 = 0; w4 < 4; ++w4) {
                uint32_t x[32];
                if (full) {
#pragma unroll
                    for (int rr = 0; rr < 32; ++rr, p += a.gs) x[rr] = __ldg(reinterpret_cast<const uint32_t*>(p));
                } else {
#pragma unroll
                    for (int rr = 0; rr < 32; ++rr, p += a.gs) {
                        x[rr] = 0u;
                        if (s_grp + 32 * w4 + rr < a.n) x[rr] = load_cols4(p, ncols);
                    }
                }
                uint32_t GA[4], GZ[4];
                bool any_dirty = false;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint32_t acc = 0u, d = 0u;
#pragma unroll
                    for (int rr = 7; rr >= 0; --rr) {  // row rr ends at bit rr of its byte
                        acc = (acc << 1) + (x[8 * q + rr] & 0x01010101u);
                        d |= x[8 * q + rr];
                    }
                    GA[q] = acc;
                    GZ[q] = 0u;
                    if (d & 0xfefefefeu) {  // some byte >= 2: build the invalid plane of this block
                        any_dirty = true;
                        uint32_t z = 0u;
#pragma unroll
                        for (int rr = 7; rr >= 0; --rr) {
                            const uint32_t v = x[8 * q + rr];
                            const uint32_t t = (((v & 0x7e7e7e7eu) + 0x7e7e7e7eu) | v) & 0x80808080u;
                            z = (z << 1) + (t >> 7);
                        }
                        GZ[q] = z;
                        if (a.flags) {
#pragma unroll
                            for (int rr = 0; rr < 8; ++rr) fl |= qc_flags4(x[8 * q + rr] & refmask);
                        }
                    }
                }
                uint32_t w[4], z[4] = {0u, 0u, 0u, 0u};
                byte_transpose4(GA, w);
                if (any_dirty) byte_transpose4(GZ, z);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    stage[(w4 * 8 + k) * kPackThreads] = w[k];
                    stage[(w4 * 8 + 4 + k) * kPackThreads] = z[k];
                }
            }
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                if (k_alt[i] >= 0) store_sector(record(a, tile, (uint32_t)k_alt[i]) + 4 * grp, stage, i, false, vm);
                if (k_ref[i] >= 0) store_sector(record(a, tile, (uint32_t)k_ref[i]) + 4 * grp, stage, i, true, vm);
            }
        } else {
#pragma unroll 1
            for (uint32_t r = 0; r < rounds; ++r) {
                uint32_t pb[2] = {0u, 0u}, lb[2] = {0u, 0u};
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    if (r < nk[i]) {
                        pb[i] = __ldg(a.key_allele + kb[i] + r);
                        if (kAnc) lb[i] = __ldg(a.key_label + kb[i] + r);
                    }
                }
                const uint32_t pat = (pb[0] | (pb[1] << 16)) * 0x0101u;
                const uint32_t lpat = (lb[0] | (lb[1] << 16)) * 0x0101u;
#pragma unroll 1
                for (int w4 = 0; w4 < 4; ++w4) {
                    uint32_t G[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int64_t s_first = s_grp + 32 * w4 + 8 * q;
                        uint32_t x[8], c[8];
#pragma unroll
                        for (int rr = 0; rr < 8; ++rr) {
                            x[rr] = 0u;
                            c[rr] = 0u;
                            if (s_first + rr < a.n) {
                                x[rr] = load_cols4(g + (s_first + rr) * a.gs, ncols);
                                if (kAnc) c[rr] = load_cols4(l + (s_first + rr) * a.as, ncols);
                            }
                        }
                        uint32_t acc = 0u;
#pragma unroll
                        for (int rr = 0; rr < 8; ++rr) {
                            uint32_t f = eq_bytes_msb(x[rr], pat);
                            if (kAnc) f &= eq_bytes_msb(c[rr], lpat);
                            acc = (acc >> 1) | f;  // row rr enters at bit 7, ends at bit rr
                            if (r == 0 && a.flags) fl |= qc_flags4(x[rr] & refmask);
                        }
                        G[q] = acc;
                    }
                    uint32_t w[4];
                    byte_transpose4(G, w);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        stage[(w4 * 8 + k) * kPackThreads] = w[k];
                        stage[(w4 * 8 + 4 + k) * kPackThreads] = 0u;
                    }
                }
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    if (r < nk[i]) store_sector(record(a, tile, kb[i] + r) + 4 * grp, stage, i, false, vm);
                }
            }
        }
    }
    publish_flags(a.flags, fl);
}

// -----------------------------------------------------------------------------------------
// sample-major kernel, biallelic fast path (s_strand == 1, s_var == 2, 16-byte aligned rows)
//   CTA  = (tile of 1024 samples) x (64 consecutive columns); 8 warps = 2 row halves x 4
//          column chunks of 16 columns
//   warp = 16 words (512 rows) x 16 columns
//   lane = (word wl = lane / 2, sub-chunk = lane % 2): 8 columns = 16 bytes of each of the 32
//          rows of ONE word -- two lanes make a full 32-byte sector of every row
// The lane folds its 32 rows into byte-wide shift registers and byte-transposes (as above);
// it then owns word `wl` of 8 columns x 2 strands, so the lanes with the same sub-chunk hold
// 16 consecutive words of a key: every store instruction writes 2 x 128 contiguous bytes of
// plane records and no shared-memory staging is needed.  (A profile of the previous mapping
// showed its scattered 16-byte stores doubling the kernel time although planes are 20 % of
// the bytes.)  Only keys "allele 0" / "allele 1" are handled (dense table col_key01 from the
// planner); variants with other keys go through the generic kernel (var_sel).
// Missing / multi-allelic bytes (>= 2) are exact: a lane that saw one re-reads its rows and
// builds the invalid plane.
// -----------------------------------------------------------------------------------------
constexpr int kSm2CtaCols = 64;

__device__ __forceinline__ uint4 load_row16(const uint8_t* p, int ncols) {
    if (ncols >= 8) return __ldg(reinterpret_cast<const uint4*>(p));
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {  // word j = my columns 2j, 2j+1 (2 bytes each; rows are 2-byte aligned)
        w[j] = 0u;
        if (2 * j < ncols) w[j] = *reinterpret_cast<const uint16_t*>(p + 4 * j);
        if (2 * j + 1 < ncols) w[j] |= (uint32_t)*reinterpret_cast<const uint16_t*>(p + 4 * j + 2) << 16;
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// byte k of `src` (k = 0..3) replaces byte q of each of dst[0..3]
__device__ __forceinline__ void insert_bytes(uint32_t (&dst)[4], uint32_t src, int q) {
    // PRMT selector: identity 0x3210 with nibble q replaced by 4 + k (bytes 4..7 = src)
    const uint32_t keep = 0x3210u & ~(0xfu << (4 * q));
#pragma unroll
    for (int k = 0; k < 4; ++k) dst[k] = prmt(dst[k], src, keep | ((4u + k) << (4 * q)));
}

// kSlab: the mapping for LONG rows (pitch of a MB and more: every row of a tile lies in its own 2 MB
// page, 1,024 pages per tile against a TLB reach of 128).  CTA = 4 words (128 rows) x 512 columns, the
// grid is slab-major inside a tile, so all CTAs in flight read the same 128 rows (128 pages) and a CTA
// reads 1 KB of every row; warp = 4 words x 64 columns (a load instruction reads 4 rows x 128 bytes), a
// record is stored 4 words (one 32-byte sector) per instruction.  n_cblocks then counts (slab, 512-column
// block) pairs: 8 * ceil(n_cols / 512).
constexpr int kSm2SlabCols = 512, kSm2SlabsPerTile = 8;
constexpr int64_t kSm2SlabMinPitch = 0;  // bytes between rows from which the slab mapping is used: always (it is
                                         // faster at every pitch measured, 100 KB to 2 MB: profiles/exp_pack_pitch2.py)

template <bool kSlab>
__global__ void __launch_bounds__(kPackThreads, 3)
pack_sm2_kernel(const PackArgs a, uint32_t n_cblocks, int64_t cblock0, int64_t tile0) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t tile = tile0 + blockIdx.x / n_cblocks;
    const uint32_t xb = blockIdx.x % n_cblocks;
    const uint32_t n_cb = kSlab ? n_cblocks / kSm2SlabsPerTile : n_cblocks;
    const int word = kSlab ? (int)(xb / n_cb) * 4 + (lane >> 3) : (warp >> 2) * 16 + (lane >> 1);
    const int64_t col0 = kSlab ? (cblock0 + xb % n_cb) * kSm2SlabCols + warp * 64 + (lane & 7) * 8
                               : (cblock0 + xb) * kSm2CtaCols + (warp & 3) * 16 + (lane & 1) * 8;
    const int ncols = (int)min((int64_t)8, a.n_cols - col0);  // may be <= 0
    if (ncols <= 0) return;
    const int64_t s0 = tile * kTileSamples + 32 * word;
    const int nrows = (int)min((int64_t)32, a.n - s0);  // <= 0: the word lies beyond the samples

    // keys of my 8 columns: {allele 0, allele 1}, -1 = no such key
    int2 kk[8];
    bool any = false;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        kk[i] = make_int2(-1, -1);
        if (i < ncols) kk[i] = __ldg(reinterpret_cast<const int2*>(a.col_key01) + col0 + i);
        any = any || kk[i].x >= 0 || kk[i].y >= 0;
    }
    if (!any) return;

    const uint8_t* base = a.gt + s0 * a.gs + col0 * 2;
    const bool full = nrows == 32 && ncols == 8;
    // W[j][k]: word of byte column 4*j + k = my column 2*j + k/2, strand k & 1
    uint32_t W[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int k = 0; k < 4; ++k) W[j][k] = 0u;
    uint32_t d = 0u;
    {
        const uint8_t* p = base;
#pragma unroll 1
        for (int q = 0; q < 4; ++q) {  // 8-row blocks; 8 independent 16-byte loads in flight
            uint4 x[8];
            if (full) {
#pragma unroll
                for (int rr = 0; rr < 8; ++rr, p += a.gs) x[rr] = __ldg(reinterpret_cast<const uint4*>(p));
            } else {
#pragma unroll
                for (int rr = 0; rr < 8; ++rr, p += a.gs) {
                    x[rr] = make_uint4(0u, 0u, 0u, 0u);
                    if (8 * q + rr < nrows) x[rr] = load_row16(p, ncols);
                }
            }
            uint32_t a0 = 0u, a1 = 0u, a2 = 0u, a3 = 0u;
#pragma unroll
            for (int rr = 7; rr >= 0; --rr) {  // row rr ends at bit rr of its byte
                a0 = (a0 << 1) + (x[rr].x & 0x01010101u);
                a1 = (a1 << 1) + (x[rr].y & 0x01010101u);
                a2 = (a2 << 1) + (x[rr].z & 0x01010101u);
                a3 = (a3 << 1) + (x[rr].w & 0x01010101u);
                d |= (x[rr].x | x[rr].y) | (x[rr].z | x[rr].w);
            }
            insert_bytes(W[0], a0, q);
            insert_bytes(W[1], a1, q);
            insert_bytes(W[2], a2, q);
            insert_bytes(W[3], a3, q);
        }
    }
    const uint32_t vm = valid_mask(a.n, tile, word);
    if (!(d & 0xfefefefeu)) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const uint32_t w0 = W[i >> 1][2 * (i & 1)], w1 = W[i >> 1][2 * (i & 1) + 1];
            if (kk[i].y >= 0) record(a, tile, (uint32_t)kk[i].y)[word] = make_uint2(w0 & vm, w1 & vm);
            if (kk[i].x >= 0) record(a, tile, (uint32_t)kk[i].x)[word] = make_uint2(~w0 & vm, ~w1 & vm);
        }
        return;
    }
    // some byte >= 2 (missing or multi-allelic call): build the invalid plane from a second
    // read of my rows (L1/L2 hits), one row at a time
    uint32_t Z[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int k = 0; k < 4; ++k) Z[j][k] = 0u;
    uint32_t fl = 0u;
    uint32_t refmask[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
        refmask[j] = ((kk[2 * j].x >= 0 || kk[2 * j].y >= 0) ? 0x0000ffffu : 0u) |
                     ((kk[2 * j + 1].x >= 0 || kk[2 * j + 1].y >= 0) ? 0xffff0000u : 0u);
    const uint8_t* p = base;
#pragma unroll 1
    for (int q = 0; q < 4; ++q) {
        uint32_t z0 = 0u, z1 = 0u, z2 = 0u, z3 = 0u;
#pragma unroll 1
        for (int rr = 0; rr < 8; ++rr, p += a.gs) {  // row rr enters at bit 7, ends at bit rr
            uint4 x = make_uint4(0u, 0u, 0u, 0u);
            if (8 * q + rr < nrows) x = load_row16(p, ncols);
            z0 = (z0 >> 1) | nz_bytes_msb(x.x & 0xfefefefeu);
            z1 = (z1 >> 1) | nz_bytes_msb(x.y & 0xfefefefeu);
            z2 = (z2 >> 1) | nz_bytes_msb(x.z & 0xfefefefeu);
            z3 = (z3 >> 1) | nz_bytes_msb(x.w & 0xfefefefeu);
            if (a.flags)
                fl |= qc_flags4(x.x & refmask[0]) | qc_flags4(x.y & refmask[1]) |
                      qc_flags4(x.z & refmask[2]) | qc_flags4(x.w & refmask[3]);
        }
        insert_bytes(Z[0], z0, q);
        insert_bytes(Z[1], z1, q);
        insert_bytes(Z[2], z2, q);
        insert_bytes(Z[3], z3, q);
    }
    if (a.flags && fl) {
        if (fl & ~*reinterpret_cast<volatile uint32_t*>(a.flags)) atomicOr(a.flags, fl);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const uint32_t w0 = W[i >> 1][2 * (i & 1)], w1 = W[i >> 1][2 * (i & 1) + 1];
        const uint32_t ok0 = ~Z[i >> 1][2 * (i & 1)] & vm, ok1 = ~Z[i >> 1][2 * (i & 1) + 1] & vm;
        if (kk[i].y >= 0) record(a, tile, (uint32_t)kk[i].y)[word] = make_uint2(w0 & ok0, w1 & ok1);
        if (kk[i].x >= 0) record(a, tile, (uint32_t)kk[i].x)[word] = make_uint2(~w0 & ok0, ~w1 & ok1);
    }
}

// -----------------------------------------------------------------------------------------
// sample-major kernel with local ancestry, same mapping as pack_sm2_kernel (lane = one word x
// 8 columns; 128-byte coalesced record stores).  One pass over the genotype rows builds the
// allele-1 plane, one pass over the ancestry rows the three low bit-planes of the label bytes
// (their 8-row block accumulators are parked in thread-private shared memory); every key
// (allele a <= 1, label l < 8) is then  alt^ma & L0^m0 & L1^m1 & L2^m2 & ~invalid,  where the
// invalid plane (GT byte >= 2 or label byte >= 8) is only built by lanes that saw such a byte.
// The keys of a warp's 16 columns are a contiguous range: their {first key, count} pairs and
// (allele, label) bytes are staged once per warp.  Variants with other keys -> generic kernel.
// -----------------------------------------------------------------------------------------
constexpr int kAncKeyStage = 128;  // staged (allele, label) pairs per warp
constexpr size_t kSm2AncSmem = sizeof(uint32_t) * 48 * kPackThreads + sizeof(int2) * kPackWarps * 16 +
                               sizeof(uint16_t) * kPackWarps * kAncKeyStage;

__global__ void __launch_bounds__(kPackThreads, 2)
pack_sm2_anc_kernel(const PackArgs a, uint32_t n_cblocks, int64_t cblock0, int64_t tile0) {
    extern __shared__ uint32_t s_dyn[];
    uint32_t* gl = s_dyn + threadIdx.x;  // [(q * 3 + bit) * 4 + j][thread]
    int2* s_kr = reinterpret_cast<int2*>(s_dyn + 48 * kPackThreads);
    uint16_t* s_ki = reinterpret_cast<uint16_t*>(s_kr + kPackWarps * 16);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t tile = tile0 + blockIdx.x / n_cblocks;
    const int word = (warp >> 2) * 16 + (lane >> 1);
    const int64_t wcol0 = (cblock0 + blockIdx.x % n_cblocks) * kSm2CtaCols + (warp & 3) * 16;
    if (wcol0 >= a.n_cols) return;  // warp-uniform
    const int64_t col0 = wcol0 + (lane & 1) * 8;
    const int ncols = (int)max((int64_t)0, min((int64_t)8, a.n_cols - col0));
    const int64_t s0 = tile * kTileSamples + 32 * word;
    const int nrows = (int)max((int64_t)0, min((int64_t)32, a.n - s0));
    const bool have_rows = ncols > 0 && nrows > 0;
    const bool full = nrows == 32 && ncols == 8;
    int2* my_kr = s_kr + warp * 16;
    uint16_t* my_ki = s_ki + warp * kAncKeyStage;

    // (1) key ranges of the warp's 16 columns (lanes 0..15 fetch one each)
    int2 kr = make_int2(0, 0);
    if (lane < 16 && wcol0 + lane < a.n_cols) kr = __ldg(reinterpret_cast<const int2*>(a.col_key_range) + wcol0 + lane);

    // (2) genotype rows -> allele-1 plane words W[j][k] (byte column 4j + k)
    uint32_t W[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int k = 0; k < 4; ++k) W[j][k] = 0u;
    uint32_t d = 0u;
    const uint8_t* gbase = a.gt + s0 * a.gs + col0 * 2;
    const uint8_t* lbase = a.anc + s0 * a.as + col0 * 2;
    if (have_rows) {
        const uint8_t* p = gbase;
#pragma unroll 1
        for (int q = 0; q < 4; ++q) {
            uint4 x[8];
            if (full) {
#pragma unroll
                for (int rr = 0; rr < 8; ++rr, p += a.gs) x[rr] = __ldg(reinterpret_cast<const uint4*>(p));
            } else {
#pragma unroll
                for (int rr = 0; rr < 8; ++rr, p += a.gs) {
                    x[rr] = make_uint4(0u, 0u, 0u, 0u);
                    if (8 * q + rr < nrows) x[rr] = load_row16(p, ncols);
                }
            }
            uint32_t a0 = 0u, a1 = 0u, a2 = 0u, a3 = 0u;
#pragma unroll
            for (int rr = 7; rr >= 0; --rr) {
                a0 = (a0 << 1) + (x[rr].x & 0x01010101u);
                a1 = (a1 << 1) + (x[rr].y & 0x01010101u);
                a2 = (a2 << 1) + (x[rr].z & 0x01010101u);
                a3 = (a3 << 1) + (x[rr].w & 0x01010101u);
                d |= ((x[rr].x | x[rr].y) | (x[rr].z | x[rr].w)) & 0xfefefefeu;
            }
            insert_bytes(W[0], a0, q);
            insert_bytes(W[1], a1, q);
            insert_bytes(W[2], a2, q);
            insert_bytes(W[3], a3, q);
        }
    }

    // (3) publish the key ranges, find the warp's key span, fetch its (allele, label) bytes
    if (lane < 16) my_kr[lane] = kr;
    int kmin = kr.y > 0 ? kr.x : 0x7fffffff, kmax = kr.y > 0 ? kr.x + kr.y : 0;
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
        kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
    }
    if (kmax <= kmin) return;  // none of the warp's columns is referenced (warp-uniform)
    const int n_k = kmax - kmin;
    const bool ki_staged = n_k <= kAncKeyStage;
    uint32_t ki_reg[kAncKeyStage / 32];
#pragma unroll
    for (int t = 0; t < kAncKeyStage / 32; ++t) {
        const int idx = lane + 32 * t;
        ki_reg[t] = 0u;
        if (ki_staged && idx < n_k)
            ki_reg[t] = (uint32_t)__ldg(a.key_allele + kmin + idx) | ((uint32_t)__ldg(a.key_label + kmin + idx) << 8);
    }

    // (4) ancestry rows -> block accumulators of label bits 0..2, parked in shared memory
    if (have_rows) {
        const uint8_t* p = lbase;
#pragma unroll 1
        for (int q = 0; q < 4; ++q) {
            uint4 c[8];
            if (full) {
#pragma unroll
                for (int rr = 0; rr < 8; ++rr, p += a.as) c[rr] = __ldg(reinterpret_cast<const uint4*>(p));
            } else {
#pragma unroll
                for (int rr = 0; rr < 8; ++rr, p += a.as) {
                    c[rr] = make_uint4(0u, 0u, 0u, 0u);
                    if (8 * q + rr < nrows) c[rr] = load_row16(p, ncols);
                }
            }
#pragma unroll
            for (int bit = 0; bit < 3; ++bit) {
                uint32_t a0 = 0u, a1 = 0u, a2 = 0u, a3 = 0u;
#pragma unroll
                for (int rr = 7; rr >= 0; --rr) {
                    a0 = (a0 << 1) + ((c[rr].x >> bit) & 0x01010101u);
                    a1 = (a1 << 1) + ((c[rr].y >> bit) & 0x01010101u);
                    a2 = (a2 << 1) + ((c[rr].z >> bit) & 0x01010101u);
                    a3 = (a3 << 1) + ((c[rr].w >> bit) & 0x01010101u);
                }
                uint32_t* dst = gl + ((q * 3 + bit) * 4) * kPackThreads;
                dst[0] = a0;
                dst[kPackThreads] = a1;
                dst[2 * kPackThreads] = a2;
                dst[3 * kPackThreads] = a3;
            }
#pragma unroll
            for (int rr = 0; rr < 8; ++rr) d |= ((c[rr].x | c[rr].y) | (c[rr].z | c[rr].w)) & 0xf8f8f8f8u;
        }
    } else {
#pragma unroll 1
        for (int i = 0; i < 48; ++i) gl[i * kPackThreads] = 0u;
    }

    // (5) publish the key bytes
    if (ki_staged) {
#pragma unroll
        for (int t = 0; t < kAncKeyStage / 32; ++t)
            if (lane + 32 * t < n_k) my_ki[lane + 32 * t] = (uint16_t)ki_reg[t];
    }
    __syncwarp();

    // (6) invalid plane, only for lanes that saw a GT byte >= 2 or a label byte >= 8
    uint32_t Z[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int k = 0; k < 4; ++k) Z[j][k] = 0u;
    if (d) {
        uint32_t fl = 0u, refmask[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c_lo = (lane & 1) * 8 + 2 * j;
            refmask[j] = (my_kr[c_lo].y > 0 ? 0x0000ffffu : 0u) | (my_kr[c_lo + 1].y > 0 ? 0xffff0000u : 0u);
        }
        const uint8_t* pg = gbase;
        const uint8_t* pl = lbase;
#pragma unroll 1
        for (int q = 0; q < 4; ++q) {
            uint32_t z0 = 0u, z1 = 0u, z2 = 0u, z3 = 0u;
#pragma unroll 1
            for (int rr = 0; rr < 8; ++rr, pg += a.gs, pl += a.as) {
                uint4 x = make_uint4(0u, 0u, 0u, 0u), c = make_uint4(0u, 0u, 0u, 0u);
                if (8 * q + rr < nrows) {
                    x = load_row16(pg, ncols);
                    c = load_row16(pl, ncols);
                }
                z0 = (z0 >> 1) | nz_bytes_msb((x.x & 0xfefefefeu) | (c.x & 0xf8f8f8f8u));
                z1 = (z1 >> 1) | nz_bytes_msb((x.y & 0xfefefefeu) | (c.y & 0xf8f8f8f8u));
                z2 = (z2 >> 1) | nz_bytes_msb((x.z & 0xfefefefeu) | (c.z & 0xf8f8f8f8u));
                z3 = (z3 >> 1) | nz_bytes_msb((x.w & 0xfefefefeu) | (c.w & 0xf8f8f8f8u));
                if (a.flags)
                    fl |= qc_flags4(x.x & refmask[0]) | qc_flags4(x.y & refmask[1]) |
                          qc_flags4(x.z & refmask[2]) | qc_flags4(x.w & refmask[3]);
            }
            insert_bytes(Z[0], z0, q);
            insert_bytes(Z[1], z1, q);
            insert_bytes(Z[2], z2, q);
            insert_bytes(Z[3], z3, q);
        }
        if (a.flags && fl) {
            if (fl & ~*reinterpret_cast<volatile uint32_t*>(a.flags)) atomicOr(a.flags, fl);
        }
    }

    // (7) every key of my 8 columns: combine the planes, store word `word` of both strands
    if (ncols <= 0) return;
    const uint32_t vm = valid_mask(a.n, tile, word);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint32_t L[3][4];
#pragma unroll
        for (int bit = 0; bit < 3; ++bit) {
            uint32_t G[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) G[q] = gl[((q * 3 + bit) * 4 + j) * kPackThreads];
            byte_transpose4(G, L[bit]);
        }
#pragma unroll
        for (int cp = 0; cp < 2; ++cp) {
            const int2 r = my_kr[(lane & 1) * 8 + 2 * j + cp];
#pragma unroll 1
            for (int t = 0; t < r.y; ++t) {
                const int key = r.x + t;
                const uint32_t info = ki_staged ? (uint32_t)my_ki[key - kmin]
                                                : ((uint32_t)__ldg(a.key_allele + key) | ((uint32_t)__ldg(a.key_label + key) << 8));
                const uint32_t mA = (info & 1u) ? 0u : 0xffffffffu, m0 = (info & 0x100u) ? 0u : 0xffffffffu;
                const uint32_t m1 = (info & 0x200u) ? 0u : 0xffffffffu, m2 = (info & 0x400u) ? 0u : 0xffffffffu;
                uint2 o;
                o.x = (W[j][2 * cp] ^ mA) & (L[0][2 * cp] ^ m0) & (L[1][2 * cp] ^ m1) & (L[2][2 * cp] ^ m2) &
                      ~Z[j][2 * cp] & vm;
                o.y = (W[j][2 * cp + 1] ^ mA) & (L[0][2 * cp + 1] ^ m0) & (L[1][2 * cp + 1] ^ m1) &
                      (L[2][2 * cp + 1] ^ m2) & ~Z[j][2 * cp + 1] & vm;
                record(a, tile, (uint32_t)key)[word] = o;
            }
        }
    }
}

// -----------------------------------------------------------------------------------------
// Line-consuming form of pack_sm2_anc_kernel: warp = 4 words x the CTA's 64 columns, so that one load
// instruction reads 4 rows x 128 contiguous bytes (whole cache lines: no line waits in L2 for the other
// three warps that own a sector of it); lane = one word x 8 columns as before.  The price: a key's
// record is stored 4 words (32 bytes, one sector) per store instruction instead of 16.  All warps of a
// CTA share the same 64 columns: key ranges and (allele, label) bytes are staged once per CTA.  The
// allele-1 accumulators are parked in shared memory like the label bits and the key loop is rolled
// (93 registers).  Default form for ancestry plans: 18.0 ms against 19.7 ms on the ancestry config
// (0.74 of the measured HBM peak); 3 CTAs per SM (80 registers) are SLOWER (22.9 ms), see DESIGN.md.
// -----------------------------------------------------------------------------------------
constexpr int kAncLineKeyStage = 2048;  // staged (allele, label) pairs per CTA (its 64 or 512 columns)
constexpr size_t sm2_anc_line_smem(bool slab) {
    return sizeof(uint32_t) * 64 * kPackThreads + sizeof(int2) * (slab ? kSm2SlabCols : kSm2CtaCols) +
           sizeof(uint16_t) * kAncLineKeyStage + sizeof(int) * 2;
}

#ifndef HT_ANC_LINE_MIN_CTAS
#define HT_ANC_LINE_MIN_CTAS 2
#endif
// kSlab: CTA = 512 columns, walking the 8 slabs (4 words = 128 rows each) of its tile one after the other
// with the key tables staged once (n_cblocks counts 512-column blocks); else CTA = 32 words x 64 columns.
template <bool kSlab>
__global__ void __launch_bounds__(kPackThreads, HT_ANC_LINE_MIN_CTAS)
pack_sm2_anc_line_kernel(const PackArgs a, uint32_t n_cblocks, int64_t cblock0, int64_t tile0) {
    constexpr int kCols = kSlab ? kSm2SlabCols : kSm2CtaCols;  // columns of the CTA
    extern __shared__ uint32_t s_dyn[];
    uint32_t* gl = s_dyn + threadIdx.x;  // [(q * 4 + plane) * 4 + j][thread]; plane 0 = allele 1, 1..3 = label bits
    int2* s_kr = reinterpret_cast<int2*>(s_dyn + 64 * kPackThreads);  // key range of each of the CTA's columns
    uint16_t* s_ki = reinterpret_cast<uint16_t*>(s_kr + kCols);
    int* s_span = reinterpret_cast<int*>(s_ki + kAncLineKeyStage);     // {first key, end key} of the CTA

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t tile = tile0 + blockIdx.x / n_cblocks;
    const int64_t wcol0 = (cblock0 + blockIdx.x % n_cblocks) * kCols;  // the CTA's first column
    if (wcol0 >= a.n_cols) return;  // CTA-uniform
    const int ccol = (kSlab ? warp * 64 : 0) + (lane & 7) * 8;  // my first column inside the CTA
    const int64_t col0 = wcol0 + ccol;
    const int ncols = (int)max((int64_t)0, min((int64_t)8, a.n_cols - col0));

    // (1) key ranges of the CTA's columns, its key span, the (allele, label) bytes of those keys
    if (threadIdx.x == 0) {
        s_span[0] = 0x7fffffff;
        s_span[1] = 0;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < kCols; c += kPackThreads) {
        int2 kr = make_int2(0, 0);
        if (wcol0 + c < a.n_cols) kr = __ldg(reinterpret_cast<const int2*>(a.col_key_range) + wcol0 + c);
        s_kr[c] = kr;
        if (kr.y > 0) {
            atomicMin(&s_span[0], kr.x);
            atomicMax(&s_span[1], kr.x + kr.y);
        }
    }
    __syncthreads();
    const int kmin = s_span[0], kmax = s_span[1];
    if (kmax <= kmin) return;  // none of the CTA's columns is referenced (CTA-uniform)
    const int n_k = kmax - kmin;
    const bool ki_staged = n_k <= kAncLineKeyStage;
    if (ki_staged) {
        for (int idx = threadIdx.x; idx < n_k; idx += kPackThreads)
            s_ki[idx] = (uint16_t)((uint32_t)__ldg(a.key_allele + kmin + idx) | ((uint32_t)__ldg(a.key_label + kmin + idx) << 8));
    }
    __syncthreads();
    if (ncols <= 0) return;  // no barrier below

    uint32_t fl = 0u;
#pragma unroll 1
    for (int part = 0; part < (kSlab ? kSm2SlabsPerTile : 1); ++part) {
        const int word = kSlab ? part * 4 + (lane >> 3) : warp * 4 + (lane >> 3);
        const int64_t s0 = tile * kTileSamples + 32 * word;
        const int nrows = (int)max((int64_t)0, min((int64_t)32, a.n - s0));
        const bool full = nrows == 32 && ncols == 8;
        const uint8_t* gbase = a.gt + s0 * a.gs + col0 * 2;
        const uint8_t* lbase = a.anc + s0 * a.as + col0 * 2;
        uint32_t d = 0u;

        // (2) genotype rows -> 8-row block accumulators of the allele-1 plane, parked in thread-private
        //     shared memory like those of the label bits below: registers are the load buffers
        if (nrows > 0) {
            const uint8_t* p = gbase;
#pragma unroll 1
            for (int q = 0; q < 4; ++q) {
                uint4 x[8];
                if (full) {
#pragma unroll
                    for (int rr = 0; rr < 8; ++rr, p += a.gs) x[rr] = __ldg(reinterpret_cast<const uint4*>(p));
                } else {
#pragma unroll
                    for (int rr = 0; rr < 8; ++rr, p += a.gs) {
                        x[rr] = make_uint4(0u, 0u, 0u, 0u);
                        if (8 * q + rr < nrows) x[rr] = load_row16(p, ncols);
                    }
                }
                uint32_t a0 = 0u, a1 = 0u, a2 = 0u, a3 = 0u;
#pragma unroll
                for (int rr = 7; rr >= 0; --rr) {  // row rr ends at bit rr of its byte
                    a0 = (a0 << 1) + (x[rr].x & 0x01010101u);
                    a1 = (a1 << 1) + (x[rr].y & 0x01010101u);
                    a2 = (a2 << 1) + (x[rr].z & 0x01010101u);
                    a3 = (a3 << 1) + (x[rr].w & 0x01010101u);
                    d |= ((x[rr].x | x[rr].y) | (x[rr].z | x[rr].w)) & 0xfefefefeu;
                }
                uint32_t* dst = gl + (q * 16) * kPackThreads;
                dst[0] = a0;
                dst[kPackThreads] = a1;
                dst[2 * kPackThreads] = a2;
                dst[3 * kPackThreads] = a3;
            }
            // (3) ancestry rows -> block accumulators of label bits 0..2
            p = lbase;
#pragma unroll 1
            for (int q = 0; q < 4; ++q) {
                uint4 c[8];
                if (full) {
#pragma unroll
                    for (int rr = 0; rr < 8; ++rr, p += a.as) c[rr] = __ldg(reinterpret_cast<const uint4*>(p));
                } else {
#pragma unroll
                    for (int rr = 0; rr < 8; ++rr, p += a.as) {
                        c[rr] = make_uint4(0u, 0u, 0u, 0u);
                        if (8 * q + rr < nrows) c[rr] = load_row16(p, ncols);
                    }
                }
#pragma unroll
                for (int bit = 0; bit < 3; ++bit) {
                    uint32_t a0 = 0u, a1 = 0u, a2 = 0u, a3 = 0u;
#pragma unroll
                    for (int rr = 7; rr >= 0; --rr) {
                        a0 = (a0 << 1) + ((c[rr].x >> bit) & 0x01010101u);
                        a1 = (a1 << 1) + ((c[rr].y >> bit) & 0x01010101u);
                        a2 = (a2 << 1) + ((c[rr].z >> bit) & 0x01010101u);
                        a3 = (a3 << 1) + ((c[rr].w >> bit) & 0x01010101u);
                    }
                    uint32_t* dst = gl + ((q * 4 + bit + 1) * 4) * kPackThreads;
                    dst[0] = a0;
                    dst[kPackThreads] = a1;
                    dst[2 * kPackThreads] = a2;
                    dst[3 * kPackThreads] = a3;
                }
#pragma unroll
                for (int rr = 0; rr < 8; ++rr) d |= ((c[rr].x | c[rr].y) | (c[rr].z | c[rr].w)) & 0xf8f8f8f8u;
            }
        } else {
#pragma unroll 1
            for (int i = 0; i < 64; ++i) gl[i * kPackThreads] = 0u;
        }

        // (4) every key of my 8 columns, one pair of columns (= one 4-byte word of my rows) at a time:
        //     combine the planes, store word `word` of both strands.  The loop is rolled and nothing but the
        //     shared-memory accumulators is carried into it.
        const uint32_t vm = valid_mask(a.n, tile, word);
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {
            uint32_t Wj[4], L[3][4], Zj[4] = {0u, 0u, 0u, 0u};
#pragma unroll
            for (int plane = 0; plane < 4; ++plane) {
                uint32_t G[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) G[q] = gl[((q * 4 + plane) * 4 + j) * kPackThreads];
                if (plane == 0) byte_transpose4(G, Wj);
                else byte_transpose4(G, L[plane - 1]);
            }
            if (d) {
                // invalid plane (GT byte >= 2 or label byte >= 8), only for lanes that saw such a byte: a
                // second read of my rows' word j (L1/L2 hits), one row at a time
                const int c_lo = ccol + 2 * j;
                const uint32_t refmask = (s_kr[c_lo].y > 0 ? 0x0000ffffu : 0u) | (s_kr[c_lo + 1].y > 0 ? 0xffff0000u : 0u);
                const int nc = ncols - 2 * j;  // columns of this word that exist (may be <= 0)
                const uint8_t* pg = gbase + 4 * j;
                const uint8_t* pl = lbase + 4 * j;
#pragma unroll 1
                for (int q = 0; q < 4; ++q) {
                    uint32_t z = 0u;
#pragma unroll 1
                    for (int rr = 0; rr < 8; ++rr, pg += a.gs, pl += a.as) {  // row rr enters at bit 7, ends at bit rr
                        uint32_t x = 0u, c = 0u;
                        if (8 * q + rr < nrows && nc > 0) {
                            x = load_cols4(pg, nc);
                            c = load_cols4(pl, nc);
                        }
                        z = (z >> 1) | nz_bytes_msb((x & 0xfefefefeu) | (c & 0xf8f8f8f8u));
                        if (a.flags) fl |= qc_flags4(x & refmask);
                    }
                    insert_bytes(Zj, z, q);
                }
            }
#pragma unroll
            for (int cp = 0; cp < 2; ++cp) {
                const int2 r = s_kr[ccol + 2 * j + cp];
#pragma unroll 1
                for (int t = 0; t < r.y; ++t) {
                    const int key = r.x + t;
                    const uint32_t info = ki_staged ? (uint32_t)s_ki[key - kmin]
                                                    : ((uint32_t)__ldg(a.key_allele + key) | ((uint32_t)__ldg(a.key_label + key) << 8));
                    const uint32_t mA = (info & 1u) ? 0u : 0xffffffffu, m0 = (info & 0x100u) ? 0u : 0xffffffffu;
                    const uint32_t m1 = (info & 0x200u) ? 0u : 0xffffffffu, m2 = (info & 0x400u) ? 0u : 0xffffffffu;
                    uint2 o;
                    o.x = (Wj[2 * cp] ^ mA) & (L[0][2 * cp] ^ m0) & (L[1][2 * cp] ^ m1) & (L[2][2 * cp] ^ m2) &
                          ~Zj[2 * cp] & vm;
                    o.y = (Wj[2 * cp + 1] ^ mA) & (L[0][2 * cp + 1] ^ m0) & (L[1][2 * cp + 1] ^ m1) &
                          (L[2][2 * cp + 1] ^ m2) & ~Zj[2 * cp + 1] & vm;
                    record(a, tile, (uint32_t)key)[word] = o;
                }
            }
        }
    }
    if (a.flags && fl) {
        if (fl & ~*reinterpret_cast<volatile uint32_t*>(a.flags)) atomicOr(a.flags, fl);
    }
}

static bool aligned(const void* p, int64_t stride, int to) {
    return ((reinterpret_cast<uintptr_t>(p) | (uintptr_t)stride) & (uintptr_t)(to - 1)) == 0;
}

// Launches `kernel` over all tiles; blockIdx.x = tile-major (tile, x-block).  The kernels take
// (PackArgs, uint32 n_xblocks, [int64 extra,] int64 tile0).
template <class K>
static int launch_tiled(K kernel, const PackArgs& a, int64_t n_xblocks, int64_t extra,
                        bool has_extra, cudaStream_t st, size_t smem = 0) {
    const int64_t n_tiles = ht_num_tiles(a.n);
    if (n_xblocks <= 0 || n_tiles <= 0) return HT_OK;
    if (n_xblocks > 0x7fffffff) {
        set_error("pack: too many variant blocks (%lld)", (long long)n_xblocks);
        return HT_ERR_UNSUPPORTED;
    }
    uint32_t nxb = (uint32_t)n_xblocks;
    if (smem > 48 * 1024)
        HT_CUDA_TRY(cudaFuncSetAttribute((const void*)kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // a grid holds at most 2^31-1 blocks: launch ranges of tiles
    const int64_t tiles_per_launch = max((int64_t)1, (int64_t)0x7fffffff / n_xblocks);
    for (int64_t t0 = 0; t0 < n_tiles; t0 += tiles_per_launch) {
        const int64_t nt = min(tiles_per_launch, n_tiles - t0);
        void* args_extra[] = {(void*)&a, (void*)&nxb, (void*)&extra, (void*)&t0};
        void* args_plain[] = {(void*)&a, (void*)&nxb, (void*)&t0};
        HT_CUDA_TRY(cudaLaunchKernel((const void*)kernel, dim3((unsigned)(nt * n_xblocks)),
                                     dim3(kPackThreads), has_extra ? args_extra : args_plain, smem, st));
    }
    return HT_OK;
}

// variant-major kernel over all (tile, variant) items
static int launch_pack_vm(const PackArgs& a, bool anc, cudaStream_t st) {
    const int64_t n_tiles = ht_num_tiles(a.n);
    const int64_t want = (n_tiles * a.n_vars + kPackWarps - 1) / kPackWarps;
    if (want <= 0) return HT_OK;
    // Grid: whole waves of resident CTAs (4 per SM), about 48 items per warp -- 2 waves on the
    // 1000G-scale config (103.6 us against 107.1 at 8 and 126.9 at 32), 32 waves at 200 k samples x
    // 50 k variants (3.69 ms against 3.75 at 8); profiles/exp_pack_vm.py.  HT_VM_WAVES / HT_VM_ORDER
    // override the choice (measurements, tests).
    const char* waves_env = getenv("HT_VM_WAVES");
    const int env_waves = waves_env ? atoi(waves_env) : 0;
    const int64_t n_items = n_tiles * a.n_vars, wave_warps = (int64_t)148 * 4 * kPackWarps;
    const int64_t waves = env_waves > 0 ? env_waves : max((int64_t)2, min((int64_t)32, (n_items + 24 * wave_warps) / (48 * wave_warps)));
    const unsigned grid = (unsigned)min(want, (int64_t)148 * 4 * waves);
    // Item order: variant-first (neighbouring warps read neighbouring pieces of one variant row) is
    // faster at every size measured: 107.1 against 114.8 us (1000G-scale), 3.75 against 3.83 ms (200 k)
    const char* order = getenv("HT_VM_ORDER");  // "t": tile-first
    const bool vfirst = order ? order[0] != 't' : true;
    if (anc) pack_vm_kernel<true><<<grid, kPackThreads, 0, st>>>(a, n_tiles, vfirst);
    else pack_vm_kernel<false><<<grid, kPackThreads, 0, st>>>(a, n_tiles, vfirst);
    HT_CUDA_TRY(cudaGetLastError());
    return HT_OK;
}

// -----------------------------------------------------------------------------------------
// pgenlib int32 allele codes -> bytes, referenced rows only: row i of `dst` = the 2n codes of variant
// var_col[i] as the reference stores them in its uint8 matrix (haptools/data/genotypes.py:1388-1401: -9 -> 255,
// everything else modulo 256), i.e. a variant-major byte matrix the variant-major pack kernel reads.  Pure
// streaming: a lane turns one 16-byte load (2 samples x 2 strands) into one 4-byte store.
// -----------------------------------------------------------------------------------------
constexpr int kNarrowPerThread = 4;  // 16-byte loads per thread

__device__ __forceinline__ uint32_t narrow4(int4 v) {
    const uint32_t b0 = v.x == -9 ? 255u : (uint32_t)v.x, b1 = v.y == -9 ? 255u : (uint32_t)v.y;
    const uint32_t b2 = v.z == -9 ? 255u : (uint32_t)v.z, b3 = v.w == -9 ? 255u : (uint32_t)v.w;
    return prmt(prmt(b0, b1, 0x0040), prmt(b2, b3, 0x0040), 0x5410);  // low bytes of the four
}

__global__ void __launch_bounds__(256)
narrow_i32_rows_kernel(const uint8_t* __restrict__ src, int64_t src_pitch, const uint32_t* __restrict__ var_col,
                       int64_t n_values, uint32_t blocks_per_row, uint8_t* __restrict__ dst, int64_t dst_pitch,
                       uint32_t* __restrict__ ident) {
    const uint32_t row = blockIdx.x / blocks_per_row, xb = blockIdx.x % blocks_per_row;
    if (xb == 0 && threadIdx.x == 0) ident[row] = row;
    const int4* s = reinterpret_cast<const int4*>(src + (int64_t)__ldg(var_col + row) * src_pitch);
    uint32_t* d = reinterpret_cast<uint32_t*>(dst + (int64_t)row * dst_pitch);
    const int64_t n4 = n_values >> 2;
    const int64_t i0 = (int64_t)xb * (256 * kNarrowPerThread) + threadIdx.x;
    int4 v[kNarrowPerThread];
#pragma unroll
    for (int u = 0; u < kNarrowPerThread; ++u)
        if (i0 + 256 * u < n4) v[u] = __ldg(s + i0 + 256 * u);
#pragma unroll
    for (int u = 0; u < kNarrowPerThread; ++u)
        if (i0 + 256 * u < n4) d[i0 + 256 * u] = narrow4(v[u]);
    if (xb == blocks_per_row - 1 && threadIdx.x < (n_values & 3)) {  // 2n values: a tail of 2 when n is odd
        const int32_t x = reinterpret_cast<const int32_t*>(s)[4 * n4 + threadIdx.x];
        reinterpret_cast<uint8_t*>(d)[4 * n4 + threadIdx.x] = x == -9 ? (uint8_t)255 : (uint8_t)x;
    }
}

static int check_pack(const PackArgs& a) {
    if (a.n < 0 || a.n_vars < 0 || a.n_keys < 0 || a.key_base < 0) return bad_arg("negative size");
    if (a.n_vars == 0 || a.n == 0) return HT_OK;
    if (!a.gt) return bad_arg("gt is NULL");
    if (!a.var_col || !a.var_key_off || !a.key_allele) return bad_arg("variant/key tables are NULL");
    if (!a.planes) return bad_arg("planes is NULL");
    if (reinterpret_cast<uintptr_t>(a.planes) & 15) return bad_arg("planes must be 16-byte aligned");
    if ((a.anc != nullptr) != (a.key_label != nullptr))
        return bad_arg("anc and key_label must be given together");
    return HT_OK;
}

}  // namespace ht

extern "C" {

int ht_pack_u8(const uint8_t* gt, int64_t gt_s_samp, int64_t gt_s_var, int64_t gt_s_strand,
               const uint8_t* anc, int64_t anc_s_samp, int64_t anc_s_var, int64_t anc_s_strand,
               int64_t n_samples, const ht_pack_tables* t, uint32_t* planes, uint32_t* flags,
               int mode, void* stream) {
    using namespace ht;
    if (!t) return bad_arg("tables is NULL");
    PackArgs a;
    a.gt = gt; a.gs = gt_s_samp; a.gv = gt_s_var; a.gc = gt_s_strand;
    a.anc = anc; a.as = anc_s_samp; a.av = anc_s_var; a.ac = anc_s_strand;
    a.n = n_samples;
    a.var_col = t->var_col; a.var_key_off = t->var_key_off;
    a.key_allele = t->key_allele; a.key_label = t->key_label;
    a.col_var = t->col_var; a.col_key01 = t->col_key01; a.col_key_range = t->col_key_range; a.var_sel = nullptr;
    a.n_cols = t->n_cols;
    a.n_vars = t->n_vars; a.n_keys = t->n_keys; a.key_base = 0;
    a.planes = planes; a.flags = flags;
    if (int rc = check_pack(a)) return rc;
    if (a.n_vars == 0 || n_samples == 0) return HT_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n_cols = t->n_cols;

    const bool vm_ok = gt_s_strand == 1 && gt_s_samp == 2 && aligned(gt, gt_s_var, 16) &&
                       (!anc || (anc_s_strand == 1 && anc_s_samp == 2 && aligned(anc, anc_s_var, 16)));
    const bool sm_ok = gt_s_strand == 1 && gt_s_var == 2 && aligned(gt, gt_s_samp, 4) && t->col_var &&
                       n_cols > 0 &&
                       (!anc || (anc_s_strand == 1 && anc_s_var == 2 && aligned(anc, anc_s_samp, 4)));
    const bool sm2_ok = gt_s_strand == 1 && gt_s_var == 2 && aligned(gt, gt_s_samp, 16) && n_cols > 0 &&
                        t->n_var_other >= 0 && (t->n_var_other == 0 || t->var_other) &&
                        (anc ? (anc_s_strand == 1 && anc_s_var == 2 && aligned(anc, anc_s_samp, 16) &&
                                t->col_key_range && aligned(t->col_key_range, 0, 8))
                             : (t->col_key01 && aligned(t->col_key01, 0, 8)));
    if (mode == HT_PACK_AUTO)
        mode = vm_ok ? HT_PACK_VARIANT_MAJOR
                     : (sm2_ok ? HT_PACK_SAMPLE_MAJOR_BIALLELIC
                               : (sm_ok ? HT_PACK_SAMPLE_MAJOR : HT_PACK_GENERIC));

    const int64_t n_vblocks = (a.n_vars + kPackWarps - 1) / kPackWarps;
    switch (mode) {
        case HT_PACK_GENERIC:
            return anc ? launch_tiled(pack_generic_kernel<LoadU8, true>, a, n_vblocks, 0, false, st)
                       : launch_tiled(pack_generic_kernel<LoadU8, false>, a, n_vblocks, 0, false, st);
        case HT_PACK_VARIANT_MAJOR:
            if (!vm_ok) {
                set_error("variant-major pack needs s_strand=1, s_samp=2 and 16-byte aligned rows");
                return HT_ERR_UNSUPPORTED;
            }
            return launch_pack_vm(a, anc != nullptr, st);
        case HT_PACK_SAMPLE_MAJOR: {
            if (!sm_ok) {
                set_error("sample-major pack needs s_strand=1, s_var=2, 4-byte aligned rows and col_var");
                return HT_ERR_UNSUPPORTED;
            }
            const int64_t n_cblocks = (n_cols + kSmCtaCols - 1) / kSmCtaCols;
            return anc ? launch_tiled(pack_sm_kernel<true>, a, n_cblocks, 0, true, st,
                                      sizeof(uint32_t) * kSmStageAnc * kPackThreads)
                       : launch_tiled(pack_sm_kernel<false>, a, n_cblocks, 0, true, st,
                                      sizeof(uint32_t) * kSmStagePlain * kPackThreads);
        }
        case HT_PACK_SAMPLE_MAJOR_BIALLELIC: {
            if (!sm2_ok) {
                set_error("biallelic sample-major pack needs s_strand=1, s_var=2, 16-byte aligned rows, "
                          "var_other and col_key01 (col_key_range with ancestry)");
                return HT_ERR_UNSUPPORTED;
            }
            const int64_t n_cblocks = (n_cols + kSm2CtaCols - 1) / kSm2CtaCols;
            // long rows (see pack_sm2_kernel<true>); HT_PACK_SM2_KERNEL=slab|tile overrides (measurements, tests)
            const char* sm2_form = getenv("HT_PACK_SM2_KERNEL");
            const bool sm2_slab = sm2_form ? sm2_form[0] == 's' : gt_s_samp >= kSm2SlabMinPitch;
            const int64_t n_slab_blocks = kSm2SlabsPerTile * ((n_cols + kSm2SlabCols - 1) / kSm2SlabCols);
            const char* anc_form = getenv("HT_PACK_ANC_KERNEL");  // "warp" = pack_sm2_anc_kernel (A/B measurements, tests)
            const bool anc_line = !(anc_form && anc_form[0] == 'w');
            // default with ancestry: the TMA-fed kernel (ht_pack_anc_tma.cu); "line" / "warp" select the register-load
            // forms, which also serve drivers that cannot encode tensor maps
            const bool anc_tma_forced = anc_form && anc_form[0] == 't';
            if (anc && !(sm2_form && sm2_form[0]) && (anc_tma_forced || (!anc_form && pack_anc_tma_available()))) {
                const int rc = launch_pack_sm2_anc_tma(a, st);
                if (rc != HT_ERR_UNSUPPORTED || anc_tma_forced) {
                    if (rc) return rc;
                    if (t->n_var_other > 0) {  // the few variants with other keys: generic kernel
                        PackArgs o = a;
                        o.var_sel = t->var_other;
                        o.n_vars = t->n_var_other;
                        return launch_tiled(pack_generic_kernel<LoadU8, true>, o, (o.n_vars + kPackWarps - 1) / kPackWarps, 0, false, st);
                    }
                    return HT_OK;
                }
            }
            // the ancestry kernel keeps the tile mapping: its line-consuming loads make it insensitive to the row
            // pitch (4.7 TB/s from 100 KB to 2 MB) and the slab mapping repeats its key staging 8 x per tile
            // (18.2 vs 21.2 ms on the ancestry config, profiles/exp_pack_pitch3.py); slab only on request
            const bool anc_slab = sm2_form && sm2_form[0] == 's';
            if (int rc = anc ? (anc_line ? (anc_slab ? launch_tiled(pack_sm2_anc_line_kernel<true>, a, n_slab_blocks / kSm2SlabsPerTile, 0, true, st, sm2_anc_line_smem(true))
                                                     : launch_tiled(pack_sm2_anc_line_kernel<false>, a, n_cblocks, 0, true, st, sm2_anc_line_smem(false)))
                                         : launch_tiled(pack_sm2_anc_kernel, a, n_cblocks, 0, true, st, kSm2AncSmem))
                             : (sm2_slab ? launch_tiled(pack_sm2_kernel<true>, a, n_slab_blocks, 0, true, st)
                                         : launch_tiled(pack_sm2_kernel<false>, a, n_cblocks, 0, true, st)))
                return rc;
            if (t->n_var_other > 0) {  // the few variants with other keys: generic kernel
                PackArgs o = a;
                o.var_sel = t->var_other;
                o.n_vars = t->n_var_other;
                const int64_t nvb = (o.n_vars + kPackWarps - 1) / kPackWarps;
                return anc ? launch_tiled(pack_generic_kernel<LoadU8, true>, o, nvb, 0, false, st)
                           : launch_tiled(pack_generic_kernel<LoadU8, false>, o, nvb, 0, false, st);
            }
            return HT_OK;
        }
        default:
            return bad_arg("unknown pack mode");
    }
}

int ht_pack_i32(const int32_t* alleles, int64_t row_pitch_elems, int64_t n_samples,
                const uint32_t* var_col, const uint32_t* var_key_off, const uint8_t* key_allele,
                int64_t n_vars, int64_t key_base, int64_t n_keys_total, uint32_t* planes,
                uint32_t* flags, void* stream) {
    return ht_pack_i32_anc(alleles, row_pitch_elems, n_samples, nullptr, 0, 0, 0, var_col, var_key_off, key_allele,
                           nullptr, n_vars, key_base, n_keys_total, planes, flags, stream);
}

int ht_pack_i32_anc(const int32_t* alleles, int64_t row_pitch_elems, int64_t n_samples, const uint8_t* anc,
                    int64_t anc_s_samp, int64_t anc_s_var, int64_t anc_s_strand, const uint32_t* var_col,
                    const uint32_t* var_key_off, const uint8_t* key_allele, const uint8_t* key_label,
                    int64_t n_vars, int64_t key_base, int64_t n_keys_total, uint32_t* planes,
                    uint32_t* flags, void* stream) {
    using namespace ht;
    PackArgs a;
    a.gt = reinterpret_cast<const uint8_t*>(alleles);
    a.gs = 8; a.gv = row_pitch_elems * 4; a.gc = 4;
    a.anc = anc; a.as = anc_s_samp; a.av = anc_s_var; a.ac = anc_s_strand;
    a.n = n_samples;
    a.var_col = var_col; a.var_key_off = var_key_off;
    a.key_allele = key_allele; a.key_label = key_label;
    a.col_var = nullptr; a.col_key01 = nullptr; a.col_key_range = nullptr; a.var_sel = nullptr; a.n_cols = 0;
    a.n_vars = n_vars; a.n_keys = n_keys_total; a.key_base = key_base;
    a.planes = planes; a.flags = flags;
    if (int rc = check_pack(a)) return rc;
    if (n_vars == 0 || n_samples == 0) return HT_OK;
    if (row_pitch_elems < 2 * n_samples) return bad_arg("row_pitch_elems < 2 * n_samples");
    if (!aligned(alleles, a.gv, 8)) return bad_arg("int32 allele rows must be 8-byte aligned");
    // Without ancestry labels: narrow the referenced rows to bytes (what the reference's uint8 matrix holds) and run
    // the variant-major pack kernel on them -- 2 streaming kernels, 4 x the rate of the generic ballot kernel on the
    // streamed-PGEN config.  HT_I32_KERNEL=generic keeps the one-kernel form (measurements, tests).
    const char* form = getenv("HT_I32_KERNEL");
    if (!anc && !(form && form[0] == 'g') && aligned(alleles, a.gv, 16) && n_vars < ((int64_t)1 << 31)) {
        cudaStream_t st = (cudaStream_t)stream;
        const int64_t n_values = 2 * n_samples, dst_pitch = (n_values + 15) / 16 * 16;
        const int64_t blocks_per_row = (n_values / 4 + 256 * kNarrowPerThread - 1) / (256 * kNarrowPerThread);
        if (blocks_per_row > 0 && blocks_per_row * n_vars < ((int64_t)1 << 31)) {
            uint8_t* scratch = nullptr;
            const size_t bytes = (size_t)n_vars * (size_t)dst_pitch;
            HT_CUDA_TRY(cudaMallocAsync((void**)&scratch, bytes + sizeof(uint32_t) * (size_t)n_vars, st));
            uint32_t* ident = reinterpret_cast<uint32_t*>(scratch + bytes);  // bytes is a multiple of 16
            narrow_i32_rows_kernel<<<(unsigned)(blocks_per_row * n_vars), 256, 0, st>>>(
                a.gt, a.gv, var_col, n_values, (uint32_t)blocks_per_row, scratch, dst_pitch, ident);
            cudaError_t e = cudaGetLastError();
            int rc = HT_OK;
            if (e == cudaSuccess) {
                PackArgs b = a;
                b.gt = scratch; b.gs = 2; b.gv = dst_pitch; b.gc = 1;
                b.var_col = ident;
                rc = launch_pack_vm(b, false, st);
            }
            const cudaError_t ef = cudaFreeAsync(scratch, st);
            if (e != cudaSuccess) return cuda_fail(e, "narrow_i32_rows_kernel");
            if (rc) return rc;
            HT_CUDA_TRY(ef);
            return HT_OK;
        }
    }
    const int64_t n_vblocks = (n_vars + kPackWarps - 1) / kPackWarps;
    return anc ? launch_tiled(pack_generic_kernel<LoadI32, true>, a, n_vblocks, 0, false, (cudaStream_t)stream)
               : launch_tiled(pack_generic_kernel<LoadI32, false>, a, n_vblocks, 0, false, (cudaStream_t)stream);
}

}  // extern "C"
