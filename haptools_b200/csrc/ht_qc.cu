// Genotype QC on the device: ONE pass over the (n, p, 2|3) uint8 matrix yields what the
// reference's four checks each derive with full numpy passes on the host before `transform` runs
// (haptools/transform.py:623-631 -> haptools/data/genotypes.py):
//
//   check_missing    :444-485   np.any(data[:, :, :2] >= 254, axis=2)          (GenotypesAncestry: == 255,
//                               haptools/transform.py:353-383) -> first (sample, variant) / samples to discard
//   check_biallelic  :487-528   np.any(data[:, :, :2] > 1, axis=2)             -> first pair / variants to discard /
//                               number of pairs (its log line counts pairs)
//   check_phase      :530-557   (bool(d0) ^ bool(d1)) & ~bool(phase)           -> first pair
//   check_maf        :559-625   data[:, :, :2].astype(bool).sum(axis=(0, 2))   -> per-variant count
//
// "first" is np.nonzero's order on an (n, p) array: the minimum of (sample << 32 | variant) -- an atomicMin,
// like ht_bp_ancestry's `bad` word.  The kernel reads every byte once (HBM-bound); the thread layout follows
// the contiguous axis: with variants contiguous (PGEN reader layout) a thread owns 4 variants and walks
// samples, with samples contiguous (VCF reader layout) it owns 4 samples and walks variants.
#include "ht_common.cuh"

namespace ht {

struct QcArgs {
    const uint8_t* gt;
    int64_t ss, sv, sc;  // byte strides (sample, variant, strand)
    int depth;           // 3: a phase byte at strand index 2
    int64_t n, p, sample_offset;
    uint32_t missing_min;
    ht_qc_summary* sum;
    uint8_t* samp_missing;
    uint8_t* var_multi;
    unsigned long long* var_alt;
};

constexpr int kQcOuter = 64;  // outer-axis items per CTA
constexpr unsigned long long kNone = ~0ull;

struct Cell {
    uint32_t g0, g1, ph;
};

// 4 consecutive inner items at once when their bytes are one aligned run of 4 * depth bytes
template <int kDepth>
__device__ __forceinline__ void load4(const uint8_t* p, Cell (&c)[4]) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(p);
    if (kDepth == 2) {
        const uint2 v = *reinterpret_cast<const uint2*>(w);
        c[0] = {v.x & 0xff, (v.x >> 8) & 0xff, 1u};
        c[1] = {(v.x >> 16) & 0xff, v.x >> 24, 1u};
        c[2] = {v.y & 0xff, (v.y >> 8) & 0xff, 1u};
        c[3] = {(v.y >> 16) & 0xff, v.y >> 24, 1u};
    } else {
        const uint32_t a = w[0], b = w[1], d = w[2];
        c[0] = {a & 0xff, (a >> 8) & 0xff, (a >> 16) & 0xff};
        c[1] = {a >> 24, b & 0xff, (b >> 8) & 0xff};
        c[2] = {(b >> 16) & 0xff, b >> 24, d & 0xff};
        c[3] = {(d >> 8) & 0xff, (d >> 16) & 0xff, d >> 24};
    }
}

// kVarInner: inner axis = variants (thread owns 4 variants, walks samples); else inner = samples.
// kVec: 0 = byte loads with any strides, 2 / 3 = load4<depth> (contiguous aligned cells)
template <bool kVarInner, int kVec>
__global__ void __launch_bounds__(256) qc_kernel(const QcArgs a) {
    const int64_t n_inner = kVarInner ? a.p : a.n, n_outer = kVarInner ? a.n : a.p;
    const int64_t s_inner = kVarInner ? a.sv : a.ss, s_outer = kVarInner ? a.ss : a.sv;
    const int64_t i0 = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4;
    const int64_t o0 = (int64_t)blockIdx.y * kQcOuter, o1 = min(n_outer, o0 + kQcOuter);
    const int n_mine = (int)max((int64_t)0, min((int64_t)4, n_inner - i0));
    unsigned long long first_missing = kNone, first_multi = kNone, first_unphased = kNone;
    uint32_t n_multi = 0, n_missing = 0;
    uint32_t in_alt[4] = {0, 0, 0, 0}, in_flag[4] = {0, 0, 0, 0};  // per inner item: alt count / multi (kVarInner) or missing flag
    const bool has_phase = a.depth >= 3;
#pragma unroll 1
    for (int64_t o = o0; o < o1; ++o) {
        Cell c[4];
        const uint8_t* base = a.gt + o * s_outer + i0 * s_inner;
        if (kVec != 0 && n_mine == 4) {
            load4<kVec == 0 ? 2 : kVec>(base, c);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                c[j] = {0u, 0u, 1u};
                if (j < n_mine) {
                    const uint8_t* q = base + j * s_inner;
                    c[j].g0 = q[0];
                    c[j].g1 = q[a.sc];
                    if (has_phase) c[j].ph = q[2 * a.sc];
                }
            }
        }
        uint32_t o_alt = 0, o_multi = 0, o_missing = 0;  // per outer item, over my inner items
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (j >= n_mine) continue;
            const int64_t s = kVarInner ? o : i0 + j, v = kVarInner ? i0 + j : o;
            const unsigned long long key = ((unsigned long long)(s + a.sample_offset) << 32) | (unsigned long long)v;
            const uint32_t g0 = c[j].g0, g1 = c[j].g1;
            const bool missing = g0 >= a.missing_min || g1 >= a.missing_min;
            const bool multi = g0 > 1 || g1 > 1;
            const bool unphased = has_phase && ((g0 != 0) != (g1 != 0)) && c[j].ph == 0;
            const uint32_t alt = (g0 != 0) + (g1 != 0);
            if (missing) {
                first_missing = min(first_missing, key);
                ++n_missing;
            }
            if (multi) {
                first_multi = min(first_multi, key);
                ++n_multi;
            }
            if (unphased) first_unphased = min(first_unphased, key);
            if (kVarInner) {
                in_alt[j] += alt;
                in_flag[j] |= multi;
                o_missing |= missing;
            } else {
                in_flag[j] |= missing;
                o_alt += alt;
                o_multi |= multi;
            }
        }
        if (kVarInner) {
            if (o_missing && a.samp_missing) a.samp_missing[o] = 1;  // idempotent byte store
        } else {
            // per-variant totals over the warp's 128 samples: one atomic per warp and variant
            const uint32_t alt_w = __reduce_add_sync(0xffffffffu, o_alt);
            const uint32_t multi_w = __reduce_or_sync(0xffffffffu, o_multi);
            if ((threadIdx.x & 31) == 0) {
                if (alt_w && a.var_alt) atomicAdd(a.var_alt + o, (unsigned long long)alt_w);
                if (multi_w && a.var_multi) a.var_multi[o] = 1;
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (j >= n_mine) continue;
        if (kVarInner) {
            if (in_alt[j] && a.var_alt) atomicAdd(a.var_alt + i0 + j, (unsigned long long)in_alt[j]);
            if (in_flag[j] && a.var_multi) a.var_multi[i0 + j] = 1;
        } else if (in_flag[j] && a.samp_missing) {
            a.samp_missing[i0 + j] = 1;
        }
    }
    if (first_missing != kNone) atomicMin(reinterpret_cast<unsigned long long*>(&a.sum->first_missing), first_missing);
    if (first_multi != kNone) atomicMin(reinterpret_cast<unsigned long long*>(&a.sum->first_multiallelic), first_multi);
    if (first_unphased != kNone) atomicMin(reinterpret_cast<unsigned long long*>(&a.sum->first_unphased), first_unphased);
    const uint32_t nm = __reduce_add_sync(0xffffffffu, n_multi), ns = __reduce_add_sync(0xffffffffu, n_missing);
    if ((threadIdx.x & 31) == 0) {
        if (nm) atomicAdd(reinterpret_cast<unsigned long long*>(&a.sum->n_multiallelic), (unsigned long long)nm);
        if (ns) atomicAdd(reinterpret_cast<unsigned long long*>(&a.sum->n_missing), (unsigned long long)ns);
    }
}

template <bool kVarInner>
static int launch_qc(const QcArgs& a, cudaStream_t st) {
    const int64_t n_inner = kVarInner ? a.p : a.n, n_outer = kVarInner ? a.n : a.p;
    const int64_t s_inner = kVarInner ? a.sv : a.ss, s_outer = kVarInner ? a.ss : a.sv;
    const int64_t gx = (n_inner + 1023) / 1024, gy = (n_outer + kQcOuter - 1) / kQcOuter;
    if (gx > 0x7fffffff) return bad_arg("ht_qc_u8: inner axis too long");
    // contiguous cells (strand stride 1, inner stride = depth) in 4-byte aligned rows: 4 cells per 8 / 12-byte load
    const bool vec = a.sc == 1 && s_inner == a.depth && (s_outer & 3) == 0 && (reinterpret_cast<uintptr_t>(a.gt) & 7) == 0 &&
                     (a.depth == 3 || (s_outer & 7) == 0);
    for (int64_t y0 = 0; y0 < gy; y0 += 65535) {
        const dim3 grid((unsigned)gx, (unsigned)min((int64_t)65535, gy - y0));
        QcArgs b = a;
        const int64_t skip = y0 * kQcOuter;  // outer items handled by earlier launches
        b.gt = a.gt + skip * s_outer;
        if (kVarInner) {
            b.n = a.n - skip;
            b.sample_offset = a.sample_offset + skip;
            if (b.samp_missing) b.samp_missing += skip;
        } else {
            // variant indices must stay absolute for the first_* keys: handled by offsetting the tables instead
            b.p = a.p - skip;
            if (b.var_multi) b.var_multi += skip;
            if (b.var_alt) b.var_alt += skip;
            if (skip) return bad_arg("ht_qc_u8: more than 65535 * 64 variants in a variant-major array; call in pieces");
        }
        if (vec && a.depth == 3)
            qc_kernel<kVarInner, 3><<<grid, 256, 0, st>>>(b);
        else if (vec)
            qc_kernel<kVarInner, 2><<<grid, 256, 0, st>>>(b);
        else
            qc_kernel<kVarInner, 0><<<grid, 256, 0, st>>>(b);
        HT_CUDA_TRY(cudaGetLastError());
    }
    return HT_OK;
}

}  // namespace ht

extern "C" int ht_qc_u8(const uint8_t* gt, int64_t s_samp, int64_t s_var, int64_t s_strand, int depth,
                        int64_t n_samples, int64_t n_vars, int64_t sample_offset, int missing_min,
                        ht_qc_summary* summary, uint8_t* samp_missing, uint8_t* var_multiallelic,
                        uint64_t* var_alt_count, void* stream) {
    using namespace ht;
    if (n_samples < 0 || n_vars < 0 || sample_offset < 0) return bad_arg("negative size");
    if (depth != 2 && depth != 3) return bad_arg("ht_qc_u8: depth must be 2 or 3");
    if (missing_min < 2 || missing_min > 255) return bad_arg("ht_qc_u8: missing_min must be in 2..255");
    if (!summary) return bad_arg("ht_qc_u8: summary is NULL");
    if (n_samples == 0 || n_vars == 0) return HT_OK;
    if (!gt) return bad_arg("ht_qc_u8: gt is NULL");
    if (n_samples + sample_offset > 0xffffffffLL || n_vars > 0xffffffffLL) return bad_arg("ht_qc_u8: index does not fit 32 bits");
    QcArgs a{gt, s_samp, s_var, s_strand, depth, n_samples, n_vars, sample_offset, (uint32_t)missing_min, summary,
             samp_missing, var_multiallelic, reinterpret_cast<unsigned long long*>(var_alt_count)};
    const int64_t av = s_var < 0 ? -s_var : s_var, as = s_samp < 0 ? -s_samp : s_samp;
    return av <= as ? launch_qc<true>(a, (cudaStream_t)stream) : launch_qc<false>(a, (cudaStream_t)stream);
}
