// Local-ancestry labels from breakpoint blocks, on the device.
//
// Replaces Breakpoints.population_array / _find_blocks
// (haptools/data/breakpoints.py:225-330): for every (sample, strand) the reference runs
// np.searchsorted(block end positions, variant positions, side="left") per chromosome in a
// Python loop and gathers the block's population label into an (n, p, 2) matrix that
// transform_haps then hands to the ancestry transform (haptools/transform.py:644-660).
// Here one thread per (sample, variant) does the two binary searches; the result is written
// straight into the device array the ancestry pack kernel reads.
#include "ht_common.cuh"

namespace ht {

struct BpArgs {
    const unsigned long long* blk_key;  // (chrom index << 32) | block end position
    const uint8_t* blk_pop;
    const uint32_t* seg_off;            // [2n + 1], segment of (sample s, strand c) = 2s + c
    const unsigned long long* var_key;  // (chrom index << 32) | variant position
    uint8_t* anc;
    int64_t ss, sv, sc;
    int64_t n, p;
    unsigned long long* bad;  // min over failures of ((2s + c) << 32 | v), init ~0
};

__device__ __forceinline__ uint32_t lower_bound_u64(const unsigned long long* a, uint32_t lo, uint32_t hi,
                                                    unsigned long long key) {
    while (lo < hi) {  // first index with a[idx] >= key (np.searchsorted side="left")
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if (__ldg(a + mid) < key) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(256) bp_ancestry_kernel(const BpArgs a, bool variant_fast) {
    const int64_t total = a.n * a.p;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = variant_fast ? i / a.p : i % a.n;
        const int64_t v = variant_fast ? i % a.p : i / a.n;
        const unsigned long long key = __ldg(a.var_key + v);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const uint32_t lo = __ldg(a.seg_off + 2 * s + c), hi = __ldg(a.seg_off + 2 * s + c + 1);
            const uint32_t idx = lower_bound_u64(a.blk_key, lo, hi, key);
            uint8_t lab = 255;
            // the block must exist and lie on the variant's chromosome
            if (idx < hi && (__ldg(a.blk_key + idx) >> 32) == (key >> 32)) lab = __ldg(a.blk_pop + idx);
            else if (a.bad) atomicMin(a.bad, ((unsigned long long)(2 * s + c) << 32) | (unsigned long long)v);
            a.anc[s * a.ss + v * a.sv + c * a.sc] = lab;
        }
    }
}

}  // namespace ht

extern "C" int ht_bp_ancestry(const uint64_t* blk_key, const uint8_t* blk_pop, const uint32_t* seg_off,
                              const uint64_t* var_key, uint8_t* anc, int64_t s_samp, int64_t s_var,
                              int64_t s_strand, int64_t n_samples, int64_t n_vars, uint64_t* bad,
                              void* stream) {
    using namespace ht;
    if (n_samples < 0 || n_vars < 0) return bad_arg("negative size");
    if (n_samples == 0 || n_vars == 0) return HT_OK;
    if (!seg_off || !var_key || !anc) return bad_arg("ht_bp_ancestry: NULL pointer");
    if (n_samples * n_vars > 0 && n_vars > 0xffffffffLL) return bad_arg("too many variants");
    BpArgs a{reinterpret_cast<const unsigned long long*>(blk_key), blk_pop, seg_off,
             reinterpret_cast<const unsigned long long*>(var_key), anc, s_samp, s_var, s_strand,
             n_samples, n_vars, reinterpret_cast<unsigned long long*>(bad)};
    const int64_t total = n_samples * n_vars;
    const int64_t blocks = min((total + 255) / 256, (int64_t)148 * 32);
    bp_ancestry_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a, s_var <= s_samp);
    HT_CUDA_TRY(cudaGetLastError());
    return HT_OK;
}
