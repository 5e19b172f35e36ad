// Deterministic synthetic genotypes / local-ancestry labels for the benchmark configs.
// Not part of the reference: BASELINE.json's large configs do not fit a host, so they are
// generated on the device from a counter-based hash that haptools_b200/synth.py restates in
// numpy (any sample slice can be regenerated on the host and pushed through the oracle).
#include "ht_common.cuh"

namespace ht {

struct SynthArgs {
    uint8_t* out;
    int64_t ss, sv, sc;  // byte strides: sample, variant, strand
    int64_t n, p;
    int64_t sample_offset;
    uint64_t seed;
    int mode;
    uint32_t p0, p1;  // gt: missing_per_64k, -   ancestry: n_pops, tract_len
};

// genotype of (global sample s, variant v): both strands
__device__ __forceinline__ uint32_t synth_gt2(const SynthArgs& a, uint64_t s, uint64_t v) {
    const uint64_t hs = splitmix64(a.seed + s);
    const uint64_t r = splitmix64(hs ^ v);
    uint32_t g0, g1;
    if (a.mode == 0) {
        g0 = (uint32_t)(r & 1u);
        g1 = (uint32_t)((r >> 1) & 1u);
    } else {
        // "blocky": inside each 64-variant block every (sample, strand) copies one of 16
        // founder haplotypes, so multi-variant haplotypes have non-trivial frequencies
        const uint64_t rb = splitmix64(hs ^ ((1ull << 40) | (v >> 6)));
        const uint64_t f0 = rb & 15u, f1 = (rb >> 4) & 15u;
        g0 = (uint32_t)(splitmix64(splitmix64(a.seed ^ 0xF00Dull ^ (f0 << 48)) ^ v) & 1u);
        g1 = (uint32_t)(splitmix64(splitmix64(a.seed ^ 0xF00Dull ^ (f1 << 48)) ^ v) & 1u);
    }
    if (a.p0) {
        if (((r >> 16) & 0xffffu) < a.p0) g0 = 255u;
        if (((r >> 32) & 0xffffu) < a.p0) g1 = 255u;
    }
    return g0 | (g1 << 8);
}

// local-ancestry label: piecewise constant along variants, tracts of p1 variants with a
// per-(sample, strand) phase, label uniform in [0, p0)
__device__ __forceinline__ uint32_t synth_anc2(const SynthArgs& a, uint64_t s, uint64_t v) {
    const uint64_t hs = splitmix64(a.seed + s);
    uint32_t out = 0;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        const uint64_t off = splitmix64(hs ^ ((2ull << 40) | (uint64_t)c));
        const uint64_t tract = (v + (off & 0xffffu)) / a.p1;
        const uint32_t lab = (uint32_t)(splitmix64((off >> 16) + tract) % a.p0);
        out |= lab << (8 * c);
    }
    return out;
}

template <bool kAncestry>
__global__ void __launch_bounds__(256) synth_kernel(const SynthArgs a, bool variant_fast) {
    const int64_t total = a.n * a.p;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        // consecutive threads walk the axis with the smaller stride
        const int64_t s = variant_fast ? i / a.p : i % a.n;
        const int64_t v = variant_fast ? i % a.p : i / a.n;
        const uint32_t g = kAncestry ? synth_anc2(a, (uint64_t)(s + a.sample_offset), (uint64_t)v)
                                     : synth_gt2(a, (uint64_t)(s + a.sample_offset), (uint64_t)v);
        uint8_t* o = a.out + s * a.ss + v * a.sv;
        if (a.sc == 1 && ((reinterpret_cast<uintptr_t>(o) & 1) == 0)) {
            *reinterpret_cast<uint16_t*>(o) = (uint16_t)g;
        } else {
            o[0] = (uint8_t)(g & 0xffu);
            o[a.sc] = (uint8_t)(g >> 8);
        }
    }
}

// dst(s, v, c) = src(s, v, c) for c in {0, 1}: re-layout of a strided genotype chunk (e.g. the
// reference's arrays that still interleave the phase column) into the dense
// 2-bytes-per-genotype form the fast pack kernels read.  Only used on host-fed chunks, where
// PCIe, not HBM, is the bound.
struct RelayoutArgs {
    const uint8_t* src;
    int64_t ss, sv, sc;
    uint8_t* dst;
    int64_t ds, dv;
    int64_t n, p;
};

__global__ void __launch_bounds__(256) relayout_kernel(const RelayoutArgs a, bool variant_fast) {
    const int64_t total = a.n * a.p;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = variant_fast ? i / a.p : i % a.n;
        const int64_t v = variant_fast ? i % a.p : i / a.n;
        const uint8_t* in = a.src + s * a.ss + v * a.sv;
        const uint16_t g = (uint16_t)in[0] | ((uint16_t)in[a.sc] << 8);
        *reinterpret_cast<uint16_t*>(a.dst + s * a.ds + v * a.dv) = g;
    }
}

static int launch_synth(SynthArgs a, bool ancestry, void* stream) {
    if (a.n < 0 || a.p < 0) return bad_arg("negative size");
    if (a.n == 0 || a.p == 0) return HT_OK;
    if (!a.out) return bad_arg("output is NULL");
    const bool variant_fast = a.sv <= a.ss;
    const int64_t total = a.n * a.p;
    const int64_t blocks = min((total + 255) / 256, (int64_t)148 * 64);
    if (ancestry)
        synth_kernel<true><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a, variant_fast);
    else
        synth_kernel<false><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a, variant_fast);
    HT_CUDA_TRY(cudaGetLastError());
    return HT_OK;
}

}  // namespace ht

extern "C" {

int ht_synth_gt(uint8_t* gt, int64_t s_samp, int64_t s_var, int64_t s_strand, int64_t n_samples,
                int64_t n_vars, int64_t sample_offset, uint64_t seed, int mode,
                uint32_t missing_per_64k, void* stream) {
    using namespace ht;
    if (mode != 0 && mode != 1) return bad_arg("ht_synth_gt: mode must be 0 or 1");
    SynthArgs a{gt, s_samp, s_var, s_strand, n_samples, n_vars, sample_offset, seed, mode,
                missing_per_64k, 0u};
    return launch_synth(a, false, stream);
}

int ht_relayout_u8(const uint8_t* src, int64_t s_samp, int64_t s_var, int64_t s_strand, int64_t n_samples,
                   int64_t n_vars, uint8_t* dst, int64_t d_samp, int64_t d_var, void* stream) {
    using namespace ht;
    if (n_samples < 0 || n_vars < 0) return bad_arg("negative size");
    if (n_samples == 0 || n_vars == 0) return HT_OK;
    if (!src || !dst) return bad_arg("ht_relayout_u8: NULL pointer");
    if (((reinterpret_cast<uintptr_t>(dst) | (uintptr_t)d_samp | (uintptr_t)d_var) & 1) != 0)
        return bad_arg("ht_relayout_u8: destination must be 2-byte aligned with even strides");
    RelayoutArgs a{src, s_samp, s_var, s_strand, dst, d_samp, d_var, n_samples, n_vars};
    const int64_t total = n_samples * n_vars;
    const int64_t blocks = min((total + 255) / 256, (int64_t)148 * 64);
    relayout_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a, d_var <= d_samp);
    HT_CUDA_TRY(cudaGetLastError());
    return HT_OK;
}

int ht_synth_ancestry(uint8_t* anc, int64_t s_samp, int64_t s_var, int64_t s_strand,
                      int64_t n_samples, int64_t n_vars, int64_t sample_offset, uint64_t seed,
                      uint32_t n_pops, uint32_t tract_len, void* stream) {
    using namespace ht;
    if (n_pops == 0 || n_pops > 255 || tract_len == 0) return bad_arg("ht_synth_ancestry: bad n_pops/tract_len");
    SynthArgs a{anc, s_samp, s_var, s_strand, n_samples, n_vars, sample_offset, seed, 0, n_pops,
                tract_len};
    return launch_synth(a, true, stream);
}

}  // extern "C"
