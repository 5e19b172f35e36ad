// Bit-packed result -> haplotype-major allele matrix for the PGEN writer (sm_100a).
//
// GenotypesPLINK.write (haptools/data/genotypes.py:1569-1659) transposes the (n, H, 2) result to
// (H, n, 2), widens a chunk of haplotype rows to int32 and flattens it to (rows, 2n) for
// pgenlib's PgenWriter.append_alleles_batch (:1618-1642).  The transform's bit-packed result
// (tiles, H, 32, 2) already is haplotype-major per tile, so that hand-off is a plain expansion:
// row r of `out` = haplotype h0 + r, element 2 s + c = presence on strand c of sample s, as
// uint8 (1 byte per call over PCIe, the host widens a contiguous row) or int32 (what pgenlib
// takes as is).
//
// warp = one (haplotype, tile): lane w loads word w of both strands (one coalesced 256-byte
// record), then the warp writes the record's 2048 elements in 512-byte fully coalesced store
// instructions (the source word comes over a shuffle).
#include "ht_common.cuh"

namespace ht {

constexpr int kHmWarps = 8;

// bits 0..7 of a and b interleaved: bit 2i = a_i, bit 2i + 1 = b_i
__device__ __forceinline__ uint32_t interleave8(uint32_t a, uint32_t b) {
    uint32_t t = (a & 0xffu) | ((b & 0xffu) << 16);
    t = (t | (t << 4)) & 0x0f0f0f0fu;
    t = (t | (t << 2)) & 0x33333333u;
    t = (t | (t << 1)) & 0x55555555u;
    return (t & 0xffffu) | ((t >> 16) << 1);
}

template <typename T>
__global__ void __launch_bounds__(kHmWarps * 32)
hapmajor_kernel(const uint32_t* __restrict__ bits, int64_t n_samples, int64_t n_haps, int64_t h0,
                int64_t n_rows, int64_t n_tiles, T* __restrict__ out, int64_t pitch) {
    const int lane = threadIdx.x & 31;
    const int64_t item = (int64_t)blockIdx.x * kHmWarps + (threadIdx.x >> 5);
    if (item >= n_rows * n_tiles) return;
    // tile fastest: the warps of a CTA write consecutive pieces of one output row
    const int64_t r = item / n_tiles, tile = item % n_tiles;
    const uint2 w = __ldg(reinterpret_cast<const uint2*>(bits) + ((size_t)tile * (size_t)n_haps + (size_t)(h0 + r)) * kTileWords + lane);
    const int64_t s_tile = tile * kTileSamples;
    T* row = out + (size_t)r * (size_t)pitch + 2 * s_tile;
    const int64_t valid = min((int64_t)kTileSamples, n_samples - s_tile);  // samples of this tile
    if (sizeof(T) == 4) {
        // store i: elements [128 i, 128 i + 128) = samples [64 i, 64 i + 64) = words 2 i, 2 i + 1;
        // lane l writes samples 64 i + 2 l, + 1 (4 elements)
#pragma unroll 4
        for (int i = 0; i < 16; ++i) {
            const int src = 2 * i + (lane >> 4), b = (2 * lane) & 31;
            const uint32_t x = __shfl_sync(0xffffffffu, w.x, src) >> b, y = __shfl_sync(0xffffffffu, w.y, src) >> b;
            const int s = 64 * i + 2 * lane;
            int4 v = make_int4(x & 1, y & 1, (x >> 1) & 1, (y >> 1) & 1);
            T* p = row + 2 * s;
            if (s + 1 < valid) {
                *reinterpret_cast<int4*>(p) = v;
            } else if (s < valid) {
                reinterpret_cast<int2*>(p)[0] = make_int2(v.x, v.y);
            }
        }
    } else {
        // store i: bytes [512 i, 512 i + 512) = samples [256 i, 256 i + 256) = words 8 i .. 8 i + 7;
        // lane l writes samples 256 i + 8 l .. + 7 (16 bytes)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int src = 8 * i + (lane >> 2), b = 8 * (lane & 3);
            const uint32_t x = __shfl_sync(0xffffffffu, w.x, src) >> b, y = __shfl_sync(0xffffffffu, w.y, src) >> b;
            const uint32_t z = interleave8(x, y);  // 16 bits: element order (sample, strand)
            const int s = 256 * i + 8 * lane;
            uint4 v;
            v.x = spread4(z & 15u);
            v.y = spread4((z >> 4) & 15u);
            v.z = spread4((z >> 8) & 15u);
            v.w = spread4((z >> 12) & 15u);
            uint8_t* p = reinterpret_cast<uint8_t*>(row) + 2 * s;
            if (s + 8 <= valid && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
                *reinterpret_cast<uint4*>(p) = v;
            } else {
                const uint32_t q[4] = {v.x, v.y, v.z, v.w};
                for (int e = 0; e < 16; ++e)
                    if (s + (e >> 1) < valid) p[e] = (uint8_t)((q[e >> 2] >> (8 * (e & 3))) & 0xffu);
            }
        }
    }
}

}  // namespace ht

extern "C" int ht_unpack_bits_hapmajor(const uint32_t* bits, int64_t n_samples, int64_t n_haps, int64_t h0,
                                       int64_t n_rows, int elem_size, void* out, int64_t out_pitch_elems,
                                       void* stream) {
    using namespace ht;
    if (n_samples < 0 || n_haps < 0 || h0 < 0 || n_rows < 0 || h0 + n_rows > n_haps) return bad_arg("bad haplotype range");
    if (elem_size != 1 && elem_size != 4) return bad_arg("elem_size must be 1 (uint8) or 4 (int32)");
    if (n_rows == 0 || n_samples == 0) return HT_OK;
    if (!bits || !out) return bad_arg("bits / out is NULL");
    if (out_pitch_elems < 2 * n_samples) return bad_arg("out_pitch_elems < 2 * n_samples");
    if (elem_size == 4 && ((reinterpret_cast<uintptr_t>(out) & 15) || (out_pitch_elems & 3)))
        return bad_arg("int32 output must be 16-byte aligned with a row pitch that is a multiple of 4 elements");
    const int64_t n_tiles = ht_num_tiles(n_samples);
    const int64_t blocks = (n_rows * n_tiles + kHmWarps - 1) / kHmWarps;
    if (blocks > 0x7fffffff) {
        set_error("too many (haplotype, tile) items (%lld): unpack fewer rows per call", (long long)(n_rows * n_tiles));
        return HT_ERR_UNSUPPORTED;
    }
    if (elem_size == 4)
        hapmajor_kernel<int32_t><<<(unsigned)blocks, kHmWarps * 32, 0, (cudaStream_t)stream>>>(
            bits, n_samples, n_haps, h0, n_rows, n_tiles, reinterpret_cast<int32_t*>(out), out_pitch_elems);
    else
        hapmajor_kernel<uint8_t><<<(unsigned)blocks, kHmWarps * 32, 0, (cudaStream_t)stream>>>(
            bits, n_samples, n_haps, h0, n_rows, n_tiles, reinterpret_cast<uint8_t*>(out), out_pitch_elems);
    HT_CUDA_TRY(cudaGetLastError());
    return HT_OK;
}

// ---------------------------------------------------------------------------------------
// VCF text of the GT columns.  GenotypesVCF.write (haptools/data/genotypes.py:752-813) sets
// record.samples[s]["GT"] sample by sample in Python and lets htslib print "a|b" per sample,
// tab-separated, newline-terminated.  For the transform's result (phased, 0/1, never missing)
// that text is fixed-width: 4 bytes per sample, "a|b\t" ("a|b\n" for the last sample), so a
// haplotype row is n * 4 bytes and is produced straight from the bit-packed result.
// warp = one (haplotype, tile); store i covers samples [128 i, 128 i + 128) = 512 contiguous
// bytes, lane l the 4 samples 128 i + 4 l .. + 3 (one 16-byte store).
// ---------------------------------------------------------------------------------------
namespace ht {

__global__ void __launch_bounds__(kHmWarps * 32)
vcf_gt_text_kernel(const uint32_t* __restrict__ bits, int64_t n_samples, int64_t n_haps, int64_t h0,
                   int64_t n_rows, int64_t n_tiles, uint8_t* __restrict__ out, int64_t pitch) {
    const int lane = threadIdx.x & 31;
    const int64_t item = (int64_t)blockIdx.x * kHmWarps + (threadIdx.x >> 5);
    if (item >= n_rows * n_tiles) return;
    const int64_t r = item / n_tiles, tile = item % n_tiles;
    const uint2 w = __ldg(reinterpret_cast<const uint2*>(bits) + ((size_t)tile * (size_t)n_haps + (size_t)(h0 + r)) * kTileWords + lane);
    const int64_t s_tile = tile * kTileSamples;
    uint8_t* row = out + (size_t)r * (size_t)pitch + 4 * s_tile;
    const int64_t valid = min((int64_t)kTileSamples, n_samples - s_tile);
    const int64_t last = n_samples - 1 - s_tile;  // tile-local index of the cohort's last sample
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int src = 4 * i + (lane >> 3), b = 4 * (lane & 7);
        const uint32_t x = __shfl_sync(0xffffffffu, w.x, src) >> b, y = __shfl_sync(0xffffffffu, w.y, src) >> b;
        const int s = 128 * i + 4 * lane;
        uint32_t t[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            // '0' + a, '|', '0' + b, '\t'  (little endian)
            t[e] = 0x09307c30u + ((x >> e) & 1u) + (((y >> e) & 1u) << 16);
            if (s + e == last) t[e] = (t[e] & 0x00ffffffu) | 0x0a000000u;  // '\n' ends the row
        }
        uint8_t* p = row + 4 * s;
        if (s + 4 <= valid && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
            *reinterpret_cast<uint4*>(p) = make_uint4(t[0], t[1], t[2], t[3]);
        } else {
            for (int e = 0; e < 4; ++e)
                if (s + e < valid)
                    for (int k = 0; k < 4; ++k) p[4 * e + k] = (uint8_t)(t[e] >> (8 * k));
        }
    }
}

}  // namespace ht

extern "C" int ht_format_vcf_gt(const uint32_t* bits, int64_t n_samples, int64_t n_haps, int64_t h0,
                                int64_t n_rows, uint8_t* out, int64_t out_pitch, void* stream) {
    using namespace ht;
    if (n_samples < 0 || n_haps < 0 || h0 < 0 || n_rows < 0 || h0 + n_rows > n_haps) return bad_arg("bad haplotype range");
    if (n_rows == 0 || n_samples == 0) return HT_OK;
    if (!bits || !out) return bad_arg("bits / out is NULL");
    if (out_pitch < 4 * n_samples) return bad_arg("out_pitch < 4 * n_samples");
    const int64_t n_tiles = ht_num_tiles(n_samples);
    const int64_t blocks = (n_rows * n_tiles + kHmWarps - 1) / kHmWarps;
    if (blocks > 0x7fffffff) {
        set_error("too many (haplotype, tile) items (%lld): format fewer rows per call", (long long)(n_rows * n_tiles));
        return HT_ERR_UNSUPPORTED;
    }
    vcf_gt_text_kernel<<<(unsigned)blocks, kHmWarps * 32, 0, (cudaStream_t)stream>>>(bits, n_samples, n_haps, h0, n_rows,
                                                                                   n_tiles, out, out_pitch);
    HT_CUDA_TRY(cudaGetLastError());
    return HT_OK;
}
