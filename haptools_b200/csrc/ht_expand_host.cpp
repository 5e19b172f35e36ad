// Host side of the bit-packed result hand-off: bits -> the 0x00 / 0x01 bytes of numpy's bool, written
// straight into the caller's (n, H, 2) array (haptools/data/haplotypes.py:1408 allocates it with np.empty).
//
// This is data movement, not the transform: every result BIT was computed on the GPU (ht_transform_bits +
// ht_bits_to_rowbits); the host only widens a bit to the byte the reference's dtype asks for, because 1 bit
// instead of 1 byte per hap-call is what crosses PCIe.  Plain C++ (compiled by g++ through nvcc) so that the
// x86 vector forms can be selected at run time: AVX-512BW (mask -> bytes in one instruction), AVX2, or a
// portable 64-bit multiply.  Streaming (non-temporal) stores whenever the destination allows: the result is
// written once and read by somebody else later.
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#if defined(__x86_64__) || defined(__i386__)
#include <immintrin.h>
#define HT_X86 1
#endif

namespace ht {

namespace {

inline uint64_t load_bits64(const uint8_t* bits, int64_t bit0, int64_t n_bits_total) {
    // 64 bits starting at bit `bit0` of the little-endian bit string; bits beyond n_bits_total read as 0 and
    // no byte beyond ceil(n_bits_total / 8) is touched
    const int64_t byte0 = bit0 >> 3;
    const int sh = (int)(bit0 & 7);
    const int64_t n_bytes = (n_bits_total + 7) >> 3;
    uint64_t lo = 0;
    uint8_t hi = 0;
    const int64_t avail = n_bytes - byte0;
    if (avail >= 9) {
        memcpy(&lo, bits + byte0, 8);
        hi = bits[byte0 + 8];
    } else if (avail > 0) {
        uint8_t tmp[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        memcpy(tmp, bits + byte0, (size_t)avail);
        memcpy(&lo, tmp, 8);
        hi = tmp[8];
    }
    return sh ? (lo >> sh) | ((uint64_t)hi << (64 - sh)) : lo;
}

inline uint64_t spread8(uint32_t b) {
    // 8 bits -> 8 bytes of 0/1 (byte i = bit i): the multiply puts a copy of b into every byte, the mask keeps
    // bit i in byte i, and adding 0x7f carries a set bit into bit 7 of its own byte (0x80 + 0x7f = 0xff: no
    // carry into the next byte)
    const uint64_t x = ((uint64_t)b * 0x0101010101010101ull) & 0x8040201008040201ull;
    return ((x + 0x7f7f7f7f7f7f7f7full) >> 7) & 0x0101010101010101ull;
}

void expand_scalar(const uint8_t* bits, int64_t bit0, int64_t n_bits_total, uint8_t* out, int64_t n) {
    int64_t j = 0;
    for (; j + 8 <= n; j += 8) {
        const uint64_t m = load_bits64(bits, bit0 + j, n_bits_total);
        const uint64_t v = spread8((uint32_t)(m & 0xff));
        memcpy(out + j, &v, 8);
    }
    if (j < n) {
        const uint64_t m = load_bits64(bits, bit0 + j, n_bits_total);
        for (int64_t i = 0; j + i < n; ++i) out[j + i] = (uint8_t)((m >> i) & 1);
    }
}

#ifdef HT_X86
__attribute__((target("avx512f,avx512bw"))) void expand_body_avx512(const uint8_t* bits, int64_t bit0, int64_t n_bits_total,
                                                                   uint8_t* out, int64_t n_blocks64, bool stream) {
    // out is 64-byte aligned here
    const __m512i one = _mm512_set1_epi8(1);
    const int sh = (int)(bit0 & 7);
    const uint8_t* p = bits + (bit0 >> 3);
    const int64_t n_bytes = (n_bits_total + 7) >> 3;
    const int64_t safe = (n_bytes - (bit0 >> 3) - 9) / 8;  // blocks whose 9-byte window lies inside the string
    int64_t i = 0;
    for (; i < n_blocks64 && i < safe; ++i) {
        uint64_t lo;
        memcpy(&lo, p + 8 * i, 8);
        const uint64_t m = sh ? (lo >> sh) | ((uint64_t)p[8 * i + 8] << (64 - sh)) : lo;
        const __m512i v = _mm512_maskz_mov_epi8((__mmask64)m, one);
        if (stream)
            _mm512_stream_si512((__m512i*)(out + 64 * i), v);
        else
            _mm512_storeu_si512((void*)(out + 64 * i), v);
    }
    for (; i < n_blocks64; ++i) {
        const uint64_t m = load_bits64(bits, bit0 + 64 * i, n_bits_total);
        const __m512i v = _mm512_maskz_mov_epi8((__mmask64)m, one);
        if (stream)
            _mm512_stream_si512((__m512i*)(out + 64 * i), v);
        else
            _mm512_storeu_si512((void*)(out + 64 * i), v);
    }
}

__attribute__((target("avx2"))) void expand_body_avx2(const uint8_t* bits, int64_t bit0, int64_t n_bits_total,
                                                       uint8_t* out, int64_t n_blocks64, bool stream) {
    // byte k of a 32-byte half takes bit k: broadcast the 32 bits, route byte k / 8 of them to byte k, test bit k % 8
    const __m256i route = _mm256_setr_epi8(0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 3, 3,
                                           3, 3, 3, 3, 3, 3);
    const __m256i bit = _mm256_set1_epi64x((long long)0x8040201008040201ull);
    const __m256i one = _mm256_set1_epi8(1);
    for (int64_t i = 0; i < n_blocks64; ++i) {
        const uint64_t m = load_bits64(bits, bit0 + 64 * i, n_bits_total);
        for (int h = 0; h < 2; ++h) {
            const __m256i b = _mm256_set1_epi32((int)(uint32_t)(m >> (32 * h)));
            const __m256i t = _mm256_and_si256(_mm256_shuffle_epi8(b, route), bit);
            const __m256i v = _mm256_and_si256(_mm256_cmpeq_epi8(t, bit), one);
            if (stream)
                _mm256_stream_si256((__m256i*)(out + 64 * i + 32 * h), v);
            else
                _mm256_storeu_si256((__m256i*)(out + 64 * i + 32 * h), v);
        }
    }
}

int detect_isa() {
    __builtin_cpu_init();
    if (__builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512f")) return 2;
    if (__builtin_cpu_supports("avx2")) return 1;
    return 0;
}
#else
int detect_isa() { return 0; }
#endif

int g_isa = -1, g_isa_forced = -1;

}  // namespace

// 0: portable, 1: AVX2, 2: AVX-512BW; `force` >= 0 caps the choice (tests run every form the CPU has)
int expand_isa(int force) {
    if (g_isa < 0) g_isa = detect_isa();
    if (force >= 0) g_isa_forced = force < g_isa ? force : g_isa;
    if (force == -2) g_isa_forced = -1;
    return g_isa_forced >= 0 ? g_isa_forced : g_isa;
}

// out[j] = bit j of the little-endian bit string `bits` (byte j / 8, bit j % 8), j < n_out; 0x00 or 0x01.
void expand_bits_row(const uint8_t* bits, uint8_t* out, int64_t n_out) {
    const int isa = expand_isa(-1);
    int64_t j = 0;
#ifdef HT_X86
    if (isa > 0 && n_out >= 256) {
        const int64_t head = (int64_t)((64 - ((uintptr_t)out & 63)) & 63);
        expand_scalar(bits, 0, n_out, out, head);
        const int64_t blocks = (n_out - head) / 64;
        if (isa == 2)
            expand_body_avx512(bits, head, n_out, out + head, blocks, true);
        else
            expand_body_avx2(bits, head, n_out, out + head, blocks, true);
        j = head + 64 * blocks;
    }
#endif
    expand_scalar(bits, j, n_out, out + j, n_out - j);
}

// rows of a pitched matrix; ends with a store fence so that the streaming stores are visible to whoever is
// told next that the rows are there
void expand_bits_rows(const uint8_t* bits, int64_t bits_pitch, uint8_t* out, int64_t out_pitch, int64_t n_out, int64_t rows) {
    for (int64_t r = 0; r < rows; ++r) expand_bits_row(bits + r * bits_pitch, out + r * out_pitch, n_out);
#ifdef HT_X86
    _mm_sfence();
#endif
}

}  // namespace ht
