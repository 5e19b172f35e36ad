/*
 * A host written in plain C99 driving the hot path through include/haptools_cuda.h alone
 * (no Python, no torch): cudaMalloc'ed buffers, the planner tables built here, ht_pack_u8 +
 * ht_transform_u8 / ht_transform_bits + ht_unpack_bits_u8 + ht_count_bits, result compared byte
 * for byte with the definition of the transform
 *
 *     out[s, h, c] = AND over (column v, allele a) of haplotype h of ( gt[s, v, c] == a )
 *
 * (haptools/data/haplotypes.py:1393-1412) evaluated by the loops in expect() below.  TEST
 * INFRASTRUCTURE: tests/test_c_host.py compiles it with gcc against libhaptools_cuda.so (CPU
 * suite: compile + link only) and runs it on the GPU box (`-m gpu`).
 *
 *     abi_roundtrip [n_samples [n_cols [n_haps [seed]]]]     exit code 0 = every check passed
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cuda_runtime_api.h>

#include "haptools_cuda.h"

#define MAX_ALLELE 3 /* allele indices 0..2: index 2 exercises the var_other (generic) route */

static uint64_t rng_state;
static uint32_t rnd(void) { /* splitmix64, high half */
    uint64_t z = (rng_state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return (uint32_t)((z ^ (z >> 31)) >> 32);
}

#define CU(call)                                                                          \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess) {                                                          \
            fprintf(stderr, "%s:%d: %s\n", __FILE__, __LINE__, cudaGetErrorString(e_));   \
            return 2;                                                                     \
        }                                                                                 \
    } while (0)
#define HT(call)                                                                          \
    do {                                                                                  \
        int rc_ = (call);                                                                 \
        if (rc_ != HT_OK) {                                                               \
            fprintf(stderr, "%s:%d: rc %d: %s\n", __FILE__, __LINE__, rc_, ht_last_error()); \
            return 3;                                                                     \
        }                                                                                 \
    } while (0)

static void* to_device(const void* host, size_t bytes) {
    void* d = NULL;
    if (cudaMalloc(&d, bytes ? bytes : 16) != cudaSuccess) return NULL;
    if (bytes && cudaMemcpy(d, host, bytes, cudaMemcpyHostToDevice) != cudaSuccess) return NULL;
    return d;
}

/* the definition, one byte at a time */
static void expect(const uint8_t* gt, int64_t pitch, int64_t n, const int64_t* hap_off, const int32_t* ref_col,
                   const uint8_t* ref_allele, int64_t H, uint8_t* out) {
    for (int64_t s = 0; s < n; ++s)
        for (int64_t h = 0; h < H; ++h)
            for (int c = 0; c < 2; ++c) {
                uint8_t all = 1;
                for (int64_t r = hap_off[h]; r < hap_off[h + 1]; ++r)
                    all &= (uint8_t)(gt[s * pitch + 2 * ref_col[r] + c] == ref_allele[r]);
                out[(s * H + h) * 2 + c] = all;
            }
}

int main(int argc, char** argv) {
    const int64_t n = argc > 1 ? atoll(argv[1]) : 2504;
    const int64_t p = argc > 2 ? atoll(argv[2]) : 96;
    const int64_t H = argc > 3 ? atoll(argv[3]) : 300;
    rng_state = argc > 4 ? strtoull(argv[4], NULL, 10) : 1;
    if (ht_version() != HT_ABI_VERSION) {
        fprintf(stderr, "library ABI %d, header %d\n", ht_version(), HT_ABI_VERSION);
        return 1;
    }

    /* genotypes: sample-major, rows padded to 16 bytes; mostly 0/1, some 2 (multi-allelic), some 255 (missing) */
    const int64_t pitch = (2 * p + 15) / 16 * 16;
    uint8_t* gt = (uint8_t*)calloc((size_t)(n * pitch), 1);
    for (int64_t s = 0; s < n; ++s)
        for (int64_t j = 0; j < 2 * p; ++j) {
            const uint32_t r = rnd() % 100;
            gt[s * pitch + j] = r < 70 ? 0 : (r < 94 ? 1 : (r < 98 ? 2 : 255));
        }

    /* haplotypes: 0..5 alleles on distinct columns; mostly allele 0 so that results are not all zero */
    int64_t* hap_off = (int64_t*)malloc(sizeof(int64_t) * (size_t)(H + 1));
    int32_t* ref_col = (int32_t*)malloc(sizeof(int32_t) * (size_t)(H * 5 + 1));
    uint8_t* ref_allele = (uint8_t*)malloc((size_t)(H * 5 + 1));
    int64_t K = 0;
    for (int64_t h = 0; h < H; ++h) {
        hap_off[h] = K;
        int k = (int)(rnd() % 6);
        if (k > p) k = (int)p;
        for (int i = 0; i < k; ++i) {
            int32_t col;
            int dup;
            do {
                col = (int32_t)(rnd() % (uint32_t)p);
                dup = 0;
                for (int64_t r = hap_off[h]; r < K; ++r) dup |= ref_col[r] == col;
            } while (dup);
            const uint32_t r = rnd() % 20;
            ref_col[K] = col;
            ref_allele[K] = r < 14 ? 0 : (r < 19 ? 1 : 2);
            ++K;
        }
    }
    hap_off[H] = K;

    /* planner: keys = distinct (column, allele), numbered by (column, allele) */
    int32_t* key_of = (int32_t*)malloc(sizeof(int32_t) * (size_t)(p * MAX_ALLELE));
    memset(key_of, 0, sizeof(int32_t) * (size_t)(p * MAX_ALLELE));
    for (int64_t r = 0; r < K; ++r) key_of[ref_col[r] * MAX_ALLELE + ref_allele[r]] = 1;
    uint32_t* var_col = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(p + 1));
    uint32_t* var_key_off = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(p + 2));
    uint8_t* key_allele = (uint8_t*)malloc((size_t)(p * MAX_ALLELE + 1));
    int32_t* col_var = (int32_t*)malloc(sizeof(int32_t) * (size_t)p);
    int32_t* col_key01 = (int32_t*)malloc(sizeof(int32_t) * (size_t)(2 * p));
    uint32_t* var_other = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(p + 1));
    int64_t V = 0, A = 0, n_other = 0;
    for (int64_t v = 0; v < p; ++v) {
        int used = 0;
        for (int a = 0; a < MAX_ALLELE; ++a) used |= key_of[v * MAX_ALLELE + a];
        col_var[v] = used ? (int32_t)V : -1;
        col_key01[2 * v] = col_key01[2 * v + 1] = -1;
        if (!used) continue;
        var_col[V] = (uint32_t)v;
        var_key_off[V] = (uint32_t)A;
        for (int a = 0; a < MAX_ALLELE; ++a) {
            if (!key_of[v * MAX_ALLELE + a]) {
                key_of[v * MAX_ALLELE + a] = -1;
                continue;
            }
            key_allele[A] = (uint8_t)a;
            key_of[v * MAX_ALLELE + a] = (int32_t)A++;
        }
        if (key_of[v * MAX_ALLELE + 2] >= 0) {
            var_other[n_other++] = (uint32_t)V; /* a key outside {0, 1}: the whole variant goes the generic way */
        } else {
            col_key01[2 * v] = key_of[v * MAX_ALLELE + 0];
            col_key01[2 * v + 1] = key_of[v * MAX_ALLELE + 1];
        }
        ++V;
    }
    var_key_off[V] = (uint32_t)A;
    uint32_t* h_off32 = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(H + 1));
    uint32_t* h_key = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(K + 1));
    for (int64_t h = 0; h <= H; ++h) h_off32[h] = (uint32_t)hap_off[h];
    for (int64_t r = 0; r < K; ++r) h_key[r] = (uint32_t)key_of[ref_col[r] * MAX_ALLELE + ref_allele[r]];

    uint8_t* want = (uint8_t*)malloc((size_t)(n * H * 2 + 1));
    expect(gt, pitch, n, hap_off, ref_col, ref_allele, H, want);
    uint64_t* want_counts = (uint64_t*)calloc((size_t)H + 1, sizeof(uint64_t));
    for (int64_t s = 0; s < n; ++s)
        for (int64_t h = 0; h < H; ++h) want_counts[h] += want[(s * H + h) * 2] + want[(s * H + h) * 2 + 1];

    /* device buffers: the caller owns all of them */
    ht_pack_tables t;
    memset(&t, 0, sizeof(t));
    t.n_cols = p; t.n_vars = V; t.n_keys = A; t.n_var_other = n_other;
    t.var_col = (const uint32_t*)to_device(var_col, sizeof(uint32_t) * (size_t)V);
    t.var_key_off = (const uint32_t*)to_device(var_key_off, sizeof(uint32_t) * (size_t)(V + 1));
    t.key_allele = (const uint8_t*)to_device(key_allele, (size_t)A);
    t.col_var = (const int32_t*)to_device(col_var, sizeof(int32_t) * (size_t)p);
    t.col_key01 = (const int32_t*)to_device(col_key01, sizeof(int32_t) * (size_t)(2 * p));
    t.var_other = (const uint32_t*)to_device(var_other, sizeof(uint32_t) * (size_t)n_other);
    uint8_t* d_gt = (uint8_t*)to_device(gt, (size_t)(n * pitch));
    uint32_t* d_hap_off = (uint32_t*)to_device(h_off32, sizeof(uint32_t) * (size_t)(H + 1));
    uint32_t* d_hap_key = (uint32_t*)to_device(h_key, sizeof(uint32_t) * (size_t)K);
    const size_t plane_bytes = ht_plane_bytes(n, A);
    const int64_t tiles = ht_num_tiles(n), out_pitch = (2 * H + 15) / 16 * 16;
    uint32_t *d_planes = NULL, *d_bits = NULL;
    uint8_t* d_out = NULL;
    uint64_t* d_counts = NULL;
    CU(cudaMalloc((void**)&d_planes, plane_bytes ? plane_bytes : 16));
    CU(cudaMalloc((void**)&d_out, (size_t)(n * out_pitch) + 16));
    CU(cudaMalloc((void**)&d_bits, (size_t)(tiles * H) * HT_RECORD_WORDS * sizeof(uint32_t) + 16));
    CU(cudaMalloc((void**)&d_counts, sizeof(uint64_t) * (size_t)(H + 1)));
    if (!t.var_col || !t.var_key_off || !t.key_allele || !t.col_var || !t.col_key01 || !t.var_other || !d_gt ||
        !d_hap_off || !d_hap_key) {
        fprintf(stderr, "device allocation failed\n");
        return 2;
    }

    uint8_t* got = (uint8_t*)malloc((size_t)(n * H * 2 + 1));
    uint64_t* got_counts = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(H + 1));
    static const int modes[] = {HT_PACK_AUTO, HT_PACK_GENERIC, HT_PACK_SAMPLE_MAJOR, HT_PACK_SAMPLE_MAJOR_BIALLELIC};
    static const char* mode_names[] = {"auto", "generic", "sample-major", "sample-major biallelic"};
    int failures = 0;
    for (int m = 0; m < 4; ++m) {
        /* uint8 result with the fused allele counts */
        CU(cudaMemset(d_planes, 0xA5, plane_bytes));
        CU(cudaMemset(d_out, 0x77, (size_t)(n * out_pitch)));
        HT(ht_pack_u8(d_gt, pitch, 2, 1, NULL, 0, 0, 0, n, &t, d_planes, NULL, modes[m], NULL));
        HT(ht_transform_u8(d_planes, n, A, d_hap_off, d_hap_key, NULL, NULL, H, d_out, out_pitch, d_counts, NULL));
        HT(ht_copy2d(got, 2 * H, d_out, out_pitch, 2 * H, n, HT_COPY_D2H, NULL));
        CU(cudaMemcpy(got_counts, d_counts, sizeof(uint64_t) * (size_t)H, cudaMemcpyDeviceToHost));
        CU(cudaDeviceSynchronize());
        int bad = memcmp(got, want, (size_t)(n * H * 2)) != 0;
        int bad_counts = memcmp(got_counts, want_counts, sizeof(uint64_t) * (size_t)H) != 0;
        /* bit-packed result, expanded and counted by the library */
        CU(cudaMemset(d_out, 0x77, (size_t)(n * out_pitch)));
        HT(ht_transform_bits(d_planes, n, A, d_hap_off, d_hap_key, H, d_bits, NULL));
        HT(ht_unpack_bits_u8(d_bits, n, H, d_out, out_pitch, NULL));
        HT(ht_count_bits(d_bits, n, H, d_counts, NULL));
        HT(ht_copy2d(got, 2 * H, d_out, out_pitch, 2 * H, n, HT_COPY_D2H, NULL));
        CU(cudaMemcpy(got_counts, d_counts, sizeof(uint64_t) * (size_t)H, cudaMemcpyDeviceToHost));
        CU(cudaDeviceSynchronize());
        int bad_bits = memcmp(got, want, (size_t)(n * H * 2)) != 0;
        bad_counts |= memcmp(got_counts, want_counts, sizeof(uint64_t) * (size_t)H) != 0;
        printf("pack %-24s u8 %s, bits %s, counts %s\n", mode_names[m], bad ? "DIFFERS" : "ok",
               bad_bits ? "DIFFERS" : "ok", bad_counts ? "DIFFER" : "ok");
        failures += bad + bad_bits + bad_counts;
    }
    /* error reporting across the boundary: codes and a message, no exceptions */
    if (ht_transform_u8(d_planes, -1, A, d_hap_off, d_hap_key, NULL, NULL, H, d_out, out_pitch, NULL, NULL) != HT_ERR_BAD_ARG ||
        !strlen(ht_last_error())) {
        printf("bad argument was not reported\n");
        ++failures;
    }
    uint64_t ones = 0;
    for (int64_t i = 0; i < n * H * 2; ++i) ones += want[i];
    printf("%lld samples x %lld columns x %lld haplotypes (%lld alleles, %lld keys, %lld variants the generic way), "
           "%.3f of the result set: %s\n",
           (long long)n, (long long)p, (long long)H, (long long)K, (long long)A, (long long)n_other,
           (double)ones / (double)(n > 0 && H > 0 ? n * H * 2 : 1), failures ? "FAILED" : "all checks passed");
    return failures ? 4 : 0;
}
