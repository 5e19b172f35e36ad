// Microbenchmark: what L2 can deliver for the transform kernel's access mix.
//   (a) random 256-byte record reads (one coalesced uint2 per lane, U in flight per warp) out of a
//       working set of W MB that stays in L2 -- the AND phase's reads;
//   (b) the same reads interleaved with 16-byte streaming stores to a large buffer at the transform's
//       byte ratio (S stored bytes per R read bytes) -- both phases at once, no dependence between them.
// Prints achieved read / write / total TB/s.   nvcc -arch=sm_100a -O3 l2bw.cu -o l2bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t mix(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

// Each warp: `iters` rounds of U record reads (+ stores_per_round 512-byte warp stores).
template <int U>
__global__ void __launch_bounds__(256, 4) mixk(const uint2* __restrict__ recs, uint32_t n_recs, int iters,
                                               uint4* __restrict__ out, size_t out_vec16, int stores_per_round,
                                               uint32_t* __restrict__ sink, int l1_bypass) {
    const int lane = threadIdx.x & 31;
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    uint32_t acc0 = 0xffffffffu, acc1 = 0xffffffffu;
    uint32_t seed = gw * 0x9e3779b9u + 12345u;
    size_t so = (size_t)gw * 32 + lane;  // warp-contiguous 512-byte chunks, grid-strided
    for (int it = 0; it < iters; ++it) {
        uint2 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            seed = mix(seed + u + 1);
            const uint32_t r = (uint32_t)(((uint64_t)seed * n_recs) >> 32);
            const uint2* p = recs + (size_t)r * 32 + lane;
            if (l1_bypass) asm volatile("ld.global.cg.v2.u32 {%0, %1}, [%2];" : "=r"(v[u].x), "=r"(v[u].y) : "l"(p));
            else v[u] = __ldg(p);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) { acc0 &= v[u].x; acc1 &= v[u].y; }
        for (int s = 0; s < stores_per_round; ++s) {
            const uint4 w = make_uint4(acc0, acc1, it, s);
            asm volatile("st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(out + so % out_vec16), "r"(w.x), "r"(w.y), "r"(w.z), "r"(w.w) : "memory");
            so += (size_t)n_warps * 32;
        }
    }
    if ((acc0 ^ acc1) == 0x12345678u) *sink = acc0;
}

template <int U>
void run(const uint2* recs, double w_mb, int stores_per_round, uint4* out, size_t out_bytes, uint32_t* sink, int bypass) {
    const uint32_t n_recs = (uint32_t)(w_mb * 1e6 / 256);
    const int ctas = 148 * 4, iters = 2000 / U * 8 / 8;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    mixk<U><<<ctas, 256>>>(recs, n_recs, iters, out, out_bytes / 16, stores_per_round, sink, bypass);
    cudaEventRecord(e0);
    const int reps = 3;
    for (int i = 0; i < reps; ++i) mixk<U><<<ctas, 256>>>(recs, n_recs, iters, out, out_bytes / 16, stores_per_round, sink, bypass);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
    const double warps = ctas * 8.0;
    const double rd = warps * iters * U * 256.0, wr = warps * iters * (double)stores_per_round * 512.0;
    printf("W=%6.1f MB U=%2d %s stores/round=%d: %7.3f ms  read %6.2f TB/s  write %6.2f TB/s  total %6.2f TB/s\n", w_mb, U,
           bypass ? "ld.cg" : "ldg  ", stores_per_round, ms, rd / ms / 1e9, wr / ms / 1e9, (rd + wr) / ms / 1e9);
}

int main() {
    const size_t rec_bytes = 128ull << 20, out_bytes = 8ull << 30;
    uint2* recs; uint4* out; uint32_t* sink;
    cudaMalloc(&recs, rec_bytes); cudaMalloc(&out, out_bytes); cudaMalloc(&sink, 4);
    cudaMemset(recs, 0xff, rec_bytes);
    for (int bypass = 0; bypass < 2; ++bypass)
        for (double w : {1.0, 5.1, 12.8, 25.6, 51.2, 102.4}) run<8>(recs, w, 0, out, out_bytes, sink, bypass);
    run<16>(recs, 25.6, 0, out, out_bytes, sink, 0);
    run<16>(recs, 12.8, 0, out, out_bytes, sink, 0);
    // stores only (reads of a tiny set: L1 hits), then the transform's mix: 8 reads (2 KB) : 3 stores (1.5 KB) ~ 55 : 40
    run<8>(recs, 0.01, 3, out, out_bytes, sink, 0);
    run<8>(recs, 0.01, 6, out, out_bytes, sink, 0);
    for (double w : {5.1, 12.8, 25.6, 51.2}) run<8>(recs, w, 3, out, out_bytes, sink, 0);
    for (double w : {5.1, 12.8, 25.6}) run<8>(recs, w, 6, out, out_bytes, sink, 0);
    for (double w : {12.8, 25.6}) run<8>(recs, w, 3, out, out_bytes, sink, 1);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
