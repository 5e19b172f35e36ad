// Microbenchmark: achieved HBM read bandwidth of a "column chunk down the rows" access pattern
// (what the sample-major pack kernel does) as a function of load width per lane, loads in
// flight per thread and CTAs per SM.   nvcc -arch=sm_100a -O3 ldbw.cu -o ldbw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <typename T> __device__ __forceinline__ uint32_t fold(T v);
template <> __device__ __forceinline__ uint32_t fold<uint32_t>(uint32_t v) { return v; }
template <> __device__ __forceinline__ uint32_t fold<uint2>(uint2 v) { return v.x ^ v.y; }
template <> __device__ __forceinline__ uint32_t fold<uint4>(uint4 v) { return v.x ^ v.y ^ v.z ^ v.w; }

// grid: blockIdx.x = column block, blockIdx.y = row block of `rows_per_cta` rows
template <typename T, int U>
__global__ void __launch_bounds__(256) rd(const uint8_t* __restrict__ base, size_t row_bytes, int rows_per_cta,
                                          uint32_t* __restrict__ sink) {
    const size_t col = ((size_t)blockIdx.x * 256 + threadIdx.x) * sizeof(T);
    if (col + sizeof(T) > row_bytes) return;
    const uint8_t* p = base + (size_t)blockIdx.y * rows_per_cta * row_bytes + col;
    uint32_t acc = 0;
    for (int r = 0; r < rows_per_cta; r += U) {
        T v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = __ldg(reinterpret_cast<const T*>(p + (size_t)(r + u) * row_bytes));
#pragma unroll
        for (int u = 0; u < U; ++u) acc += fold<T>(v[u]);
    }
    if (acc == 0x12345678u) *sink = acc;
}

template <typename T, int U>
void run(const char* name, const uint8_t* d, size_t rows, size_t row_bytes, uint32_t* sink) {
    const int rows_per_cta = 1024;
    dim3 grid((unsigned)((row_bytes / sizeof(T) + 255) / 256), (unsigned)(rows / rows_per_cta));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 2; ++i) rd<T, U><<<grid, 256>>>(d, row_bytes, rows_per_cta, sink);
    cudaEventRecord(e0);
    const int reps = 5;
    for (int i = 0; i < reps; ++i) rd<T, U><<<grid, 256>>>(d, row_bytes, rows_per_cta, sink);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
    int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, rd<T, U>, 256, 0);
    printf("%-10s U=%2d  ctas/SM=%d  %8.3f ms  %7.0f GB/s\n", name, U, occ, ms, rows * row_bytes / ms / 1e6);
}

int main() {
    const size_t row_bytes = 100000, rows = 160 * 1024;  // 16.4 GB
    uint8_t* d; uint32_t* sink;
    cudaMalloc(&d, rows * row_bytes); cudaMalloc(&sink, 4);
    cudaMemset(d, 1, rows * row_bytes);
    run<uint32_t, 8>("LDG.32", d, rows, row_bytes, sink);
    run<uint32_t, 16>("LDG.32", d, rows, row_bytes, sink);
    run<uint32_t, 32>("LDG.32", d, rows, row_bytes, sink);
    run<uint2, 8>("LDG.64", d, rows, row_bytes, sink);
    run<uint2, 16>("LDG.64", d, rows, row_bytes, sink);
    run<uint2, 32>("LDG.64", d, rows, row_bytes, sink);
    run<uint4, 4>("LDG.128", d, rows, row_bytes, sink);
    run<uint4, 8>("LDG.128", d, rows, row_bytes, sink);
    run<uint4, 16>("LDG.128", d, rows, row_bytes, sink);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
