// Microbenchmark for the transform's store phase: CTA = half tile (512 rows) x 256 B of an (n, 2H) byte matrix.
//  (a) st.global.cs.v4 from registers in the kernel's pattern (thread (cb, g): rows 32 g + slot(m), + 4)
//  (b) the same bytes staged in shared memory per iteration (2 boxes of 16 rows x 256 B, the rows 32 apart) and
//      written by 3-D tensor stores (cp.async.bulk.tensor.3d.global.shared::cta), 3 staging buffers, ONE
//      __syncthreads per iteration
//   nvcc -arch=sm_100a -O3 tmastore3d.cu -lcuda -o tmastore3d
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>

#define CK(x) do { CUresult r_ = (x); if (r_ != CUDA_SUCCESS) { const char* s_; cuGetErrorString(r_, &s_); printf("%s failed: %s\n", #x, s_); exit(1); } } while (0)

__device__ __forceinline__ int slot_sample(int m) { return 8 * (m & 3) + (m >> 2); }
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(256, 4) stg_kernel(uint8_t* out, int64_t pitch, uint32_t n_hb) {
    extern __shared__ __align__(128) uint8_t pad[];
    const int64_t ht = blockIdx.x / n_hb, hb = blockIdx.x % n_hb;
    const int cb = threadIdx.x & 15, g = threadIdx.x >> 4;
    uint8_t* p0 = out + hb * 256 + 16 * cb + (ht * 512 + 32 * g) * pitch;
    const uint32_t v = threadIdx.x & 1;
#pragma unroll 4
    for (int m = 0; m < 16; ++m) {
        uint8_t* p = p0 + (int64_t)slot_sample(m) * pitch;
        asm volatile("st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v), "r"(v), "r"(v), "r"(v) : "memory");
        asm volatile("st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p + 4 * pitch), "r"(v), "r"(v), "r"(v), "r"(v) : "memory");
    }
}

// CTA = 512 rows x W bytes (W = 512 or 1024): thread = 16 bytes of a row, W / 16 threads per row
template <int W>
__global__ void __launch_bounds__(256, 4) stg_wide_kernel(uint8_t* out, int64_t pitch, uint32_t n_wb) {
    extern __shared__ __align__(128) uint8_t pad[];
    constexpr int kPerRow = W / 16, kRowsPerPass = 256 / kPerRow;
    const int64_t ht = blockIdx.x / n_wb, wb = blockIdx.x % n_wb;
    const int cb = threadIdx.x % kPerRow, g = threadIdx.x / kPerRow;
    uint8_t* p0 = out + wb * W + 16 * cb + (ht * 512 + g) * pitch;
    const uint32_t v = threadIdx.x & 1;
#pragma unroll 4
    for (int m = 0; m < 512 / kRowsPerPass; ++m) {
        uint8_t* p = p0 + (int64_t)m * kRowsPerPass * pitch;
        if (wb * W + 16 * cb < pitch)
            asm volatile("st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v), "r"(v), "r"(v), "r"(v) : "memory");
    }
}

template <int NBUF>
__global__ void __launch_bounds__(256, 4) tma3d_kernel(const __grid_constant__ CUtensorMap tmap, uint32_t n_hb) {
    extern __shared__ __align__(128) uint8_t smem[];  // NBUF x 2 boxes x 4 KB (+ padding that sets the occupancy)
    const int ht = blockIdx.x / n_hb, hb = blockIdx.x % n_hb;
    const int cb = threadIdx.x & 15, g = threadIdx.x >> 4;
    const uint32_t v = threadIdx.x & 1;
#pragma unroll 1
    for (int m = 0; m < 16; ++m) {
        uint8_t* buf = smem + (m % NBUF) * 8192;
        *reinterpret_cast<uint4*>(buf + g * 256 + cb * 16) = make_uint4(v, v, v, v);
        *reinterpret_cast<uint4*>(buf + 4096 + g * 256 + cb * 16) = make_uint4(v, v, v, v);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (threadIdx.x == 0) {
            const int j = slot_sample(m);
            asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(&tmap),
                         "r"(hb * 256), "r"(j), "r"(ht * 16), "r"(smem_u32(buf)) : "memory");
            asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(&tmap),
                         "r"(hb * 256), "r"(j + 4), "r"(ht * 16), "r"(smem_u32(buf + 4096)) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(NBUF - 2) : "memory");
        }
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <typename F>
static float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaEventRecord(e0);
    for (int i = 0; i < 3; ++i) f();
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaError_t err = cudaGetLastError(); if (err != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(err));
    return ms / 3;
}

int main() {
    cudaFree(0);
    const int64_t n = 200 * 1024, H = 100000, pitch = 2 * H;
    const uint32_t n_half = n / 512, n_hb = (H + 127) / 128;
    const double bytes = (double)n * pitch;
    uint8_t* out; cudaMalloc(&out, n * pitch + 4096);
    for (int pad_kb : {0, 29, 53}) {
        cudaFuncSetAttribute(stg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, pad_kb * 1024);
        int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, stg_kernel, 256, pad_kb * 1024);
        float ms = timeit([&] { stg_kernel<<<n_half * n_hb, 256, pad_kb * 1024>>>(out, pitch, n_hb); });
        printf("stg.cs.v4 from registers, %d CTAs/SM: %7.2f ms  %6.0f GB/s\n", occ, ms, bytes / ms / 1e6);
    }
    for (int pad_kb : {29, 53}) {
        int occ = 0;
        cudaFuncSetAttribute(stg_wide_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, pad_kb * 1024);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, stg_wide_kernel<512>, 256, pad_kb * 1024);
        uint32_t n_wb = (uint32_t)((pitch + 511) / 512);
        float ms = timeit([&] { stg_wide_kernel<512><<<n_half * n_wb, 256, pad_kb * 1024>>>(out, pitch, n_wb); });
        printf("stg.cs.v4, CTA = 512 rows x 512 B, %d CTAs/SM: %7.2f ms  %6.0f GB/s\n", occ, ms, bytes / ms / 1e6);
        cudaFuncSetAttribute(stg_wide_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, pad_kb * 1024);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, stg_wide_kernel<1024>, 256, pad_kb * 1024);
        n_wb = (uint32_t)((pitch + 1023) / 1024);
        ms = timeit([&] { stg_wide_kernel<1024><<<n_half * n_wb, 256, pad_kb * 1024>>>(out, pitch, n_wb); });
        printf("stg.cs.v4, CTA = 512 rows x 1024 B, %d CTAs/SM: %7.2f ms  %6.0f GB/s\n", occ, ms, bytes / ms / 1e6);
    }
    CUtensorMap m;
    cuuint64_t dims[3] = {(cuuint64_t)pitch, 32, (cuuint64_t)(n / 32)};
    cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)(32 * pitch)};
    cuuint32_t box[3] = {256, 1, 16};
    cuuint32_t estr[3] = {1, 1, 1};
    CK(cuTensorMapEncodeTiled(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, out, dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                              CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE));
    auto run = [&](auto kernel, int nbuf, int total_kb) {
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, total_kb * 1024);
        int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, 256, total_kb * 1024);
        float ms = timeit([&] { kernel<<<n_half * n_hb, 256, total_kb * 1024>>>(m, n_hb); });
        printf("tma 3-D boxes (2 x 16 rows x 256 B per iteration), %d buffers, %d KB smem, %d CTAs/SM: %7.2f ms  %6.0f GB/s\n", nbuf, total_kb, occ, ms, bytes / ms / 1e6);
    };
    run(tma3d_kernel<3>, 3, 24);
    run(tma3d_kernel<3>, 3, 53);
    run(tma3d_kernel<3>, 3, 72);
    run(tma3d_kernel<2>, 2, 45);
    run(tma3d_kernel<4>, 4, 53);
    return 0;
}
