// Microbenchmark: writing the transform's result pattern -- CTA = 1024 rows x 256 B of a
// (n, 2H) byte matrix, row pitch 2H = 200 KB -- with (a) 16-byte streaming stores from
// registers (what transform_u8_kernel does) and (b) TMA 2-D tensor stores from shared memory
// (cp.async.bulk.tensor.2d.global.shared::cta), for several box heights and CTAs per SM.
//   nvcc -arch=sm_100a -O3 tmastore.cu -lcuda -o tmastore
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>

#define CK(x) do { CUresult r_ = (x); if (r_ != CUDA_SUCCESS) { const char* s_; cuGetErrorString(r_, &s_); printf("%s failed: %s\n", #x, s_); exit(1); } } while (0)

__global__ void __launch_bounds__(256) stg_kernel(uint8_t* out, int64_t pitch, uint32_t n_hb) {
    const int64_t tile = blockIdx.x / n_hb, hb = blockIdx.x % n_hb;
    const int c0 = (threadIdx.x >> 5) * 2 + ((threadIdx.x >> 4) & 1);
    uint8_t* p = out + hb * 256 + 16 * (threadIdx.x & 15) + (tile * 1024 + c0) * pitch;
    const uint32_t v = threadIdx.x & 1;
#pragma unroll 4
    for (int L = 0; L < 64; ++L) {
        asm volatile("st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v), "r"(v), "r"(v), "r"(v) : "memory");
        p += 16 * pitch;
    }
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// box = 256 B x ROWS rows; the CTA fills a box in shared memory (STS.128, 16 lanes per row) and
// one thread issues the tensor store; NBUF boxes in flight per CTA.
template <int ROWS, int NBUF>
__global__ void __launch_bounds__(256) tma_kernel(const __grid_constant__ CUtensorMap tmap, uint32_t n_hb) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int tile = blockIdx.x / n_hb, hb = blockIdx.x % n_hb;
    const uint32_t v = threadIdx.x & 1;
    constexpr int kBoxBytes = ROWS * 256;
    for (int c = 0; c < 1024 / ROWS; ++c) {
        uint8_t* buf = smem + (c % NBUF) * kBoxBytes;
        if (c >= NBUF) {  // the store that used this buffer must have read it
            if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(NBUF - 1) : "memory");
            __syncthreads();
        }
        for (int i = threadIdx.x; i < kBoxBytes / 16; i += 256)
            *reinterpret_cast<uint4*>(buf + 16 * i) = make_uint4(v, v, v, v);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (threadIdx.x == 0) {
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(&tmap),
                         "r"(hb * 256), "r"(tile * 1024 + c * ROWS), "r"(smem_u32(buf))
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// per-warp boxes: warp w owns rows [128w, 128w + 128) of the CTA's tile and stores them in boxes
// of ROWS rows from its own NBUF staging buffers; no CTA-wide barrier in the loop.
template <int ROWS, int NBUF>
__global__ void __launch_bounds__(256) tma_warp_kernel(const __grid_constant__ CUtensorMap tmap, uint32_t n_hb) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int tile = blockIdx.x / n_hb, hb = blockIdx.x % n_hb;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t v = threadIdx.x & 1;
    constexpr int kBoxBytes = ROWS * 256;
    uint8_t* mine = smem + warp * NBUF * kBoxBytes;
    for (int c = 0; c < 128 / ROWS; ++c) {
        uint8_t* buf = mine + (c % NBUF) * kBoxBytes;
        if (c >= NBUF) {
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(NBUF - 1) : "memory");
            __syncwarp();
        }
        for (int i = lane; i < kBoxBytes / 16; i += 32)
            *reinterpret_cast<uint4*>(buf + 16 * i) = make_uint4(v, v, v, v);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(&tmap),
                         "r"(hb * 256), "r"(tile * 1024 + warp * 128 + c * ROWS), "r"(smem_u32(buf))
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
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

template <int ROWS, int NBUF>
static void run_tma(const CUtensorMap& tmap, uint32_t n_tiles, uint32_t n_hb, double bytes, int pad_kb) {
    const size_t smem = (size_t)ROWS * 256 * NBUF + (size_t)pad_kb * 1024;  // pad limits CTAs per SM
    cudaFuncSetAttribute(tma_kernel<ROWS, NBUF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, tma_kernel<ROWS, NBUF>, 256, smem);
    float ms = timeit([&] { tma_kernel<ROWS, NBUF><<<n_tiles * n_hb, 256, smem>>>(tmap, n_hb); });
    printf("tma   box %3d rows x %d bufs, %d CTAs/SM: %7.2f ms  %6.0f GB/s\n", ROWS, NBUF, occ, ms, bytes / ms / 1e6);
}

template <int ROWS, int NBUF>
static void run_tma_warp(const CUtensorMap& tmap, uint32_t n_tiles, uint32_t n_hb, double bytes, int pad_kb) {
    const size_t smem = (size_t)ROWS * 256 * NBUF * 8 + (size_t)pad_kb * 1024;
    cudaFuncSetAttribute(tma_warp_kernel<ROWS, NBUF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, tma_warp_kernel<ROWS, NBUF>, 256, smem);
    float ms = timeit([&] { tma_warp_kernel<ROWS, NBUF><<<n_tiles * n_hb, 256, smem>>>(tmap, n_hb); });
    printf("tma per-warp box %3d rows x %d bufs, %d CTAs/SM: %7.2f ms  %6.0f GB/s\n", ROWS, NBUF, occ, ms, bytes / ms / 1e6);
}

int main() {
    cudaFree(0);
    const int64_t n = 200 * 1024, H = 100000, pitch = 2 * H;
    const uint32_t n_tiles = n / 1024, n_hb = (H + 127) / 128;
    const double bytes = (double)n * pitch;
    uint8_t* out; cudaMalloc(&out, n * pitch + 4096);
    float ms = timeit([&] { stg_kernel<<<n_tiles * n_hb, 256>>>(out, pitch, n_hb); });
    printf("stg.cs.v4 from registers:              %7.2f ms  %6.0f GB/s\n", ms, bytes / ms / 1e6);

    auto make = [&](int rows) {
        CUtensorMap m;
        cuuint64_t dims[2] = {(cuuint64_t)pitch, (cuuint64_t)n};
        cuuint64_t strides[1] = {(cuuint64_t)pitch};
        cuuint32_t box[2] = {256, (cuuint32_t)rows};
        cuuint32_t estr[2] = {1, 1};
        CK(cuTensorMapEncodeTiled(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, out, dims, strides, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                  CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE));
        return m;
    };
    CUtensorMap m32 = make(32), m64 = make(64), m128 = make(128), m256 = make(256);
    run_tma<32, 2>(m32, n_tiles, n_hb, bytes, 0);
    run_tma<64, 2>(m64, n_tiles, n_hb, bytes, 0);
    run_tma<128, 2>(m128, n_tiles, n_hb, bytes, 0);
    run_tma<128, 2>(m128, n_tiles, n_hb, bytes, 40);   // 2 CTAs/SM
    run_tma<128, 2>(m128, n_tiles, n_hb, bytes, 100);  // 1 CTA/SM
    run_tma<256, 2>(m256, n_tiles, n_hb, bytes, 0);
    run_tma<64, 4>(m64, n_tiles, n_hb, bytes, 0);
    run_tma<32, 4>(m32, n_tiles, n_hb, bytes, 0);
    CUtensorMap m8 = make(8), m16 = make(16);
    run_tma_warp<16, 2>(m16, n_tiles, n_hb, bytes, 0);    // 64 KB -> 3 CTAs/SM
    run_tma_warp<16, 2>(m16, n_tiles, n_hb, bytes, 46);   // 110 KB -> 2 CTAs/SM
    run_tma_warp<16, 2>(m16, n_tiles, n_hb, bytes, 100);  // 1 CTA/SM
    run_tma_warp<8, 2>(m8, n_tiles, n_hb, bytes, 0);      // 32 KB
    run_tma_warp<8, 2>(m8, n_tiles, n_hb, bytes, 42);     // 74 KB -> 3 CTAs/SM
    run_tma_warp<8, 2>(m8, n_tiles, n_hb, bytes, 78);     // 110 KB -> 2 CTAs/SM
    run_tma_warp<32, 2>(m32, n_tiles, n_hb, bytes, 0);    // 128 KB -> 1 CTA/SM
    run_tma_warp<16, 4>(m16, n_tiles, n_hb, bytes, 0);    // 128 KB -> 1 CTA/SM
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
