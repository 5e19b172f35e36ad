// Microbenchmark: does L2 compute-data-compression (compressible VMM allocations) speed up writing a
// mostly-zero byte matrix (what the transform's result looks like: ~97 % zero bytes)?
//   nvcc -arch=sm_100a -O3 compress.cu -lcuda -o compress
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#define CK(x) do { CUresult r_ = (x); if (r_ != CUDA_SUCCESS) { const char* s_; cuGetErrorString(r_, &s_); printf("%s failed: %s\n", #x, s_); return 1; } } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t x) { x += 0x9e3779b97f4a7c15ull; x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull; x = (x ^ (x >> 27)) * 0x94d049bb133111ebull; return x ^ (x >> 31); }

// every thread writes 16 bytes; about ones_per_1024 / 1024 of the bytes are 1 (cheap pattern: one
// hash per 16 bytes decides which bytes are set)
__global__ void fill(uint4* out, size_t n16, uint32_t ones_per_1024) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
        const uint64_t r = mix(i);
        uint32_t w[4] = {0u, 0u, 0u, 0u};
        if (ones_per_1024 >= 512) {  // dense: random 0/1 bytes
            w[0] = (uint32_t)r & 0x01010101u; w[1] = (uint32_t)(r >> 1) & 0x01010101u;
            w[2] = (uint32_t)(r >> 2) & 0x01010101u; w[3] = (uint32_t)(r >> 3) & 0x01010101u;
        } else if (((r >> 8) & 1023) < 16 * ones_per_1024) {  // sparse: one byte of the 16
            w[(r >> 2) & 3] = 1u << (8 * (r & 3));
        }
        asm volatile("st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(out + i), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
    }
}
__global__ void rd(const uint4* in, size_t n16, uint32_t* sink) {
    uint32_t acc = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) { uint4 v = in[i]; acc += v.x ^ v.y ^ v.z ^ v.w; }
    if (acc == 0x12345) *sink = acc;
}

static float time_fill(uint4* p, size_t n16, uint32_t ones) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    fill<<<148 * 16, 256>>>(p, n16, ones);
    cudaEventRecord(e0);
    for (int i = 0; i < 3; ++i) fill<<<148 * 16, 256>>>(p, n16, ones);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms / 3;
}
static float time_rd(const uint4* p, size_t n16, uint32_t* sink) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    rd<<<148 * 16, 256>>>(p, n16, sink);
    cudaEventRecord(e0);
    for (int i = 0; i < 3; ++i) rd<<<148 * 16, 256>>>(p, n16, sink);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms / 3;
}

int main() {
    cudaFree(0);
    CUdevice dev; CK(cuDeviceGet(&dev, 0));
    int supported = 0;
    CK(cuDeviceGetAttribute(&supported, CU_DEVICE_ATTRIBUTE_GENERIC_COMPRESSION_SUPPORTED, dev));
    printf("generic compression supported: %d\n", supported);
    const size_t bytes = (size_t)16 << 30, n16 = bytes / 16;
    uint32_t* sink; cudaMalloc(&sink, 4);
    uint4* plain; cudaMalloc(&plain, bytes);
    for (uint32_t ones : {0u, 27u, 512u})
        printf("plain        ones=%4u/1024: write %7.0f GB/s   read %7.0f GB/s\n", ones, bytes / time_fill(plain, n16, ones) / 1e6, bytes / time_rd(plain, n16, sink) / 1e6);
    cudaFree(plain);
    if (!supported) return 0;
    CUmemAllocationProp prop = {};
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = 0;
    prop.allocFlags.compressionType = CU_MEM_ALLOCATION_COMP_GENERIC;
    size_t gran = 0; CK(cuMemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
    size_t sz = (bytes + gran - 1) / gran * gran;
    CUmemGenericAllocationHandle h; CK(cuMemCreate(&h, sz, &prop, 0));
    CUmemAllocationProp got = {}; CK(cuMemGetAllocationPropertiesFromHandle(&got, h));
    printf("allocation compressionType = %d (granularity %zu)\n", (int)got.allocFlags.compressionType, gran);
    CUdeviceptr va; CK(cuMemAddressReserve(&va, sz, 0, 0, 0)); CK(cuMemMap(va, sz, 0, h, 0));
    CUmemAccessDesc acc = {}; acc.location = prop.location; acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    CK(cuMemSetAccess(va, sz, &acc, 1));
    for (uint32_t ones : {0u, 27u, 512u})
        printf("compressible ones=%4u/1024: write %7.0f GB/s   read %7.0f GB/s\n", ones, bytes / time_fill((uint4*)va, n16, ones) / 1e6, bytes / time_rd((const uint4*)va, n16, sink) / 1e6);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
