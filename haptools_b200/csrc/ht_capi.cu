// C-ABI glue of libhaptools_cuda: version, thread-local error string, workspace sizes and a
// few host<->device copy helpers.  See include/haptools_cuda.h.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ht_common.cuh"

namespace ht {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t err, const char* what) {
    set_error("CUDA error %d (%s) in %s", (int)err, cudaGetErrorString(err), what);
    return HT_ERR_CUDA;
}

int bad_arg(const char* what) {
    set_error("bad argument: %s", what);
    return HT_ERR_BAD_ARG;
}

// ---------------------------------------------------------------------------------------
// Copies between PAGEABLE host memory and the device go through the host copy engine (ht_hostcopy.cu: a
// persistent pool of host threads with page-locked staging slots); the synchronous forms here are
// submit + wait.  cudaMemcpy2DAsync would stage such memory through the driver's own bounce buffer with one
// host thread (11 GB/s up, 17.6 GB/s down on the B200 box).
// ---------------------------------------------------------------------------------------
namespace hostcopy {
enum Kind { kUpload = 0, kDownload = 1, kDownloadExpand = 2 };
constexpr size_t kSlotBytes = 8u << 20;
constexpr int kMaxDevs = 16;
int64_t submit(int kind, void* dst, int64_t dst_pitch, const void* src, int64_t src_pitch, int64_t width, int64_t height,
               int64_t out_width, const uint32_t* row_idx, cudaStream_t st);
int wait(int64_t ticket);
}  // namespace hostcopy

static int staged_copy(void* dst, int64_t dst_pitch, const void* src, int64_t src_pitch, int64_t width, int64_t height,
                       cudaStream_t st, const uint32_t* row_idx = nullptr, bool download = false) {
    const int64_t ticket = hostcopy::submit(download ? hostcopy::kDownload : hostcopy::kUpload, dst, dst_pitch, src,
                                            src_pitch, width, height, 0, row_idx, st);
    if (ticket < 0) return HT_ERR_CUDA;
    return hostcopy::wait(ticket);
}

static bool is_pageable(const void* p) {
    cudaPointerAttributes at;
    const cudaError_t e = cudaPointerGetAttributes(&at, p);
    if (e != cudaSuccess) cudaGetLastError();  // old drivers report unregistered memory as an error
    return e != cudaSuccess || at.type == cudaMemoryTypeUnregistered;
}

// ---------------------------------------------------------------------------------------
// Row gather straight out of PAGE-LOCKED host memory: the SMs read the selected rows over PCIe (mapped
// host memory, unified addressing) and write them to the device buffer -- no host thread touches a byte
// and the call is asynchronous on the stream.  A warp per row, 16-byte loads, four in flight per lane.
// ---------------------------------------------------------------------------------------
template <int kUnroll>
__global__ void __launch_bounds__(256) gather_rows_kernel(uint8_t* __restrict__ dst, int64_t dst_pitch,
                                                          const uint8_t* __restrict__ src, int64_t src_pitch,
                                                          int64_t width, const uint32_t* __restrict__ row_idx,
                                                          int64_t n_rows) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int64_t n16 = width >> 4;
    for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < n_rows; r += warps) {
        const uint8_t* srow = src + (int64_t)row_idx[r] * src_pitch;
        uint8_t* drow = dst + r * dst_pitch;
        const uint4* s = reinterpret_cast<const uint4*>(srow);
        uint4* d = reinterpret_cast<uint4*>(drow);
        int64_t i = lane;
        for (; i + 32 * (kUnroll - 1) < n16; i += 32 * kUnroll) {
            uint4 v[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) v[u] = s[i + 32 * u];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) d[i + 32 * u] = v[u];
        }
        {  // the last, partial batch: still all of its loads before the first store
            uint4 v[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u)
                if (i + 32 * u < n16) v[u] = s[i + 32 * u];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u)
                if (i + 32 * u < n16) d[i + 32 * u] = v[u];
        }
        const int64_t t0 = n16 << 4;  // the last width % 16 bytes
        if (t0 + lane < width) drow[t0 + lane] = srow[t0 + lane];
    }
}

// true if the rows were gathered by the kernel; false: the source is not page-locked / not 16-byte aligned
static int gather_rows_from_pinned(void* dst, int64_t dst_pitch, const void* src, int64_t src_pitch, int64_t width,
                                   const uint32_t* row_idx, int64_t n_rows, cudaStream_t st, bool* done) {
    *done = false;
    const char* mode = getenv("HT_GATHER");  // "bounce": always the host threads (measurements, tests)
    if (mode && mode[0] == 'b') return HT_OK;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, src) != cudaSuccess) {
        cudaGetLastError();
        return HT_OK;
    }
    if (at.type != cudaMemoryTypeHost || !at.devicePointer) return HT_OK;
    const uint8_t* dsrc = static_cast<const uint8_t*>(at.devicePointer);
    if (((uintptr_t)dsrc | (uintptr_t)dst | (uintptr_t)src_pitch | (uintptr_t)dst_pitch) & 15) return HT_OK;
    // the row indices in stream order: allocated, filled, used and freed on `st`
    uint32_t* d_idx = nullptr;
    HT_CUDA_TRY(cudaMallocAsync((void**)&d_idx, sizeof(uint32_t) * (size_t)n_rows, st));
    cudaError_t e = cudaMemcpyAsync(d_idx, row_idx, sizeof(uint32_t) * (size_t)n_rows, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        const int64_t ctas = (n_rows + 7) / 8;
        const char* ce = getenv("HT_GATHER_CTAS");  // measurements
        const char* ue = getenv("HT_GATHER_UNROLL");
        const int64_t max_ctas = ce && atoi(ce) > 0 ? atoi(ce) : 148 * 8;
        const unsigned grid = (unsigned)(ctas < max_ctas ? ctas : max_ctas);
        // loads in flight per lane and the grid make no difference (48.0-48.8 GB/s for 4 / 8 / 16 loads and 148 to
        // 2,368 CTAs against 55.6 GB/s for a DMA of the same bytes, profiles/r2/gather.json): the link's read
        // protocol bounds SM-issued reads, not the kernel
        if (ue && atoi(ue) == 8)
            gather_rows_kernel<8><<<grid, 256, 0, st>>>(static_cast<uint8_t*>(dst), dst_pitch, dsrc, src_pitch, width, d_idx, n_rows);
        else if (ue && atoi(ue) == 16)
            gather_rows_kernel<16><<<grid, 256, 0, st>>>(static_cast<uint8_t*>(dst), dst_pitch, dsrc, src_pitch, width, d_idx, n_rows);
        else
            gather_rows_kernel<4><<<grid, 256, 0, st>>>(static_cast<uint8_t*>(dst), dst_pitch, dsrc, src_pitch, width, d_idx, n_rows);
        e = cudaGetLastError();
    }
    const cudaError_t ef = cudaFreeAsync(d_idx, st);
    if (e != cudaSuccess) return cuda_fail(e, "row gather from page-locked host memory");
    HT_CUDA_TRY(ef);
    *done = true;
    return HT_OK;
}

}  // namespace ht

extern "C" {

int ht_version(void) { return HT_ABI_VERSION; }

const char* ht_last_error(void) { return ht::g_err; }

int64_t ht_num_tiles(int64_t n_samples) {
    if (n_samples <= 0) return 0;
    return (n_samples + HT_TILE_SAMPLES - 1) / HT_TILE_SAMPLES;
}

size_t ht_plane_bytes(int64_t n_samples, int64_t n_keys) {
    if (n_samples <= 0 || n_keys <= 0) return 0;
    return (size_t)ht_num_tiles(n_samples) * (size_t)n_keys * HT_RECORD_WORDS * sizeof(uint32_t);
}

int ht_copy2d(void* dst, int64_t dst_pitch, const void* src, int64_t src_pitch,
              int64_t width_bytes, int64_t height, int kind, void* stream) {
    if (width_bytes < 0 || height < 0) return ht::bad_arg("ht_copy2d: negative extent");
    if (width_bytes == 0 || height == 0) return HT_OK;
    if (!dst || !src) return ht::bad_arg("ht_copy2d: NULL pointer");
    if (dst_pitch < width_bytes || src_pitch < width_bytes)
        return ht::bad_arg("ht_copy2d: pitch smaller than width");
    cudaMemcpyKind k;
    switch (kind) {
        case HT_COPY_H2D: k = cudaMemcpyHostToDevice; break;
        case HT_COPY_D2H: k = cudaMemcpyDeviceToHost; break;
        case HT_COPY_D2D: k = cudaMemcpyDeviceToDevice; break;
        default: return ht::bad_arg("ht_copy2d: unknown kind");
    }
    int cur_dev = 0;
    if ((kind == HT_COPY_H2D || kind == HT_COPY_D2H) && width_bytes * height >= (4 << 20) &&
        (size_t)width_bytes <= ht::hostcopy::kSlotBytes && cudaGetDevice(&cur_dev) == cudaSuccess &&
        cur_dev < ht::hostcopy::kMaxDevs && ht::is_pageable(kind == HT_COPY_H2D ? src : dst))
        return ht::staged_copy(dst, dst_pitch, src, src_pitch, width_bytes, height, (cudaStream_t)stream, nullptr,
                               kind == HT_COPY_D2H);
    if (dst_pitch == width_bytes && src_pitch == width_bytes) {  // dense on both sides: one linear copy
        HT_CUDA_TRY(cudaMemcpyAsync(dst, src, (size_t)width_bytes * (size_t)height, k, (cudaStream_t)stream));
        return HT_OK;
    }
    HT_CUDA_TRY(cudaMemcpy2DAsync(dst, (size_t)dst_pitch, src, (size_t)src_pitch,
                                  (size_t)width_bytes, (size_t)height, k,
                                  (cudaStream_t)stream));
    return HT_OK;
}

int ht_copy_rows_h2d(void* dst, int64_t dst_pitch, const void* src, int64_t src_pitch, int64_t width_bytes,
                     const uint32_t* row_idx, int64_t n_rows, void* stream) {
    if (width_bytes < 0 || n_rows < 0) return ht::bad_arg("ht_copy_rows_h2d: negative extent");
    if (width_bytes == 0 || n_rows == 0) return HT_OK;
    if (!dst || !src || !row_idx) return ht::bad_arg("ht_copy_rows_h2d: NULL pointer");
    if (dst_pitch < width_bytes || src_pitch < width_bytes) return ht::bad_arg("ht_copy_rows_h2d: pitch smaller than width");
    int dev = 0;
    HT_CUDA_TRY(cudaGetDevice(&dev));
    bool done = false;  // page-locked source: the SMs gather the rows over PCIe, asynchronously
    if (int rc = ht::gather_rows_from_pinned(dst, dst_pitch, src, src_pitch, width_bytes, row_idx, n_rows, (cudaStream_t)stream, &done))
        return rc;
    if (done) return HT_OK;
    if ((size_t)width_bytes > ht::hostcopy::kSlotBytes || dev >= ht::hostcopy::kMaxDevs) {
        ht::set_error("ht_copy_rows_h2d: rows of %lld bytes do not fit the staging slots", (long long)width_bytes);
        return HT_ERR_UNSUPPORTED;
    }
    return ht::staged_copy(dst, dst_pitch, src, src_pitch, width_bytes, n_rows, (cudaStream_t)stream, row_idx);
}

int ht_host_register(void* ptr, size_t bytes) {
    if (!ptr || bytes == 0) return ht::bad_arg("ht_host_register: empty range");
    HT_CUDA_TRY(cudaHostRegister(ptr, bytes, cudaHostRegisterDefault));
    return HT_OK;
}

int ht_host_unregister(void* ptr) {
    if (!ptr) return ht::bad_arg("ht_host_unregister: NULL");
    HT_CUDA_TRY(cudaHostUnregister(ptr));
    return HT_OK;
}

}  // extern "C"
