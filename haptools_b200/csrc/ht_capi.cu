// C-ABI glue of libhaptools_cuda: version, thread-local error string, workspace sizes and a
// few host<->device copy helpers.  See include/haptools_cuda.h.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <thread>
#include <vector>

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
// Upload from PAGEABLE host memory (what `gts.data` is when it comes out of cyvcf2 / pgenlib /
// numpy).  cudaMemcpy2DAsync stages such a source through the driver's own bounce buffer with one
// host thread (11 GB/s on the B200 box); here a few host threads copy row blocks into a ring of
// page-locked bounce buffers and every block leaves with an asynchronous DMA, so the upload runs
// at what memcpy on several cores and the PCIe link deliver (HT_HOST_COPY_THREADS, default 4; every
// thread owns two 8 MB bounce buffers and whole row blocks).  Pure data movement: no byte is
// interpreted on the host.
// ---------------------------------------------------------------------------------------
namespace {
constexpr int kMaxCopyThreads = 8, kBufsPerThread = 2;
constexpr size_t kBounceBytes = 8u << 20;
constexpr int kMaxBounceDevs = 16;
struct Bounce {
    std::mutex mu;  // one staged upload at a time per process
    uint8_t* buf[kMaxCopyThreads][kBufsPerThread] = {};
    cudaEvent_t ev[kMaxBounceDevs][kMaxCopyThreads][kBufsPerThread] = {};  // an event belongs to a device
    int last_dev[kMaxCopyThreads][kBufsPerThread];                          // device of the buffer's last DMA, -1: none
    Bounce() {
        for (auto& row : last_dev)
            for (int& v : row) v = -1;
    }
};
Bounce g_bounce;

int host_threads() {
    static const int n = [] {
        const char* e = getenv("HT_HOST_COPY_THREADS");
        int v = e ? atoi(e) : 4;
        const int hw = (int)std::thread::hardware_concurrency();
        if (hw > 0 && v > hw) v = hw;
        return v < 1 ? 1 : (v > kMaxCopyThreads ? kMaxCopyThreads : v);
    }();
    return n;
}

// worker t of nt: row blocks t, t + nt, ... through its own pair of bounce buffers
// `row_idx` (optional): output row r is source row row_idx[r] (a gather of selected rows)
void upload_blocks(int t, int nt, int dev, uint8_t* dst, int64_t dst_pitch, const uint8_t* src, int64_t src_pitch,
                   int64_t width, int64_t height, int64_t rows_per_block, const uint32_t* row_idx, cudaStream_t st,
                   cudaError_t* err) {
    cudaError_t e = t ? cudaSetDevice(dev) : cudaSuccess;  // a new host thread starts on device 0
    int b = 0;
    for (int64_t r0 = t * rows_per_block; e == cudaSuccess && r0 < height; r0 += nt * rows_per_block, b ^= 1) {
        const int64_t rows = height - r0 < rows_per_block ? height - r0 : rows_per_block;
        const int ld = g_bounce.last_dev[t][b];
        if (ld >= 0) e = cudaEventSynchronize(g_bounce.ev[ld][t][b]);  // its last DMA (on any device) has left
        if (e != cudaSuccess) break;
        uint8_t* stage = g_bounce.buf[t][b];
        const uint8_t* s = src + r0 * src_pitch;
        if (row_idx) {
            for (int64_t r = 0; r < rows; ++r) memcpy(stage + r * width, src + (int64_t)row_idx[r0 + r] * src_pitch, (size_t)width);
        } else if (src_pitch == width) {
            memcpy(stage, s, (size_t)(rows * width));
        } else {
            for (int64_t r = 0; r < rows; ++r) memcpy(stage + r * width, s + r * src_pitch, (size_t)width);
        }
        e = cudaMemcpy2DAsync(dst + r0 * dst_pitch, (size_t)dst_pitch, stage, (size_t)width, (size_t)width, (size_t)rows,
                              cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaEventRecord(g_bounce.ev[dev][t][b], st);
        g_bounce.last_dev[t][b] = e == cudaSuccess ? dev : -1;
    }
    *err = e;
}
// worker t of nt of a download into pageable memory: row blocks t, t + nt, ...; the DMA of a block into one
// bounce buffer is in flight while the previous block is copied out of the other one
void download_blocks(int t, int nt, int dev, uint8_t* dst, int64_t dst_pitch, const uint8_t* src, int64_t src_pitch,
                     int64_t width, int64_t height, int64_t rows_per_block, cudaStream_t st, cudaError_t* err) {
    cudaError_t e = t ? cudaSetDevice(dev) : cudaSuccess;
    auto drain = [&](int64_t r0, int64_t rows, int b) {  // block (r0, rows) has been asked into buffer b
        e = cudaEventSynchronize(g_bounce.ev[dev][t][b]);
        if (e != cudaSuccess) return;
        const uint8_t* stage = g_bounce.buf[t][b];
        uint8_t* d = dst + r0 * dst_pitch;
        if (dst_pitch == width) {
            memcpy(d, stage, (size_t)(rows * width));
        } else {
            for (int64_t r = 0; r < rows; ++r) memcpy(d + r * dst_pitch, stage + r * width, (size_t)width);
        }
        g_bounce.last_dev[t][b] = -1;  // free, nothing pending
    };
    int b = 0;
    int64_t prev_r0 = -1, prev_rows = 0;
    for (int64_t r0 = t * rows_per_block; e == cudaSuccess && r0 < height; r0 += nt * rows_per_block, b ^= 1) {
        const int64_t rows = height - r0 < rows_per_block ? height - r0 : rows_per_block;
        const int ld = g_bounce.last_dev[t][b];
        if (ld >= 0) e = cudaEventSynchronize(g_bounce.ev[ld][t][b]);  // an earlier upload out of this buffer
        if (e != cudaSuccess) break;
        e = cudaMemcpy2DAsync(g_bounce.buf[t][b], (size_t)width, src + r0 * src_pitch, (size_t)src_pitch, (size_t)width,
                              (size_t)rows, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaEventRecord(g_bounce.ev[dev][t][b], st);
        if (e != cudaSuccess) break;
        g_bounce.last_dev[t][b] = dev;
        if (prev_r0 >= 0) drain(prev_r0, prev_rows, b ^ 1);
        prev_r0 = r0;
        prev_rows = rows;
    }
    if (e == cudaSuccess && prev_r0 >= 0) drain(prev_r0, prev_rows, b ^ 1);
    *err = e;
}
}  // namespace

// staged copy between PAGEABLE host memory and the device; `download`: device -> host (returns when the
// destination holds the data), else host -> device (returns when the source has been read)
static int h2d_from_pageable(void* dst, int64_t dst_pitch, const void* src, int64_t src_pitch, int64_t width,
                             int64_t height, cudaStream_t st, const uint32_t* row_idx = nullptr, bool download = false) {
    std::lock_guard<std::mutex> lock(g_bounce.mu);
    int dev = 0;
    HT_CUDA_TRY(cudaGetDevice(&dev));
    const int64_t rows_per_block = (int64_t)(kBounceBytes / (size_t)width);
    const int64_t n_blocks = (height + rows_per_block - 1) / rows_per_block;
    const int nt = (int)(n_blocks < host_threads() ? n_blocks : host_threads());
    for (int t = 0; t < nt; ++t) {
        for (int b = 0; b < kBufsPerThread; ++b) {
            if (!g_bounce.buf[t][b])
                HT_CUDA_TRY(cudaHostAlloc((void**)&g_bounce.buf[t][b], kBounceBytes, cudaHostAllocPortable));
            if (!g_bounce.ev[dev][t][b])
                HT_CUDA_TRY(cudaEventCreateWithFlags(&g_bounce.ev[dev][t][b], cudaEventDisableTiming));
        }
    }
    cudaError_t errs[kMaxCopyThreads] = {};  // cudaSuccess
    std::vector<std::thread> th;
    th.reserve(nt > 1 ? nt - 1 : 0);
    uint8_t* d = static_cast<uint8_t*>(dst);
    const uint8_t* s = static_cast<const uint8_t*>(src);
    if (download) {
        for (int t = 1; t < nt; ++t)
            th.emplace_back(download_blocks, t, nt, dev, d, dst_pitch, s, src_pitch, width, height, rows_per_block, st, &errs[t]);
        download_blocks(0, nt, dev, d, dst_pitch, s, src_pitch, width, height, rows_per_block, st, &errs[0]);
    } else {
        for (int t = 1; t < nt; ++t)
            th.emplace_back(upload_blocks, t, nt, dev, d, dst_pitch, s, src_pitch, width, height, rows_per_block, row_idx, st, &errs[t]);
        upload_blocks(0, nt, dev, d, dst_pitch, s, src_pitch, width, height, rows_per_block, row_idx, st, &errs[0]);
    }
    for (auto& x : th) x.join();
    for (int t = 0; t < nt; ++t)
        if (errs[t] != cudaSuccess) return cuda_fail(errs[t], "staged copy through the bounce buffers");
    return HT_OK;
}

// ---------------------------------------------------------------------------------------
// Row gather straight out of PAGE-LOCKED host memory: the SMs read the selected rows over PCIe (mapped
// host memory, unified addressing) and write them to the device buffer -- no host thread touches a byte
// and the call is asynchronous on the stream.  A warp per row, 16-byte loads, four in flight per lane.
// ---------------------------------------------------------------------------------------
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
        for (; i + 96 < n16; i += 128) {
            const uint4 v0 = s[i], v1 = s[i + 32], v2 = s[i + 64], v3 = s[i + 96];
            d[i] = v0;
            d[i + 32] = v1;
            d[i + 64] = v2;
            d[i + 96] = v3;
        }
        for (; i < n16; i += 32) d[i] = s[i];
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
        const unsigned grid = (unsigned)(ctas < 148 * 8 ? ctas : 148 * 8);
        gather_rows_kernel<<<grid, 256, 0, st>>>(static_cast<uint8_t*>(dst), dst_pitch, dsrc, src_pitch, width, d_idx, n_rows);
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
    if (kind == HT_COPY_H2D && width_bytes * height >= (4 << 20) && (size_t)width_bytes <= ht::kBounceBytes &&
        cudaGetDevice(&cur_dev) == cudaSuccess && cur_dev < ht::kMaxBounceDevs) {
        cudaPointerAttributes at;
        const cudaError_t e = cudaPointerGetAttributes(&at, src);
        if (e != cudaSuccess) cudaGetLastError();  // old drivers report unregistered memory as an error
        if (e != cudaSuccess || at.type == cudaMemoryTypeUnregistered)
            return ht::h2d_from_pageable(dst, dst_pitch, src, src_pitch, width_bytes, height, (cudaStream_t)stream);
    }
    if (kind == HT_COPY_D2H && width_bytes * height >= (4 << 20) && (size_t)width_bytes <= ht::kBounceBytes &&
        cudaGetDevice(&cur_dev) == cudaSuccess && cur_dev < ht::kMaxBounceDevs) {
        cudaPointerAttributes at;
        const cudaError_t e = cudaPointerGetAttributes(&at, dst);
        if (e != cudaSuccess) cudaGetLastError();
        if (e != cudaSuccess || at.type == cudaMemoryTypeUnregistered)
            return ht::h2d_from_pageable(dst, dst_pitch, src, src_pitch, width_bytes, height, (cudaStream_t)stream, nullptr, true);
    }
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
    if ((size_t)width_bytes > ht::kBounceBytes || dev >= ht::kMaxBounceDevs) {
        ht::set_error("ht_copy_rows_h2d: rows of %lld bytes do not fit the bounce buffers", (long long)width_bytes);
        return HT_ERR_UNSUPPORTED;
    }
    return ht::h2d_from_pageable(dst, dst_pitch, src, src_pitch, width_bytes, n_rows, (cudaStream_t)stream, row_idx);
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
