// C-ABI glue of libhaptools_cuda: version, thread-local error string, workspace sizes and a
// few host<->device copy helpers.  See include/haptools_cuda.h.
#include <stdarg.h>
#include <stdio.h>

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
    if (dst_pitch == width_bytes && src_pitch == width_bytes) {  // dense on both sides: one linear copy
        HT_CUDA_TRY(cudaMemcpyAsync(dst, src, (size_t)width_bytes * (size_t)height, k, (cudaStream_t)stream));
        return HT_OK;
    }
    HT_CUDA_TRY(cudaMemcpy2DAsync(dst, (size_t)dst_pitch, src, (size_t)src_pitch,
                                  (size_t)width_bytes, (size_t)height, k,
                                  (cudaStream_t)stream));
    return HT_OK;
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
