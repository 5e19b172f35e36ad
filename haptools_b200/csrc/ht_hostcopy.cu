// Host copy engine of libhaptools_cuda: a persistent pool of host threads that moves data between PAGEABLE
// host memory (what `gts.data` is when it comes out of cyvcf2 / pgenlib / numpy, and what the np.empty result
// of haptools/data/haplotypes.py:1408 is) and the device through page-locked staging slots, and that widens the
// bit-packed result to the reference's bool bytes on the way down (ht_expand_host.cpp).
//
//   * one pool per process, started on first use (HT_HOST_COPY_THREADS workers, default min(cores / ranks on this host, 12)); every
//     worker owns two page-locked slots and one copy stream per device, so a worker always has one DMA in flight
//     while it copies / expands the previous block
//   * a job = one call (upload, download, download + expand) cut into row blocks; workers take blocks of the
//     first job whose device-side precondition (an event recorded on the caller's stream at submit) has been
//     met, so jobs for different devices, or an upload queued behind a download that still waits for its
//     kernels, do not hold each other up
//   * every job is asynchronous to the caller: ht_host_submit_* returns a ticket at once, ht_host_wait(ticket)
//     returns when the job's data is where it belongs (and, for uploads, is visible to any stream).  The
//     synchronous entry points (ht_copy2d, ht_copy_rows_h2d) are submit + wait.
// Pure data movement: no genotype byte is interpreted on the host.
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <unistd.h>

#include <atomic>
#include <condition_variable>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "ht_common.cuh"

namespace ht {

void expand_bits_rows(const uint8_t* bits, int64_t bits_pitch, uint8_t* out, int64_t out_pitch, int64_t n_out, int64_t rows);
int expand_isa(int force);

namespace hostcopy {

constexpr size_t kSlotBytes = 8u << 20;
constexpr int kSlots = 2;
constexpr int kMaxWorkers = 64;
constexpr int kMaxDevs = 16;

enum Kind { kUpload = 0, kDownload = 1, kDownloadExpand = 2 };

struct Job {
    int kind = 0, dev = 0;
    uint8_t* dst = nullptr;
    const uint8_t* src = nullptr;
    int64_t dst_pitch = 0, src_pitch = 0, width = 0, height = 0, rows_per_block = 1, n_blocks = 0;
    int64_t out_width = 0;              // kDownloadExpand: bytes of an expanded row
    std::vector<uint32_t> row_idx;      // kUpload: gather of selected source rows (empty: rows 0 .. height-1)
    cudaEvent_t ready = nullptr;        // recorded on the caller's stream at submit
    bool ready_seen = false;            // (under the pool mutex) the event has been seen complete
    int64_t next_block = 0;             // (under the pool mutex)
    std::atomic<int64_t> done_blocks{0};
    std::atomic<int> err{0};            // first cudaError_t
    int64_t ticket = 0;
    bool finished = false;              // (under the pool mutex)
};

struct Block {
    std::shared_ptr<Job> job;
    int64_t r0 = 0, rows = 0;
    int slot = 0;
};

struct Worker {
    uint8_t* slot[kSlots] = {};
    cudaEvent_t ev[kMaxDevs][kSlots] = {};
    int slot_dev[kSlots] = {-1, -1};  // device of the slot's last DMA that may still be in flight
    cudaStream_t stream[kMaxDevs] = {};
    int cur_dev = -1;
};

class Pool {
  public:
    static Pool& get() {
        static Pool* p = new Pool();  // never destroyed: worker threads may outlive static destruction order
        return *p;
    }

    int n_workers() const { return n_workers_; }

    int64_t submit(std::shared_ptr<Job> job, cudaStream_t user_stream) {
        // the device-side precondition of the job: everything queued on the caller's stream so far
        cudaError_t e = cudaEventCreateWithFlags(&job->ready, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventRecord(job->ready, user_stream);
        if (e != cudaSuccess) {
            if (job->ready) cudaEventDestroy(job->ready);
            cuda_fail(e, "host copy engine: recording the job's event");
            return -1;
        }
        std::unique_lock<std::mutex> lk(mu_);
        start_workers_locked();
        job->ticket = ++last_ticket_;
        jobs_[job->ticket] = job;
        queue_.push_back(job);
        lk.unlock();
        cv_work_.notify_all();
        return job->ticket;
    }

    // 0 = everything submitted so far
    int wait(int64_t ticket) {
        std::unique_lock<std::mutex> lk(mu_);
        int rc = HT_OK;
        auto reap = [&](std::map<int64_t, std::shared_ptr<Job>>::iterator it) {
            std::shared_ptr<Job> j = it->second;
            cv_done_.wait(lk, [&] { return j->finished; });
            const int err = j->err.load();
            if (j->ready) {
                cudaEventDestroy(j->ready);
                j->ready = nullptr;
            }
            if (err) rc = cuda_fail((cudaError_t)err, "host copy engine (staged copy through page-locked slots)");
        };
        if (ticket == 0) {
            while (!jobs_.empty()) {
                auto it = jobs_.begin();
                reap(it);
                jobs_.erase(it->first);
            }
            return rc;
        }
        auto it = jobs_.find(ticket);
        if (it == jobs_.end()) return HT_OK;  // waited for already
        reap(it);
        jobs_.erase(ticket);
        return rc;
    }

  private:
    Pool() {
        const char* e = getenv("HT_HOST_COPY_THREADS");
        int v = e ? atoi(e) : 0;
        const int hw = (int)std::thread::hardware_concurrency();
        if (v <= 0) {
            // default: min(cores, 12), shared fairly when a launcher runs one process per GPU on this host
            // (torchrun exports LOCAL_WORLD_SIZE): 8 ranks x 12 workers on 32 cores only fight each other
            const char* lws = getenv("LOCAL_WORLD_SIZE");
            const int ranks = lws ? atoi(lws) : 1;
            int share = hw > 0 ? hw / (ranks > 0 ? ranks : 1) : 4;
            if (share < 3) share = 3;
            v = share < 12 ? share : 12;
        }
        if (hw > 0 && v > hw) v = hw;
        n_workers_ = v < 1 ? 1 : (v > kMaxWorkers ? kMaxWorkers : v);
    }

    void start_workers_locked() {
        if (started_) return;
        started_ = true;
        for (int t = 0; t < n_workers_; ++t) std::thread(&Pool::run, this, t).detach();
    }

    // next block of the first job whose precondition has been met (else of the first job); `block` = wait for work
    bool claim(Block& b, bool block) {
        std::unique_lock<std::mutex> lk(mu_);
        while (true) {
            while (!queue_.empty() && queue_.front()->next_block >= queue_.front()->n_blocks) queue_.pop_front();
            std::shared_ptr<Job> pick;
            for (auto& j : queue_) {
                if (j->next_block >= j->n_blocks) continue;
                if (!j->ready_seen && cudaEventQuery(j->ready) != cudaErrorNotReady) j->ready_seen = true;  // done or failed
                if (j->ready_seen) {
                    pick = j;
                    break;
                }
            }
            if (!pick)
                for (auto& j : queue_)
                    if (j->next_block < j->n_blocks) {
                        pick = j;
                        break;
                    }
            if (pick) {
                const int64_t i = pick->next_block++;
                b.job = pick;
                b.r0 = i * pick->rows_per_block;
                b.rows = pick->height - b.r0 < pick->rows_per_block ? pick->height - b.r0 : pick->rows_per_block;
                return true;
            }
            if (!block) return false;
            cv_work_.wait(lk);
        }
    }

    void block_done(const Block& b, cudaError_t e) {
        Job& j = *b.job;
        if (e != cudaSuccess) {
            int expected = 0;
            j.err.compare_exchange_strong(expected, (int)e);
        }
        if (j.done_blocks.fetch_add(1) + 1 == j.n_blocks) {
            std::lock_guard<std::mutex> lk(mu_);
            j.finished = true;
            cv_done_.notify_all();
        }
    }

    static cudaError_t prepare(Worker& w, int dev) {
        cudaError_t e = cudaSuccess;
        if (w.cur_dev != dev) {
            e = cudaSetDevice(dev);
            if (e != cudaSuccess) return e;
            w.cur_dev = dev;
        }
        if (!w.stream[dev]) e = cudaStreamCreateWithFlags(&w.stream[dev], cudaStreamNonBlocking);
        for (int s = 0; s < kSlots && e == cudaSuccess; ++s) {
            if (!w.slot[s]) e = cudaHostAlloc((void**)&w.slot[s], kSlotBytes, cudaHostAllocPortable);
            if (e == cudaSuccess && !w.ev[dev][s]) e = cudaEventCreateWithFlags(&w.ev[dev][s], cudaEventDisableTiming);
        }
        return e;
    }

    // first half of a block: the slot is free, the DMA (and for uploads the host copy before it) is under way
    static cudaError_t start(Worker& w, Block& b) {
        Job& j = *b.job;
        if (j.dev < 0 || j.dev >= kMaxDevs) return cudaErrorInvalidDevice;
        cudaError_t e = prepare(w, j.dev);
        if (e != cudaSuccess) return e;
        const int s = b.slot;
        if (w.slot_dev[s] >= 0) {
            e = cudaEventSynchronize(w.ev[w.slot_dev[s]][s]);
            w.slot_dev[s] = -1;
            if (e != cudaSuccess) return e;
        }
        cudaStream_t st = w.stream[j.dev];
        e = cudaStreamWaitEvent(st, j.ready, 0);
        if (e != cudaSuccess) return e;
        uint8_t* stage = w.slot[s];
        if (j.kind == kUpload) {
            const uint8_t* src = j.src + b.r0 * j.src_pitch;
            if (!j.row_idx.empty()) {
                for (int64_t r = 0; r < b.rows; ++r)
                    memcpy(stage + r * j.width, j.src + (int64_t)j.row_idx[(size_t)(b.r0 + r)] * j.src_pitch, (size_t)j.width);
            } else if (j.src_pitch == j.width) {
                memcpy(stage, src, (size_t)(b.rows * j.width));
            } else {
                for (int64_t r = 0; r < b.rows; ++r) memcpy(stage + r * j.width, src + r * j.src_pitch, (size_t)j.width);
            }
            e = cudaMemcpy2DAsync(j.dst + b.r0 * j.dst_pitch, (size_t)j.dst_pitch, stage, (size_t)j.width, (size_t)j.width,
                                  (size_t)b.rows, cudaMemcpyHostToDevice, st);
        } else {
            e = cudaMemcpy2DAsync(stage, (size_t)j.width, j.src + b.r0 * j.src_pitch, (size_t)j.src_pitch, (size_t)j.width,
                                  (size_t)b.rows, cudaMemcpyDeviceToHost, st);
        }
        if (e == cudaSuccess) e = cudaEventRecord(w.ev[j.dev][s], st);
        if (e == cudaSuccess) w.slot_dev[s] = j.dev;
        return e;
    }

    // second half: the DMA has landed; downloads leave the slot for the caller's memory
    static cudaError_t finish(Worker& w, Block& b) {
        Job& j = *b.job;
        const int s = b.slot;
        cudaError_t e = cudaEventSynchronize(w.ev[j.dev][s]);
        w.slot_dev[s] = -1;
        if (e != cudaSuccess || j.kind == kUpload) return e;
        const uint8_t* stage = w.slot[s];
        uint8_t* d = j.dst + b.r0 * j.dst_pitch;
        if (j.kind == kDownloadExpand) {
            expand_bits_rows(stage, j.width, d, j.dst_pitch, j.out_width, b.rows);
        } else if (j.dst_pitch == j.width) {
            memcpy(d, stage, (size_t)(b.rows * j.width));
        } else {
            for (int64_t r = 0; r < b.rows; ++r) memcpy(d + r * j.dst_pitch, stage + r * j.width, (size_t)j.width);
        }
        return cudaSuccess;
    }

    void run(int /*t*/) {
        Worker w;
        Block pending;
        bool has_pending = false;
        int slot = 0;
        while (true) {
            Block nb;
            const bool got = claim(nb, !has_pending);
            cudaError_t e_start = cudaSuccess;
            if (got) {
                nb.slot = slot;
                e_start = start(w, nb);
            }
            if (has_pending) {
                block_done(pending, finish(w, pending));
                has_pending = false;
                pending.job.reset();
            }
            if (got) {
                if (e_start != cudaSuccess) {
                    cudaGetLastError();
                    block_done(nb, e_start);
                } else {
                    pending = nb;
                    has_pending = true;
                    slot ^= 1;
                }
            }
        }
    }

    std::mutex mu_;
    std::condition_variable cv_work_, cv_done_;
    std::deque<std::shared_ptr<Job>> queue_;
    std::map<int64_t, std::shared_ptr<Job>> jobs_;
    int64_t last_ticket_ = 0;
    int n_workers_ = 4;
    bool started_ = false;
};

static void hint_huge_pages(void* p, size_t bytes) {
    // a fresh np.empty result is first touched by the workers: with transparent huge pages in "madvise" mode this
    // turns 4 KB faults into 2 MB ones.  A hint; failure is fine.
    static const long page = sysconf(_SC_PAGESIZE);
    cudaPointerAttributes at;  // page-locked memory is populated already
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) cudaGetLastError();
    else if (at.type != cudaMemoryTypeUnregistered) return;
    const uintptr_t huge = 2u << 20;
    uintptr_t a = ((uintptr_t)p + huge - 1) & ~(huge - 1), b = ((uintptr_t)p + bytes) & ~(huge - 1);
    if (page > 0 && b > a) madvise((void*)a, b - a, MADV_HUGEPAGE);
}

int64_t submit(int kind, void* dst, int64_t dst_pitch, const void* src, int64_t src_pitch, int64_t width, int64_t height,
               int64_t out_width, const uint32_t* row_idx, cudaStream_t st) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev >= kMaxDevs) {
        set_error("host copy engine: no current device (or device index >= %d)", kMaxDevs);
        return -1;
    }
    if ((size_t)width > kSlotBytes) {
        set_error("host copy engine: rows of %lld bytes do not fit a staging slot (%zu bytes)", (long long)width, kSlotBytes);
        return -1;
    }
    auto job = std::make_shared<Job>();
    job->kind = kind;
    job->dev = dev;
    job->dst = static_cast<uint8_t*>(dst);
    job->src = static_cast<const uint8_t*>(src);
    job->dst_pitch = dst_pitch;
    job->src_pitch = src_pitch;
    job->width = width;
    job->height = height;
    job->out_width = out_width;
    if (row_idx) job->row_idx.assign(row_idx, row_idx + height);
    // block size: whole rows, at most a slot; small enough that every worker gets a few blocks of a big job
    // (an expanded block writes 8 x its bytes, so those are cut finer: 512 KB of bits = 4 MB of result; 16,384 samples
    // of the UKB plan end to end: 39.4 ms against 42.8 ms with 1 MB blocks and 48 ms with 128 KB; upload blocks
    // below 4 MB only lose, 64.8 ms at 1 MB for a pageable source -- profiles/r2/host_blocks.json)
    size_t target = kind == kDownloadExpand ? (512u << 10) : (4u << 20);
    if (const char* e = getenv(kind == kDownloadExpand ? "HT_HOST_EXPAND_BLOCK_KB" : "HT_HOST_BLOCK_KB"))  // measurements
        if (atol(e) > 0) target = (size_t)atol(e) << 10;
    int64_t rpb = (int64_t)(target / (size_t)width);
    const int64_t max_rpb = (int64_t)(kSlotBytes / (size_t)width);
    if (rpb < 1) rpb = 1;
    if (rpb > max_rpb) rpb = max_rpb;
    job->rows_per_block = rpb;
    job->n_blocks = (height + rpb - 1) / rpb;
    if (kind != kUpload && (size_t)height * (size_t)(kind == kDownloadExpand ? out_width : width) >= (32u << 20))
        hint_huge_pages(dst, (size_t)(height - 1) * (size_t)dst_pitch + (size_t)(kind == kDownloadExpand ? out_width : width));
    return Pool::get().submit(job, st);
}

int wait(int64_t ticket) { return Pool::get().wait(ticket); }
int workers() { return Pool::get().n_workers(); }

}  // namespace hostcopy

// ---------------------------------------------------------------------------------------
// bit-packed result (tiles, H, 32 words, 2 strands) -> one bit string per sample
// ---------------------------------------------------------------------------------------
// rowbits[s][j / 8] bit j % 8 = result byte j = 2 h + c of sample s: exactly the bytes of the reference's
// hap_gts.data[s] at 1 bit each, so the host only widens (ht_expand_host.cpp).
// CTA = (tile, 512 haplotypes); warp q = words 4q .. 4q+3 (one 32-byte sector of every record); lane = 16
// haplotypes: their 32 words (16 x 2 strands) of one 32-sample word are transposed in registers into one uint32
// per sample, and a warp store writes 128 contiguous bytes of one sample's bit string.
__global__ void __launch_bounds__(256)
bits_to_rowbits_kernel(const uint32_t* __restrict__ bits, int64_t n_samples, int64_t n_haps, uint32_t n_hap_blocks,
                       uint8_t* __restrict__ rowbits, int64_t pitch) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t tile = blockIdx.x / n_hap_blocks;
    const int64_t hb = blockIdx.x % n_hap_blocks;
    const int64_t h0 = hb * 512 + 16 * (int64_t)lane;
    if (h0 >= n_haps) return;
    const uint2* rec = reinterpret_cast<const uint2*>(bits) + ((size_t)tile * (size_t)n_haps + (size_t)h0) * kTileWords;
    const int n_my = (int)min((int64_t)16, n_haps - h0);
#pragma unroll 1
    for (int wi = 0; wi < 4; ++wi) {
        const int w = 4 * warp + wi;
        const int64_t s0 = tile * kTileSamples + 32 * (int64_t)w;
        if (s0 >= n_samples) break;
        uint32_t a[32];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            uint2 v = make_uint2(0u, 0u);
            if (i < n_my) v = __ldg(rec + (size_t)i * kTileWords + w);
            a[2 * i] = v.x;
            a[2 * i + 1] = v.y;
        }
        transpose32(a);  // a[b] = the 32 result bits (16 haplotypes x 2 strands) of sample s0 + b
        const int rows = (int)min((int64_t)32, n_samples - s0);
        uint8_t* p = rowbits + s0 * pitch + 128 * hb + 4 * lane;
#pragma unroll
        for (int b = 0; b < 32; ++b)
            if (b < rows) *reinterpret_cast<uint32_t*>(p + (int64_t)b * pitch) = a[b];
    }
}

}  // namespace ht

extern "C" {

int64_t ht_rowbits_pitch(int64_t n_haps) {
    if (n_haps <= 0) return 0;
    return ((n_haps + 15) / 16 * 4 + 15) / 16 * 16;  // one uint32 per 16 haplotypes, rows 16-byte aligned
}

int ht_bits_to_rowbits(const uint32_t* bits, int64_t n_samples, int64_t n_haps, uint8_t* rowbits, int64_t pitch,
                       void* stream) {
    using namespace ht;
    if (n_samples < 0 || n_haps < 0) return bad_arg("negative size");
    if (n_samples == 0 || n_haps == 0) return HT_OK;
    if (!bits || !rowbits) return bad_arg("bits / rowbits is NULL");
    if ((reinterpret_cast<uintptr_t>(bits) & 7) || (reinterpret_cast<uintptr_t>(rowbits) & 3) || (pitch & 3))
        return bad_arg("bits must be 8-byte aligned, rowbits and its pitch 4-byte aligned");
    if (pitch < (n_haps + 15) / 16 * 4) return bad_arg("pitch < 4 * ceil(n_haps / 16)");
    const int64_t n_hb = (n_haps + 511) / 512, tiles = ht_num_tiles(n_samples);
    if (n_hb > 0x7fffffff) return bad_arg("too many haplotypes");
    const int64_t per_launch = max((int64_t)1, (int64_t)0x7fffffff / n_hb);
    for (int64_t t0 = 0; t0 < tiles; t0 += per_launch) {
        const int64_t nt = min(per_launch, tiles - t0);
        const int64_t ns = min(n_samples - t0 * kTileSamples, nt * kTileSamples);
        bits_to_rowbits_kernel<<<(unsigned)(nt * n_hb), 256, 0, (cudaStream_t)stream>>>(
            bits + (size_t)t0 * (size_t)n_haps * kRecordWords, ns, n_haps, (uint32_t)n_hb,
            rowbits + (size_t)t0 * kTileSamples * (size_t)pitch, pitch);
        HT_CUDA_TRY(cudaGetLastError());
    }
    return HT_OK;
}

int ht_expand_rowbits_host(const uint8_t* h_rowbits, int64_t bits_pitch, int64_t n_rows, int64_t n_out_bytes,
                           uint8_t* h_out, int64_t out_pitch) {
    if (n_rows < 0 || n_out_bytes < 0) return ht::bad_arg("negative size");
    if (n_rows == 0 || n_out_bytes == 0) return HT_OK;
    if (!h_rowbits || !h_out) return ht::bad_arg("NULL pointer");
    if (bits_pitch < (n_out_bytes + 7) / 8 || out_pitch < n_out_bytes) return ht::bad_arg("pitch smaller than a row");
    ht::expand_bits_rows(h_rowbits, bits_pitch, h_out, out_pitch, n_out_bytes, n_rows);
    return HT_OK;
}

int ht_host_expand_isa(int force) { return ht::expand_isa(force); }

int64_t ht_host_submit_upload(void* dst, int64_t dst_pitch, const void* h_src, int64_t src_pitch, int64_t width_bytes,
                              int64_t height, const uint32_t* h_row_idx, void* stream) {
    if (width_bytes <= 0 || height <= 0 || !dst || !h_src || dst_pitch < width_bytes || src_pitch < width_bytes) {
        ht::bad_arg("ht_host_submit_upload: empty extent, NULL pointer or pitch smaller than width");
        return -1;
    }
    return ht::hostcopy::submit(ht::hostcopy::kUpload, dst, dst_pitch, h_src, src_pitch, width_bytes, height, 0, h_row_idx,
                                (cudaStream_t)stream);
}

int64_t ht_host_submit_download(void* h_dst, int64_t dst_pitch, const void* src, int64_t src_pitch, int64_t width_bytes,
                                int64_t height, void* stream) {
    if (width_bytes <= 0 || height <= 0 || !h_dst || !src || dst_pitch < width_bytes || src_pitch < width_bytes) {
        ht::bad_arg("ht_host_submit_download: empty extent, NULL pointer or pitch smaller than width");
        return -1;
    }
    return ht::hostcopy::submit(ht::hostcopy::kDownload, h_dst, dst_pitch, src, src_pitch, width_bytes, height, 0, nullptr,
                                (cudaStream_t)stream);
}

int64_t ht_host_submit_download_expand(void* h_out, int64_t out_pitch, int64_t n_out_bytes, const void* rowbits,
                                       int64_t bits_pitch, int64_t n_rows, void* stream) {
    const int64_t width = (n_out_bytes + 7) / 8;
    if (n_out_bytes <= 0 || n_rows <= 0 || !h_out || !rowbits || out_pitch < n_out_bytes || bits_pitch < width) {
        ht::bad_arg("ht_host_submit_download_expand: empty extent, NULL pointer or pitch smaller than a row");
        return -1;
    }
    return ht::hostcopy::submit(ht::hostcopy::kDownloadExpand, h_out, out_pitch, rowbits, bits_pitch, width, n_rows,
                                n_out_bytes, nullptr, (cudaStream_t)stream);
}

int ht_host_wait(int64_t ticket) {
    if (ticket < 0) return ht::bad_arg("ht_host_wait: negative ticket");
    return ht::hostcopy::wait(ticket);
}

int ht_host_workers(void) { return ht::hostcopy::workers(); }

int ht_host_is_pageable(const void* h_ptr) {
    if (!h_ptr) return 0;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, h_ptr) != cudaSuccess) {
        cudaGetLastError();  // old drivers report unregistered memory as an error
        return 1;
    }
    return at.type == cudaMemoryTypeUnregistered ? 1 : 0;
}

}  // extern "C"
