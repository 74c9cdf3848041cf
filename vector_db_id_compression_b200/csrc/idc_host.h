// idc_host.h -- host-side plumbing shared by the C-ABI translation units:
// error reporting, the context object, device buffers, launch accounting.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "../../include/idcodec.h"

#include <chrono>
#include <cstdlib>

namespace idc {

void set_error(const char* fmt, ...);

// IDC_TRACE_HOST=1: wall-clock time of the host-side phases of a call on stderr (where the milliseconds of a call
// with a million tiny units go)
struct HostTrace {
    bool on;
    const char* what;
    std::chrono::steady_clock::time_point t0;
    explicit HostTrace(const char* w) : on(getenv("IDC_TRACE_HOST") != nullptr), what(w), t0(std::chrono::steady_clock::now()) {}
    void mark(const char* phase) {
        if (!on) return;
        auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[idc host] %s: %-28s %8.3f ms\n", what, phase, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

#define IDC_CUDA(call)                                                                          \
    do {                                                                                        \
        cudaError_t e_ = (call);                                                                \
        if (e_ != cudaSuccess) {                                                                \
            idc::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return IDC_ERR_CUDA;                                                                \
        }                                                                                       \
    } while (0)

#define IDC_TRY(call)        \
    do {                     \
        int r_ = (call);     \
        if (r_ != IDC_OK)    \
            return r_;       \
    } while (0)

#define IDC_REQUIRE(cond, code, ...)     \
    do {                                 \
        if (!(cond)) {                   \
            idc::set_error(__VA_ARGS__); \
            return (code);               \
        }                                \
    } while (0)

// grow-only device allocation
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    ~DevBuf() { release(); }
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    int reserve(size_t bytes) {
        if (bytes <= cap) return IDC_OK;
        release();
        if (bytes == 0) return IDC_OK;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) {
            p = nullptr;
            set_error("cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
            return IDC_ERR_NOMEM;
        }
        cap = bytes;
        return IDC_OK;
    }
    template <typename T>
    T* as() const { return reinterpret_cast<T*>(p); }
};

struct KernelTime {
    const char* name;
    cudaEvent_t a, b;
};

}  // namespace idc

struct idc_ctx {
    // One call at a time per context: workspaces, staging buffers and the stream are shared state. The Faiss
    // virtuals (get_ids ...) are called from OpenMP threads concurrently, so every entry point takes this lock.
    std::mutex mu;
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::atomic<int> live_blobs{0};  // blobs created by this context and not freed yet (idc_ctx_destroy refuses while > 0)
    size_t l2_fetch_saved = 0;       // cudaLimitMaxL2FetchGranularity before this context changed it (0: untouched)
    int sm_count = 148;
    uint64_t launches = 0;
    bool timing = false;
    std::vector<idc::KernelTime> times;      // events of the current/last call
    std::vector<cudaEvent_t> event_pool;
    size_t events_used = 0;
    // constant tables (device)
    uint32_t* d_mt = nullptr;     // first kMtWords outputs of std::mt19937(1234)
    uint64_t* d_rcp64 = nullptr;  // floor((2^64-1)/d), d = 0..65536
    uint32_t* d_q31 = nullptr;    // 2^31 / d
    uint64_t* d_binom = nullptr;  // C(n, k), n, k < 64 (row-major 64 x 64), then 64 field widths: the RRR(63) block coder's
                                  // tables, allocated by the first wt_type = 1 call (wt_kernels.cu)
    // grow-only scratch reused across calls
    idc::DevBuf ws;       // order-statistic workspaces
    idc::DevBuf scratch;  // encoder word scratch / staging
    idc::DevBuf stage;    // host<->device staging of ids
    idc::DevBuf meta;     // per-call unit tables
    idc::DevBuf status;   // per-call status words
    // device memory pool for blob arrays: cudaMalloc/cudaFree of GB-sized blocks cost tens of ms each, and a
    // blob is typically rebuilt with the same shapes; freed blocks are kept and reused (exact size class).
    std::vector<std::pair<void*, size_t>> pool_free;
    std::vector<std::pair<void*, size_t>> pool_live;
    int pool_alloc(void** p, size_t bytes);
    void pool_release(void* p);
    void pool_trim();
    // auxiliary streams: size classes of one logical kernel run concurrently (fork/join around c->stream)
    std::vector<cudaStream_t> aux;
    std::vector<cudaEvent_t> aux_done;
    cudaEvent_t fork_ev = nullptr;
    // copy engine stream + reusable ordering events (host<->device copies overlapped with the kernels)
    cudaStream_t copy_stream = nullptr, copy_stream2 = nullptr;
    std::vector<cudaEvent_t> sync_events;
    size_t sync_used = 0;
    int copy_stream_get(cudaStream_t* s, int which = 0);
    // Mailbox of the small synchronous calls (one graph row, a handful of rows): pinned host memory the device addresses
    // directly. The kernel reads the row numbers from it and stores its output into it, so such a call is one launch and
    // one stream synchronisation -- no copy calls (each costs as much as the decode of a row). IDC_NO_MAILBOX=1: off.
    static constexpr size_t kMailboxBytes = 64u << 10;
    void* mailbox = nullptr;      // host address
    void* mailbox_dev = nullptr;  // the same bytes as the device sees them
    bool mailbox_off = false;
    int mailbox_get(size_t bytes, void** host, void** dev);  // *host == nullptr: not available for this size
    int sync_event(cudaEvent_t* e);  // an event for ordering only; recycled at the next begin_call()
    int fork(int n);              // make aux[0..n) wait for everything queued on `stream`
    int join(int n);              // make `stream` wait for aux[0..n)

    void begin_call();
    void mark(const char* name);   // call right before a kernel launch
    void mark_end();               // call right after it
};

namespace idc {

// member of every blob: keeps the owning context's live-blob count
struct CtxRef {
    idc_ctx* c = nullptr;
    void bind(idc_ctx* ctx) {
        c = ctx;
        c->live_blobs++;
    }
    ~CtxRef() {
        if (c) c->live_blobs--;
    }
};

// RAII helper: times one kernel launch when ctx->timing is on, counts it always
struct LaunchScope {
    idc_ctx* c;
    LaunchScope(idc_ctx* ctx, const char* name) : c(ctx) { c->mark(name); }
    ~LaunchScope() { c->mark_end(); }
};

int check_last_launch(const char* what);

}  // namespace idc
