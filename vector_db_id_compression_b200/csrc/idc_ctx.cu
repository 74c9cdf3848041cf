// idc_ctx.cu -- context, error reporting and constant tables of libidcodec.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <random>

#include "idc_core.cuh"
#include "idc_host.h"

namespace idc {

static thread_local std::string g_last_error;

void set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
}

int check_last_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("kernel launch %s failed: %s", what, cudaGetErrorString(e));
        return IDC_ERR_CUDA;
    }
    return IDC_OK;
}

}  // namespace idc

void idc_ctx::begin_call() {
    times.clear();
    events_used = 0;
    sync_used = 0;
}

void idc_ctx::mark(const char* name) {
    launches++;
    if (!timing) return;
    while (event_pool.size() < events_used + 2) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        event_pool.push_back(e);
    }
    idc::KernelTime kt{name, event_pool[events_used], event_pool[events_used + 1]};
    events_used += 2;
    cudaEventRecord(kt.a, stream);
    times.push_back(kt);
}

void idc_ctx::mark_end() {
    if (!timing || times.empty()) return;
    cudaEventRecord(times.back().b, stream);
}

int idc_ctx::pool_alloc(void** p, size_t bytes) {
    *p = nullptr;
    bytes = (std::max<size_t>(bytes, 1) + 511) & ~size_t(511);
    size_t best = pool_free.size();
    for (size_t i = 0; i < pool_free.size(); i++)
        if (pool_free[i].second >= bytes && pool_free[i].second <= bytes + bytes / 8 + 4096 &&
            (best == pool_free.size() || pool_free[i].second < pool_free[best].second))
            best = i;
    if (best != pool_free.size()) {
        *p = pool_free[best].first;
        pool_live.push_back(pool_free[best]);
        pool_free.erase(pool_free.begin() + best);
        return IDC_OK;
    }
    cudaError_t e = cudaMalloc(p, bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        pool_trim();  // give cached blocks back and retry once
        e = cudaMalloc(p, bytes);
    }
    if (e != cudaSuccess) {
        *p = nullptr;
        idc::set_error("cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
        return IDC_ERR_NOMEM;
    }
    pool_live.push_back({*p, bytes});
    return IDC_OK;
}

void idc_ctx::pool_release(void* p) {
    if (!p) return;
    for (size_t i = 0; i < pool_live.size(); i++)
        if (pool_live[i].first == p) {
            pool_free.push_back(pool_live[i]);
            pool_live.erase(pool_live.begin() + i);
            if (pool_free.size() > 64) pool_trim();
            return;
        }
    cudaFree(p);  // not ours
}

void idc_ctx::pool_trim() {
    for (auto& b : pool_free) cudaFree(b.first);
    pool_free.clear();
}

int idc_ctx::copy_stream_get(cudaStream_t* s, int which) {
    cudaStream_t& cs = which ? copy_stream2 : copy_stream;
    if (!cs) IDC_CUDA(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
    *s = cs;
    return IDC_OK;
}

int idc_ctx::mailbox_get(size_t bytes, void** host, void** dev) {
    *host = *dev = nullptr;
    if (mailbox_off || bytes > kMailboxBytes) return IDC_OK;
    if (!mailbox) {
        void* h = nullptr;
        void* d = nullptr;
        if (cudaHostAlloc(&h, kMailboxBytes, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess ||
            cudaHostGetDevicePointer(&d, h, 0) != cudaSuccess) {
            // no mapped host memory on this platform: the copy path does the same job
            if (h) cudaFreeHost(h);
            cudaGetLastError();
            mailbox_off = true;
            return IDC_OK;
        }
        mailbox = h;
        mailbox_dev = d;
    }
    *host = mailbox;
    *dev = mailbox_dev;
    return IDC_OK;
}

int idc_ctx::sync_event(cudaEvent_t* e) {
    if (sync_used == sync_events.size()) {
        cudaEvent_t ev;
        IDC_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        sync_events.push_back(ev);
    }
    *e = sync_events[sync_used++];
    return IDC_OK;
}

int idc_ctx::fork(int n) {
    if (!fork_ev) IDC_CUDA(cudaEventCreateWithFlags(&fork_ev, cudaEventDisableTiming));
    while ((int)aux.size() < n) {
        cudaStream_t s;
        cudaEvent_t e;
        // aux[0] carries the longest units, whose serial chains bound the kernel: highest priority first
        int least = 0, greatest = 0;
        cudaDeviceGetStreamPriorityRange(&least, &greatest);
        // (streams past the third carry pieces of the longest class: highest priority again)
        int prio = aux.size() >= 3 ? greatest : std::min(least, greatest + (int)aux.size());
        IDC_CUDA(cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, prio));
        IDC_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        aux.push_back(s);
        aux_done.push_back(e);
    }
    IDC_CUDA(cudaEventRecord(fork_ev, stream));
    for (int i = 0; i < n; i++) IDC_CUDA(cudaStreamWaitEvent(aux[i], fork_ev, 0));
    return IDC_OK;
}

int idc_ctx::join(int n) {
    for (int i = 0; i < n; i++) {
        IDC_CUDA(cudaEventRecord(aux_done[i], aux[i]));
        IDC_CUDA(cudaStreamWaitEvent(stream, aux_done[i], 0));
    }
    return IDC_OK;
}

extern "C" {

const char* idc_last_error(void) { return idc::g_last_error.c_str(); }

int idc_version(void) { return 100; }

int idc_ctx_create_on_stream(int device, void* cuda_stream, idc_ctx** out) {
    IDC_REQUIRE(out != nullptr, IDC_ERR_ARG, "idc_ctx_create: out is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        idc::set_error("no usable CUDA device (%s): this library has no CPU fallback",
                       e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return IDC_ERR_CUDA;
    }
    IDC_REQUIRE(device >= 0 && device < ndev, IDC_ERR_ARG, "device %d out of range (have %d)", device, ndev);
    IDC_CUDA(cudaSetDevice(device));
    idc_ctx* c = new idc_ctx();
    c->device = device;
    if (cuda_stream == (void*)-1) {
        IDC_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        c->own_stream = true;
    } else {
        c->stream = (cudaStream_t)cuda_stream;
        c->own_stream = false;
    }
    cudaDeviceProp prop;
    IDC_CUDA(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    c->mailbox_off = getenv("IDC_NO_MAILBOX") != nullptr || !prop.canMapHostMemory;
    // The ROC kernels read isolated 32-byte sectors (tree nodes, bucket records) scattered over GBs: L2 can be asked
    // not to promote such misses to 64/128-byte DRAM fetches (measured 3 sectors fetched per sector used otherwise).
    // That limit is DEVICE-GLOBAL -- it changes the behaviour of every other kernel of the process (Faiss, torch) --
    // so it is opt-in: IDC_L2_FETCH=32|64|128 sets it for the lifetime of the context, idc_ctx_destroy puts the
    // previous value back. bench.py opts in and says so in its config.
    if (const char* e = getenv("IDC_L2_FETCH")) {
        const size_t gran = (size_t)atoi(e);
        size_t before = 0;
        if ((gran == 32 || gran == 64 || gran == 128) && cudaDeviceGetLimit(&before, cudaLimitMaxL2FetchGranularity) == cudaSuccess &&
            cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran) == cudaSuccess)
            c->l2_fetch_saved = before;
        cudaGetLastError();
    }

    // mt19937(1234): the fallback word source of ANSState::stack_slice (codec.h:16-18,32-40)
    uint32_t mt[idc::kMtWords];
    {
        std::mt19937 g(1234);
        for (int i = 0; i < idc::kMtWords; i++) mt[i] = (uint32_t)g();
    }
    IDC_CUDA(cudaMalloc(&c->d_mt, sizeof(mt)));
    IDC_CUDA(cudaMemcpy(c->d_mt, mt, sizeof(mt), cudaMemcpyHostToDevice));

    // reciprocal tables for the uniform coder (codec.cpp:21-63), indexed by nmax <= 65536
    const uint32_t N = idc::kMaxUnit + 1;
    std::vector<uint64_t> rcp(N);
    std::vector<uint32_t> q31(N);
    rcp[0] = 0;
    q31[0] = 0;
    for (uint32_t d = 1; d < N; d++) {
        rcp[d] = ~0ull / d;
        q31[d] = (uint32_t)((1ull << 31) / d);
    }
    IDC_CUDA(cudaMalloc(&c->d_rcp64, N * sizeof(uint64_t)));
    IDC_CUDA(cudaMalloc(&c->d_q31, N * sizeof(uint32_t)));
    IDC_CUDA(cudaMemcpy(c->d_rcp64, rcp.data(), N * sizeof(uint64_t), cudaMemcpyHostToDevice));
    IDC_CUDA(cudaMemcpy(c->d_q31, q31.data(), N * sizeof(uint32_t), cudaMemcpyHostToDevice));
    *out = c;
    return IDC_OK;
}

int idc_ctx_create(int device, idc_ctx** out) { return idc_ctx_create_on_stream(device, (void*)-1, out); }

int idc_ctx_destroy(idc_ctx* c) {
    if (!c) return IDC_OK;
    // blobs give their arrays back to the context's pool when they are freed: the context must outlive them
    IDC_REQUIRE(c->live_blobs.load() == 0, IDC_ERR_ARG,
                "idc_ctx_destroy: %d blob(s) of this context are still alive; free them first (the context was left intact)",
                c->live_blobs.load());
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->l2_fetch_saved) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, c->l2_fetch_saved);
    for (auto e : c->event_pool) cudaEventDestroy(e);
    for (auto s : c->aux) cudaStreamDestroy(s);
    for (auto e : c->aux_done) cudaEventDestroy(e);
    if (c->fork_ev) cudaEventDestroy(c->fork_ev);
    for (auto e : c->sync_events) cudaEventDestroy(e);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->copy_stream2) cudaStreamDestroy(c->copy_stream2);
    cudaFree(c->d_mt);
    cudaFree(c->d_rcp64);
    cudaFree(c->d_q31);
    cudaFree(c->d_binom);
    if (c->mailbox) cudaFreeHost(c->mailbox);
    c->pool_trim();
    for (auto& b : c->pool_live) cudaFree(b.first);
    c->pool_live.clear();
    c->ws.release();
    c->scratch.release();
    c->stage.release();
    c->meta.release();
    c->status.release();
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
    return IDC_OK;
}

int idc_ctx_synchronize(idc_ctx* c) {
    IDC_REQUIRE(c != nullptr, IDC_ERR_ARG, "ctx is NULL");
    IDC_CUDA(cudaStreamSynchronize(c->stream));
    return IDC_OK;
}

uint64_t idc_ctx_launch_count(const idc_ctx* c) { return c ? c->launches : 0; }

int idc_ctx_set_timing(idc_ctx* c, int enable) {
    IDC_REQUIRE(c != nullptr, IDC_ERR_ARG, "ctx is NULL");
    c->timing = enable != 0;
    return IDC_OK;
}

float idc_ctx_last_kernel_ms(const idc_ctx* c) {
    if (!c) return 0.f;
    float total = 0.f;
    for (auto& kt : c->times) {
        float ms = 0.f;
        if (cudaEventSynchronize(kt.b) == cudaSuccess && cudaEventElapsedTime(&ms, kt.a, kt.b) == cudaSuccess)
            total += ms;
    }
    return total;
}

int idc_ctx_last_kernel_breakdown(const idc_ctx* c, const char** names, float* ms, int cap) {
    if (!c) return 0;
    int n = 0;
    for (auto& kt : c->times) {
        if (n >= cap) break;
        float t = 0.f;
        cudaEventSynchronize(kt.b);
        cudaEventElapsedTime(&t, kt.a, kt.b);
        names[n] = kt.name;
        ms[n] = t;
        n++;
    }
    return n;
}

}  // extern "C"
