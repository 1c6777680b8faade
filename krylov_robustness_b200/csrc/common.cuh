// common.cuh - context, error handling, device-memory pool, launch accounting.
// Part of libkrylov_b200 (sm_100a only).  No CPU fallback anywhere: every failure surfaces as a
// negative kr_status with a thread-local message.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <unordered_map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/krylov_b200.h"

namespace kr {

// ---------------------------------------------------------------------------------- errors
inline std::string& tls_error() {
    static thread_local std::string e;
    return e;
}

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

[[noreturn]] inline void fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    throw Error(code, buf);
}

#define KR_CUDA(call)                                                                           \
    do {                                                                                        \
        cudaError_t e_ = (call);                                                                \
        if (e_ != cudaSuccess)                                                                  \
            ::kr::fail(KR_ERR_CUDA, "CUDA error %s at %s:%d: %s", cudaGetErrorName(e_), __FILE__, \
                       __LINE__, cudaGetErrorString(e_));                                       \
    } while (0)
// wraps an extern "C" body: exceptions -> status + message
template <class F>
int guarded(F&& f) {
    try {
        f();
        return KR_OK;
    } catch (const Error& e) {
        tls_error() = e.what();
        return e.code;
    } catch (const std::bad_alloc&) {
        tls_error() = "host allocation failed";
        return KR_ERR_NOMEM;
    } catch (const std::exception& e) {
        tls_error() = e.what();
        return KR_ERR_ARG;
    }
}

}  // namespace kr

// ---------------------------------------------------------------------------------- context
struct kr_ctx {
    int device = 0;
    int num_sms = 148;
    size_t l2_bytes = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;   // host<->device staging that overlaps compute (created on first use)
    // counters: launches, spmm launches, matvecs, h2d bytes, d2h bytes
    int64_t counters[5] = {0, 0, 0, 0, 0};
    // optional SpMM timing
    bool timing = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> spmm_events;
    std::vector<cudaEvent_t> free_events;   // recycled timing events (no cudaEventCreate inside a timed loop)
    void timing_events(cudaEvent_t* a, cudaEvent_t* b) {
        cudaEvent_t* out[2] = {a, b};
        for (auto* o : out) {
            if (!free_events.empty()) { *o = free_events.back(); free_events.pop_back(); }
            else if (cudaEventCreate(o) != cudaSuccess) throw std::runtime_error("cudaEventCreate failed");
        }
    }
    double spmm_ms = 0.0;
    int64_t spmm_timed = 0;
    // size-bucketed free lists (bytes rounded up to a power of two >= 512)
    std::multimap<size_t, void*> pool;
    size_t pool_bytes = 0;
    std::unordered_map<void*, size_t> live;   // actual size of every block handed out (a request may get a larger one)
    // result of the last kr_fun_update, kept on the device until fetched
    struct FunUpdateResult* last_fun_update = nullptr;

    void* alloc(size_t bytes);
    void release(void* p, size_t bytes);
    void trim();
};

namespace kr {

inline size_t bucket(size_t bytes) {
    size_t b = 512;
    while (b < bytes) b <<= 1;
    // beyond 64 MiB round to 16 MiB multiples instead of powers of two (180 GB is not infinite)
    if (b > (size_t(64) << 20)) {
        const size_t g = size_t(16) << 20;
        b = (bytes + g - 1) / g * g;
    }
    return b;
}

// RAII device buffer drawing from the context pool.
template <class T>
struct DevBuf {
    kr_ctx* ctx = nullptr;
    T* p = nullptr;
    size_t count = 0;
    DevBuf() = default;
    DevBuf(kr_ctx* c, size_t n) { reset(c, n); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : ctx(o.ctx), p(o.p), count(o.count) { o.p = nullptr; o.count = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) {
            free();
            ctx = o.ctx; p = o.p; count = o.count;
            o.p = nullptr; o.count = 0;
        }
        return *this;
    }
    ~DevBuf() { free(); }
    void reset(kr_ctx* c, size_t n) {
        free();
        ctx = c;
        count = n;
        p = n ? static_cast<T*>(c->alloc(n * sizeof(T))) : nullptr;
    }
    void free() {
        if (p) ctx->release(p, count * sizeof(T));
        p = nullptr;
        count = 0;
    }
    size_t bytes() const { return count * sizeof(T); }
    void zero() { if (p) KR_CUDA(cudaMemsetAsync(p, 0, bytes(), ctx->stream)); }
    void upload(const T* h, size_t n) {
        KR_CUDA(cudaMemcpyAsync(p, h, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
        ctx->counters[3] += (int64_t)(n * sizeof(T));
    }
    void download(T* h, size_t n) const {
        KR_CUDA(cudaMemcpyAsync(h, p, n * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
        KR_CUDA(cudaStreamSynchronize(ctx->stream));
        ctx->counters[4] += (int64_t)(n * sizeof(T));
    }
    std::vector<T> to_host() const {
        std::vector<T> h(count);
        if (count) download(h.data(), count);
        return h;
    }
};

inline void check_launch(kr_ctx* ctx, const char* name) {
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
        cudaGetLastError();
        fail(KR_ERR_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(e));
    }
    ctx->counters[0] += 1;
}

#define KR_LAUNCH(ctx, kernel, grid, block, smem, ...)                          \
    do {                                                                        \
        kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);        \
        ::kr::check_launch((ctx), #kernel);                                     \
    } while (0)

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace kr

inline void* kr_ctx::alloc(size_t bytes) {
    size_t b = kr::bucket(bytes);
    // debugging aid (KR_POOL_POISON=1): every buffer handed out is filled with NaN bit patterns first, so a kernel
    // that reads memory it (or a predecessor) never wrote turns the result into NaN instead of a plausible number
    static const bool poison = getenv("KR_POOL_POISON") != nullptr;
    // best fit: the smallest cached block that holds the request, as long as it wastes less than half of itself
    // (an exact-size pool fills up with blocks of sizes that never come back: the candidate pipeline, the SLQ pass and
    // the node bases of one bench run all use different multi-GB shapes)
    auto it = pool.lower_bound(b);
    if (it != pool.end() && it->first <= b + b / 2) {
        void* p = it->second;
        const size_t actual = it->first;
        pool.erase(it);
        pool_bytes -= actual;
        live[p] = actual;
        if (poison) cudaMemsetAsync(p, 0xFF, b, stream);
        return p;
    }
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, b);
    if (e != cudaSuccess) {
        cudaGetLastError();
        trim();
        e = cudaMalloc(&p, b);
        if (e != cudaSuccess) {
            cudaGetLastError();
            kr::fail(KR_ERR_NOMEM, "device allocation of %zu bytes failed: %s", b, cudaGetErrorString(e));
        }
    }
    live[p] = b;
    if (poison) cudaMemsetAsync(p, 0xFF, b, stream);
    return p;
}

inline void kr_ctx::release(void* p, size_t bytes) {
    size_t b = kr::bucket(bytes);
    auto lv = live.find(p);
    if (lv != live.end()) { b = lv->second; live.erase(lv); }
    // keep at most 96 GiB cached (a B200 has 180 GB; the candidate-pair path cycles three ~17 GB blocks per
    // chunk and must not pay cudaFree + cudaMalloc for one of them on every chunk).  Over the cap the LARGEST cached
    // blocks go back to the driver first - the block being released is the most recently used shape and the most
    // likely to be asked for again.  alloc() trims the whole pool and retries when the driver runs out.
    const size_t cap = size_t(96) << 30;
    if (b > cap) {
        cudaStreamSynchronize(stream);
        cudaFree(p);
        return;
    }
    if (pool_bytes + b > cap) {
        cudaStreamSynchronize(stream);
        while (!pool.empty() && pool_bytes + b > cap) {
            auto last = std::prev(pool.end());
            cudaFree(last->second);
            pool_bytes -= last->first;
            pool.erase(last);
        }
    }
    pool.emplace(b, p);
    pool_bytes += b;
}

inline void kr_ctx::trim() {
    cudaStreamSynchronize(stream);
    for (auto& kv : pool) cudaFree(kv.second);
    pool.clear();
    pool_bytes = 0;
}
