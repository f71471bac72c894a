// Shared host-side plumbing for libsafe_b200: error convention, context, device buffers.
#pragma once
#include <cuda_runtime.h>

#include <chrono>
#include <cstdarg>
#include <cstdlib>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/safe_b200.h"

namespace sb {

// thread-local last-error string (sb_last_error)
std::string& last_error();
void set_error(const char* fmt, ...);

struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};

[[noreturn]] void fail(const char* fmt, ...);

#define SB_CUDA(expr)                                                                        \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess)                                                               \
            ::sb::fail("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
    } while (0)

#define SB_CHECK(cond, ...)                      \
    do {                                         \
        if (!(cond)) ::sb::fail(__VA_ARGS__);    \
    } while (0)

// Wrap the body of every extern "C" entry point.
#define SB_API_BEGIN try {
#define SB_API_END                                   \
    return 0;                                        \
    }                                                \
    catch (const std::exception& ex) {               \
        ::sb::set_error("%s", ex.what());            \
        return 1;                                    \
    }                                                \
    catch (...) {                                    \
        ::sb::set_error("unknown C++ exception");    \
        return 2;                                    \
    }

// Device allocations are stream-ordered and cached: cudaMallocAsync / cudaFreeAsync on the stream of the context
// the calling thread bound last (sb_ctx::bind), from the device's default memory pool whose release threshold
// sb_ctx_create raises to "keep everything", behind a size-matched block cache (neigh.cu).  A plan that is created
// and destroyed per call therefore costs no allocator calls after the first call.
cudaStream_t& alloc_stream();  // thread-local
void* dev_alloc(size_t bytes);
void dev_free(void* p);
void dev_cache_trim(int device);  // return the parked blocks of one device to the driver

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    void release() {
        if (p) dev_free(p);
        p = nullptr;
        n = 0;
    }
    // grow-only allocation (contents are NOT preserved)
    void reserve(size_t count) {
        if (count <= n) return;
        release();
        p = static_cast<T*>(dev_alloc(count * sizeof(T)));
        n = count;
    }
};

}  // namespace sb

// kernel classes whose device time can be accumulated with CUDA events (sb_ctx_profile / sb_ctx_kernel_ms)
enum { SB_K_GEMM = 0, SB_K_GATHER = 1, SB_K_FIXUP = 2, SB_K_SSSP = 3, SB_K_EUCLID = 4, SB_K_HYPERGEOM = 5,
       SB_K_SCORE = 6, SB_K_PREP = 7, SB_K_TAIL = 8, SB_K_FDR = 9, SB_K_JACCARD = 10, SB_K_CLASSES = 11 };

struct sb_ctx {
    int device = 0;
    int num_sms = 0;
    cudaStream_t stream = nullptr;
    int64_t launches = 0;
    bool profile = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> timers[SB_K_CLASSES];
    // grow-only scratch shared by all plans of this context (used only inside one call at a time):
    // gathered operand tiles, fix-up list, packed count accumulators
    sb::DevBuf<int8_t> ws_bcat;      // gathered operand tiles of the batch in flight
    sb::DevBuf<uint64_t> ws_flag_ij;
    sb::DevBuf<uint32_t> ws_flag_p;
    sb::DevBuf<uint32_t> ws_cpk;
    sb::DevBuf<unsigned int> ws_counter;
    size_t bcat_budget = 0;  // 1/4 of the free device memory seen at the first null of this context
    void bind() const {
        SB_CUDA(cudaSetDevice(device));
        sb::alloc_stream() = stream;
    }
};

namespace sb {
// RAII: when profiling is on, brackets the launches issued in its scope with an event pair on the ctx stream
struct KernelTimer {
    sb_ctx* ctx;
    int cls;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaStream_t st;
    KernelTimer(sb_ctx* c, int k, cudaStream_t on = nullptr) : ctx(c), cls(k), st(on ? on : c->stream) {
        if (!ctx->profile) return;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0, st);
    }
    ~KernelTimer() {
        if (!e0) return;
        cudaEventRecord(e1, st);
        ctx->timers[cls].emplace_back(e0, e1);
    }
};
}  // namespace sb

namespace sb {
// SB_TRACE=1: print host-side phase timings (each phase is closed by a stream synchronize) to stderr
struct PhaseTrace {
    sb_ctx* ctx;
    const char* name;
    int mode;  // 0 off, 1 SB_TRACE (phases closed by a stream synchronize), 2 SB_TRACE_HOST (host wall time only)
    std::chrono::steady_clock::time_point t0;
    PhaseTrace(sb_ctx* c, const char* n) : ctx(c), name(n) {
        static const int enabled = getenv("SB_TRACE") ? 1 : (getenv("SB_TRACE_HOST") ? 2 : 0);
        mode = enabled;
        if (mode) {
            if (mode == 1) cudaStreamSynchronize(ctx->stream);
            t0 = std::chrono::steady_clock::now();
        }
    }
    ~PhaseTrace() {
        if (!mode) return;
        if (mode == 1) cudaStreamSynchronize(ctx->stream);
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (mode == 1 || ms > 5.0) fprintf(stderr, "[sb_trace] %-28s %9.3f ms\n", name, ms);
    }
};
}  // namespace sb

#define SB_LAUNCH_CHECK(ctx)            \
    do {                                \
        (ctx)->launches++;              \
        SB_CUDA(cudaGetLastError());    \
    } while (0)

struct sb_neigh {
    sb_ctx* ctx = nullptr;
    int64_t n = 0;
    int64_t ld = 0;  // words per row
    uint32_t* words = nullptr;
    bool owned = false;
};

namespace sb {
// Device -> pageable host copy of a large result array (finalize.cu): pieces go through a pinned ring at full PCIe
// speed and worker threads move them into the caller's buffer, so that the first-touch page faults of a freshly
// allocated destination are taken in parallel.  Returns after the whole copy has landed.
void copy_out(sb_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes);
// The other direction: worker threads fill the pinned ring from pageable memory, the caller's thread uploads the
// slots in order.  Both fall back to one cudaMemcpyAsync for small or already pinned buffers.
void copy_in(sb_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes);
}  // namespace sb

static inline int64_t sb_ld_words(int64_t n) { return ((n + 31) / 32 + 3) / 4 * 4; }
static inline int64_t sb_ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
