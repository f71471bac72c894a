// Stage 1: neighborhood construction -> bit-packed N x N matrix.
// Replaces SAFE.define_neighborhoods (reference safepy/safe.py:369-430).
//
//   k_euclid   : tiled all-pairs distance-and-threshold; one __ballot_sync per 32 pairs produces the packed word.
//   k_sssp     : batched multi-source label-correcting frontier search on CSR with the radius cutoff applied
//                in-kernel; fp64 distances, 64-bit atomicMin on the IEEE bit pattern, warp-aggregated
//                frontier compaction; the source's row lives as a bitmap in shared memory and is stored once.
#include <cooperative_groups.h>

#include <mutex>
#include <unordered_map>
#include <vector>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace sb {

std::string& last_error() {
    thread_local std::string s;
    return s;
}
void set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    last_error() = buf;
}
cudaStream_t& alloc_stream() {
    thread_local cudaStream_t s = nullptr;
    return s;
}
// Block cache in front of cudaMallocAsync.  The driver's pool was observed to stall for 400-600 ms now and then when
// a plan's large buffers (hundreds of MB) were re-requested right after having been freed (pool growth / remapping),
// so freed blocks are parked here, keyed by (device, stream), and handed back on a size match: every
// compute_pvalues call asks for the same sizes, so steady state makes no allocator calls at all.  Reuse is safe
// because allocation and release are both ordered on the same stream.
namespace {
struct CachedBlock {
    int device;
    cudaStream_t stream;
    void* p;
    size_t bytes;
};
std::mutex g_cache_mu;
std::vector<CachedBlock> g_cache;
std::unordered_map<void*, size_t> g_live;  // size of every block handed out
size_t g_cached_bytes = 0;
constexpr size_t kCacheLimit = 96ull << 30;
}  // namespace

void* dev_alloc(size_t bytes) {
    if (bytes == 0) bytes = 1;
    int device = 0;
    cudaGetDevice(&device);
    cudaStream_t st = alloc_stream();
    {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        int best = -1;
        for (int i = 0; i < static_cast<int>(g_cache.size()); ++i) {
            const CachedBlock& b = g_cache[i];
            if (b.device != device || b.stream != st || b.bytes < bytes || b.bytes > bytes + bytes / 4 + 4096) continue;
            if (best < 0 || b.bytes < g_cache[best].bytes) best = i;
        }
        if (best >= 0) {
            void* p = g_cache[best].p;
            g_cached_bytes -= g_cache[best].bytes;
            g_live[p] = g_cache[best].bytes;
            g_cache.erase(g_cache.begin() + best);
            return p;
        }
    }
    void* p = nullptr;
    cudaError_t e = cudaMallocAsync(&p, bytes, st);
    if (e != cudaSuccess) {
        cudaGetLastError();
        dev_cache_trim(device);  // give the parked blocks back and retry once
        e = cudaMallocAsync(&p, bytes, st);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        fail("device allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
    }
    std::lock_guard<std::mutex> lk(g_cache_mu);
    g_live[p] = bytes;
    return p;
}
void dev_free(void* p) {
    if (!p) return;
    int device = 0;
    cudaGetDevice(&device);
    std::lock_guard<std::mutex> lk(g_cache_mu);
    auto it = g_live.find(p);
    const size_t bytes = it == g_live.end() ? 0 : it->second;
    if (it != g_live.end()) g_live.erase(it);
    if (bytes == 0 || g_cached_bytes + bytes > kCacheLimit) {
        cudaFreeAsync(p, alloc_stream());
        return;
    }
    g_cache.push_back(CachedBlock{device, alloc_stream(), p, bytes});
    g_cached_bytes += bytes;
}
void dev_cache_trim(int device) {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    for (size_t i = 0; i < g_cache.size();) {
        if (g_cache[i].device == device) {
            cudaFreeAsync(g_cache[i].p, g_cache[i].stream);
            g_cached_bytes -= g_cache[i].bytes;
            g_cache.erase(g_cache.begin() + i);
        } else {
            ++i;
        }
    }
}
void fail(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    throw Error(buf);
}

// ------------------------------------------------------------------------------------------------ K2 euclid
// Block = 8 warps; warp w owns a 256-column strip (8 packed words), lane l holds columns strip + l + 32q.
// The block walks a tile of EU_ROWS source rows whose coordinates sit in shared memory (broadcast reads).
// Per row a warp issues 8 ballots -> 8 consecutive words -> one full 32-byte sector store.
// fp64 ops are spelled with _rn intrinsics so that nvcc cannot contract them into FMAs: scipy's pdist
// evaluates sqrt(dx*dx + dy*dy) unfused, and `sqrt(s) < nr` is decided as `s <= s_star` with s_star the
// largest double whose correctly rounded sqrt is < nr (computed by the host wrapper; sqrt_rn is monotone).
constexpr int EU_ROWS = 128;
constexpr int EU_WARPS = 8;

__global__ void __launch_bounds__(EU_WARPS * 32) k_euclid(const double* __restrict__ x, const double* __restrict__ y,
                                                            int64_t n, double s_star, int any_pair, int64_t row0,
                                                            int64_t row1, uint32_t* __restrict__ words, int64_t ld) {
    __shared__ double sx[EU_ROWS], sy[EU_ROWS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t r_base = row0 + static_cast<int64_t>(blockIdx.y) * EU_ROWS;
    const int64_t w_base = (static_cast<int64_t>(blockIdx.x) * EU_WARPS + warp) * 8;  // first word of the strip
    for (int i = threadIdx.x; i < EU_ROWS; i += blockDim.x) {
        int64_t r = r_base + i;
        sx[i] = r < row1 ? x[r] : 0.0;
        sy[i] = r < row1 ? y[r] : 0.0;
    }
    __syncthreads();
    if (w_base >= ld) return;
    double cx[8], cy[8];
    bool valid[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        int64_t t = (w_base + q) * 32 + lane;
        valid[q] = t < n && any_pair;
        cx[q] = valid[q] ? x[t] : 0.0;
        cy[q] = valid[q] ? y[t] : 0.0;
    }
    const int rows = static_cast<int>(min(static_cast<int64_t>(EU_ROWS), row1 - r_base));
    for (int i = 0; i < rows; ++i) {
        const double xs = sx[i], ys = sy[i];
        uint32_t mine = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            double dx = __dsub_rn(xs, cx[q]);
            double dy = __dsub_rn(ys, cy[q]);
            double s = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
            uint32_t b = __ballot_sync(0xffffffffu, valid[q] && s <= s_star);
            if (lane == q) mine = b;
        }
        if (lane < 8 && w_base + lane < ld) words[(r_base + i) * ld + w_base + lane] = mine;
    }
}

// ------------------------------------------------------------------------------------------------ K1 sssp
constexpr int SP_THREADS = 256;
constexpr int SP_GROUP_DEFAULT = 16;  // lanes cooperating on one frontier node's adjacency list (measured: 16 lanes
                                      // and 8 CTAs per SM run 1.6x faster than 8 lanes and 4 CTAs; SB_SSSP_* override)
constexpr unsigned long long SP_INF = 0x7FF0000000000000ull;

struct SsspWs {
    unsigned long long* dist;  // [grid][n] fp64 bit patterns, +inf when untouched
    uint32_t* stamp;           // [grid][n] level stamp of the last enqueue
    int32_t* queue;            // [grid][2][n]
    unsigned int* next_row;    // dynamic source counter
};

template <int SP_GROUP>
__global__ void __launch_bounds__(SP_THREADS) k_sssp(const int64_t* __restrict__ indptr,
                                                      const int32_t* __restrict__ indices,
                                                      const double* __restrict__ length, int64_t n, double cutoff,
                                                      int64_t row0, int64_t row1, SsspWs ws,
                                                      uint32_t* __restrict__ words, int64_t ld) {
    extern __shared__ uint32_t s_row[];  // ld words: the current source's packed row
    __shared__ int s_cnt[2];
    __shared__ unsigned int s_src;
    __shared__ uint32_t s_stamp;

    unsigned long long* dist = ws.dist + static_cast<size_t>(blockIdx.x) * n;
    uint32_t* stamp = ws.stamp + static_cast<size_t>(blockIdx.x) * n;
    int32_t* q0 = ws.queue + static_cast<size_t>(blockIdx.x) * 2 * n;
    int32_t* q1 = q0 + n;

    for (int64_t w = threadIdx.x; w < ld; w += blockDim.x) s_row[w] = 0;
    if (threadIdx.x == 0) s_stamp = 0;
    __syncthreads();

    const int group = threadIdx.x / SP_GROUP, gl = threadIdx.x % SP_GROUP;
    const int ngroups = blockDim.x / SP_GROUP;

    while (true) {
        if (threadIdx.x == 0) s_src = atomicAdd(ws.next_row, 1u);
        __syncthreads();
        const int64_t src = row0 + s_src;
        if (src >= row1) break;
        if (threadIdx.x == 0) {
            dist[src] = 0ull;  // the source's distance is exactly 0 (networkx pushes (0, source))
            q0[0] = static_cast<int32_t>(src);
            s_cnt[0] = 1;
            s_cnt[1] = 0;
            s_row[src >> 5] |= 1u << (src & 31);
        }
        __syncthreads();
        int cur = 0;
        while (true) {
            const int nf = s_cnt[cur];
            if (nf == 0) break;
            const uint32_t level_stamp = s_stamp + 1;
            const int32_t* qc = cur ? q1 : q0;
            int32_t* qn = cur ? q0 : q1;
            int* cnt_next = &s_cnt[cur ^ 1];
            for (int f = group; f < nf; f += ngroups) {
                const int32_t u = qc[f];
                const double du = __longlong_as_double(static_cast<long long>(
                    *reinterpret_cast<volatile unsigned long long*>(&dist[u])));
                const int64_t e1 = indptr[u + 1];
                for (int64_t e = indptr[u] + gl; e < e1; e += SP_GROUP) {
                    const int32_t v = indices[e];
                    const double w = length ? length[e] : 1.0;
                    const double nd = __dadd_rn(du, w);
                    // networkx: `if vu_dist > cutoff: continue` -> membership is dist <= cutoff
                    if (nd <= cutoff) {
                        const unsigned long long nb = static_cast<unsigned long long>(__double_as_longlong(nd));
                        const unsigned long long old = atomicMin(&dist[v], nb);
                        if (nb < old) {
                            if (old == SP_INF) atomicOr(&s_row[v >> 5], 1u << (v & 31));
                            if (atomicExch(&stamp[v], level_stamp) != level_stamp) {
                                // warp-aggregated frontier append
                                cg::coalesced_group act = cg::coalesced_threads();
                                int base = 0;
                                if (act.thread_rank() == 0) base = atomicAdd(cnt_next, static_cast<int>(act.size()));
                                base = act.shfl(base, 0);
                                qn[base + act.thread_rank()] = v;
                            }
                        }
                    }
                }
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                s_cnt[cur] = 0;
                s_stamp = level_stamp;
            }
            cur ^= 1;
            __syncthreads();
        }
        // store the row once (coalesced) and restore the workspace for the next source
        uint32_t* out = words + src * ld;
        for (int64_t w = threadIdx.x; w < ld; w += blockDim.x) {
            uint32_t bits = s_row[w];
            out[w] = bits;
            s_row[w] = 0;
            while (bits) {
                int b = __ffs(bits) - 1;
                bits &= bits - 1;
                dist[w * 32 + b] = SP_INF;
            }
        }
        __syncthreads();
    }
}

__global__ void k_fill_u64(unsigned long long* p, size_t n, unsigned long long v) {
    size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    size_t step = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (; i < n; i += step) p[i] = v;
}

// ------------------------------------------------------------------------------------------------ row sums / unpack
__global__ void k_rowsums(const uint32_t* __restrict__ words, int64_t n, int64_t ld, int64_t* __restrict__ out) {
    const int64_t row = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    int64_t acc = 0;
    for (int64_t w = lane; w < ld; w += 32) acc += __popc(words[row * ld + w]);
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[row] = acc;
}

template <class T>
__global__ void k_unpack(const uint32_t* __restrict__ words, int64_t n, int64_t ld, int64_t r0, int64_t r1,
                         T* __restrict__ out) {
    const int64_t total = (r1 - r0) * n;
    int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    const int64_t step = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (; i < total; i += step) {
        const int64_t r = r0 + i / n, t = i % n;
        out[i] = static_cast<T>((words[r * ld + (t >> 5)] >> (t & 31)) & 1u);
    }
}

}  // namespace sb

using namespace sb;

// ================================================================================================ C ABI
extern "C" {

int sb_abi_version(void) { return SB_ABI_VERSION; }
const char* sb_last_error(void) { return sb::last_error().c_str(); }

int sb_ctx_create(int device, sb_ctx** out) {
    SB_API_BEGIN
    SB_CHECK(out != nullptr, "sb_ctx_create: out is NULL");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    SB_CHECK(e == cudaSuccess && count > 0,
             "sb_ctx_create: no CUDA device is visible (%s); libsafe_b200 has no CPU fallback",
             cudaGetErrorString(e));
    if (device < 0) SB_CUDA(cudaGetDevice(&device));
    SB_CHECK(device < count, "sb_ctx_create: device %d out of range (%d visible)", device, count);
    cudaDeviceProp prop;
    SB_CUDA(cudaGetDeviceProperties(&prop, device));
    SB_CHECK(prop.major == 10, "sb_ctx_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only",
             device, prop.major, prop.minor);
    SB_CUDA(cudaSetDevice(device));
    {
        // keep freed blocks in the stream-ordered pool instead of returning them to the driver
        cudaMemPool_t pool;
        SB_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
        uint64_t keep = UINT64_MAX;
        SB_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    }
    sb_ctx* c = new sb_ctx;
    c->device = device;
    c->num_sms = prop.multiProcessorCount;
    c->stream = nullptr;
    *out = c;
    SB_API_END
}

int sb_current_device(void) {
    int device = -1;
    if (cudaGetDevice(&device) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    return device;
}

int sb_ctx_destroy(sb_ctx* ctx) {
    SB_API_BEGIN
    if (ctx) {
        ctx->bind();
        cudaStreamSynchronize(ctx->stream);
        delete ctx;
    }
    SB_API_END
}

int sb_ctx_release_memory(sb_ctx* ctx) {
    SB_API_BEGIN
    SB_CHECK(ctx, "sb_ctx_release_memory: ctx is NULL");
    ctx->bind();
    SB_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->ws_bcat.release();
    ctx->ws_flag_ij.release();
    ctx->ws_flag_p.release();
    ctx->ws_cpk.release();
    ctx->ws_counter.release();
    dev_cache_trim(ctx->device);
    SB_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaMemPool_t pool;
    SB_CUDA(cudaDeviceGetDefaultMemPool(&pool, ctx->device));
    SB_CUDA(cudaMemPoolTrimTo(pool, 0));
    SB_API_END
}

int sb_ctx_set_stream(sb_ctx* ctx, void* cuda_stream) {
    SB_API_BEGIN
    SB_CHECK(ctx, "sb_ctx_set_stream: ctx is NULL");
    ctx->stream = static_cast<cudaStream_t>(cuda_stream);
    SB_API_END
}

int sb_ctx_synchronize(sb_ctx* ctx) {
    SB_API_BEGIN
    SB_CHECK(ctx, "sb_ctx_synchronize: ctx is NULL");
    ctx->bind();
    SB_CUDA(cudaStreamSynchronize(ctx->stream));
    SB_API_END
}

int64_t sb_ctx_launch_count(sb_ctx* ctx) { return ctx ? ctx->launches : -1; }

int sb_ctx_profile(sb_ctx* ctx, int enable) {
    SB_API_BEGIN
    SB_CHECK(ctx, "sb_ctx_profile: ctx is NULL");
    ctx->profile = enable != 0;
    SB_API_END
}

int sb_ctx_kernel_ms(sb_ctx* ctx, int kernel_class, double* ms_out, int64_t* count_out) {
    SB_API_BEGIN
    SB_CHECK(ctx && ms_out && count_out, "sb_ctx_kernel_ms: NULL argument");
    SB_CHECK(kernel_class >= 0 && kernel_class < SB_K_CLASSES, "sb_ctx_kernel_ms: unknown kernel class %d",
             kernel_class);
    ctx->bind();
    SB_CUDA(cudaStreamSynchronize(ctx->stream));
    double total = 0.0;
    auto& v = ctx->timers[kernel_class];
    for (auto& pr : v) {
        float ms = 0.f;
        SB_CUDA(cudaEventElapsedTime(&ms, pr.first, pr.second));
        total += ms;
        cudaEventDestroy(pr.first);
        cudaEventDestroy(pr.second);
    }
    *ms_out = total;
    *count_out = static_cast<int64_t>(v.size());
    v.clear();
    SB_API_END
}

int sb_ctx_device(sb_ctx* ctx) { return ctx ? ctx->device : -1; }

int sb_host_register(void* ptr, int64_t bytes) {
    SB_API_BEGIN
    SB_CUDA(cudaHostRegister(ptr, static_cast<size_t>(bytes), cudaHostRegisterDefault));
    SB_API_END
}
int sb_host_unregister(void* ptr) {
    SB_API_BEGIN
    SB_CUDA(cudaHostUnregister(ptr));
    SB_API_END
}

int64_t sb_neigh_ld(int64_t n) { return sb_ld_words(n); }

int sb_neigh_create(sb_ctx* ctx, int64_t n, sb_neigh** out) {
    SB_API_BEGIN
    SB_CHECK(ctx && out, "sb_neigh_create: NULL argument");
    SB_CHECK(n > 0 && n < (1ll << 31) - 64, "sb_neigh_create: n=%lld out of range", (long long)n);
    ctx->bind();
    sb_neigh* a = new sb_neigh;
    a->ctx = ctx;
    a->n = n;
    a->ld = sb_ld_words(n);
    a->owned = true;
    size_t bytes = static_cast<size_t>(n) * a->ld * sizeof(uint32_t);
    try {
        a->words = static_cast<uint32_t*>(dev_alloc(bytes));
    } catch (...) {
        delete a;
        throw;
    }
    SB_CUDA(cudaMemsetAsync(a->words, 0, bytes, ctx->stream));
    *out = a;
    SB_API_END
}

int sb_neigh_wrap_dev(sb_ctx* ctx, int64_t n, uint32_t* words_dev, sb_neigh** out) {
    SB_API_BEGIN
    SB_CHECK(ctx && out && words_dev, "sb_neigh_wrap_dev: NULL argument");
    SB_CHECK(n > 0 && n < (1ll << 31) - 64, "sb_neigh_wrap_dev: n=%lld out of range", (long long)n);
    SB_CHECK((reinterpret_cast<uintptr_t>(words_dev) & 15) == 0, "sb_neigh_wrap_dev: buffer must be 16-byte aligned");
    sb_neigh* a = new sb_neigh;
    a->ctx = ctx;
    a->n = n;
    a->ld = sb_ld_words(n);
    a->words = words_dev;
    a->owned = false;
    *out = a;
    SB_API_END
}

int sb_neigh_destroy(sb_neigh* a) {
    SB_API_BEGIN
    if (a) {
        if (a->owned && a->words) {
            a->ctx->bind();
            dev_free(a->words);
        }
        delete a;
    }
    SB_API_END
}

int64_t sb_neigh_n(const sb_neigh* a) { return a ? a->n : -1; }
uint32_t* sb_neigh_words_dev(sb_neigh* a) { return a ? a->words : nullptr; }

static void check_rows(const sb_neigh* a, int64_t r0, int64_t r1, const char* who) {
    SB_CHECK(a, "%s: neighborhood handle is NULL", who);
    SB_CHECK(0 <= r0 && r0 <= r1 && r1 <= a->n, "%s: row range [%lld,%lld) outside [0,%lld)", who, (long long)r0,
             (long long)r1, (long long)a->n);
}

int sb_neigh_euclid(sb_neigh* a, const double* x_host, const double* y_host, double nr, int64_t row0, int64_t row1) {
    SB_API_BEGIN
    check_rows(a, row0, row1, "sb_neigh_euclid");
    SB_CHECK(x_host && y_host, "sb_neigh_euclid: coordinates are NULL");
    SB_CHECK(nr == nr, "sb_neigh_euclid: radius is NaN");
    sb_ctx* ctx = a->ctx;
    ctx->bind();
    if (row0 == row1) return 0;
    const int64_t n = a->n;
    // largest double s with sqrt_rn(s) < nr  (host sqrt is correctly rounded)
    double s_star = 0.0;
    int any_pair = 0;
    if (nr > 0.0) {
        any_pair = 1;
        if (std::isinf(nr)) {
            s_star = std::numeric_limits<double>::max();
        } else {
            s_star = nr * nr;
            if (std::isinf(s_star)) s_star = std::numeric_limits<double>::max();
            while (std::sqrt(s_star) >= nr) s_star = std::nextafter(s_star, -1.0);
            while (true) {
                double up = std::nextafter(s_star, std::numeric_limits<double>::infinity());
                if (std::isinf(up) || !(std::sqrt(up) < nr)) break;
                s_star = up;
            }
        }
    }
    DevBuf<double> xy;
    xy.reserve(2 * n);
    SB_CUDA(cudaMemcpyAsync(xy.p, x_host, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    SB_CUDA(cudaMemcpyAsync(xy.p + n, y_host, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    dim3 grid(static_cast<unsigned>(sb_ceil_div(a->ld, EU_WARPS * 8)),
              static_cast<unsigned>(sb_ceil_div(row1 - row0, EU_ROWS)));
    SB_CHECK(grid.y <= 65535, "sb_neigh_euclid: row range too large for one launch (%lld rows)",
             (long long)(row1 - row0));
    {
        KernelTimer kt(ctx, SB_K_EUCLID);
        k_euclid<<<grid, EU_WARPS * 32, 0, ctx->stream>>>(xy.p, xy.p + n, n, s_star, any_pair, row0, row1, a->words,
                                                          a->ld);
        SB_LAUNCH_CHECK(ctx);
    }
    SB_CUDA(cudaStreamSynchronize(ctx->stream));
    SB_API_END
}

int sb_neigh_shortpath(sb_neigh* a, const int64_t* indptr_host, const int32_t* indices_host,
                       const double* length_host, double cutoff, int64_t row0, int64_t row1) {
    SB_API_BEGIN
    check_rows(a, row0, row1, "sb_neigh_shortpath");
    SB_CHECK(indptr_host && indices_host, "sb_neigh_shortpath: CSR arrays are NULL");
    SB_CHECK(cutoff == cutoff, "sb_neigh_shortpath: cutoff is NaN");
    sb_ctx* ctx = a->ctx;
    ctx->bind();
    if (row0 == row1) return 0;
    const int64_t n = a->n;
    const int64_t nnz = indptr_host[n];
    SB_CHECK(indptr_host[0] == 0 && nnz >= 0, "sb_neigh_shortpath: malformed indptr");
    for (int64_t i = 0; i < n; ++i)
        SB_CHECK(indptr_host[i] <= indptr_host[i + 1], "sb_neigh_shortpath: indptr not monotone at row %lld",
                 (long long)i);
    for (int64_t e = 0; e < nnz; ++e)
        SB_CHECK(indices_host[e] >= 0 && indices_host[e] < n, "sb_neigh_shortpath: column index %d out of range",
                 indices_host[e]);
    if (length_host)
        for (int64_t e = 0; e < nnz; ++e)
            SB_CHECK(length_host[e] >= 0.0, "sb_neigh_shortpath: edge length %g at entry %lld is negative or NaN",
                     length_host[e], (long long)e);

    const size_t smem = static_cast<size_t>(a->ld) * sizeof(uint32_t);
    SB_CHECK(smem <= 200 * 1024, "sb_neigh_shortpath: n=%lld exceeds the shared-memory row bitmap (max ~1.6M nodes)",
             (long long)n);
    static const int group = getenv("SB_SSSP_GROUP") ? atoi(getenv("SB_SSSP_GROUP")) : SP_GROUP_DEFAULT;
    static const int max_per_sm = getenv("SB_SSSP_PER_SM") ? atoi(getenv("SB_SSSP_PER_SM")) : 8;
    auto kern = group == 4 ? k_sssp<4> : (group == 8 ? k_sssp<8> : (group == 32 ? k_sssp<32> : k_sssp<16>));
    SB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    int per_sm = 0;
    SB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, SP_THREADS, smem));
    SB_CHECK(per_sm >= 1, "sb_neigh_shortpath: kernel does not fit on an SM");
    per_sm = std::min(per_sm, max_per_sm);
    int64_t grid = std::min<int64_t>(static_cast<int64_t>(per_sm) * ctx->num_sms, row1 - row0);

    DevBuf<int64_t> d_indptr;
    DevBuf<int32_t> d_indices;
    DevBuf<double> d_len;
    DevBuf<unsigned long long> d_dist;
    DevBuf<uint32_t> d_stamp;
    DevBuf<int32_t> d_queue;
    DevBuf<unsigned int> d_next;
    d_indptr.reserve(n + 1);
    d_indices.reserve(std::max<int64_t>(nnz, 1));
    if (length_host) d_len.reserve(std::max<int64_t>(nnz, 1));
    d_dist.reserve(static_cast<size_t>(grid) * n);
    d_stamp.reserve(static_cast<size_t>(grid) * n);
    d_queue.reserve(static_cast<size_t>(grid) * 2 * n);
    d_next.reserve(1);
    cudaStream_t st = ctx->stream;
    SB_CUDA(cudaMemcpyAsync(d_indptr.p, indptr_host, (n + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    SB_CUDA(cudaMemcpyAsync(d_indices.p, indices_host, nnz * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    if (length_host)
        SB_CUDA(cudaMemcpyAsync(d_len.p, length_host, nnz * sizeof(double), cudaMemcpyHostToDevice, st));
    k_fill_u64<<<ctx->num_sms * 4, 256, 0, st>>>(d_dist.p, static_cast<size_t>(grid) * n, SP_INF);
    SB_LAUNCH_CHECK(ctx);
    SB_CUDA(cudaMemsetAsync(d_stamp.p, 0, static_cast<size_t>(grid) * n * sizeof(uint32_t), st));
    SB_CUDA(cudaMemsetAsync(d_next.p, 0, sizeof(unsigned int), st));
    SsspWs ws{d_dist.p, d_stamp.p, d_queue.p, d_next.p};
    {
        KernelTimer kt(ctx, SB_K_SSSP);
        kern<<<static_cast<unsigned>(grid), SP_THREADS, smem, st>>>(d_indptr.p, d_indices.p,
                                                                      length_host ? d_len.p : nullptr, n, cutoff,
                                                                      row0, row1, ws, a->words, a->ld);
        SB_LAUNCH_CHECK(ctx);
    }
    SB_CUDA(cudaStreamSynchronize(st));
    SB_API_END
}

int sb_neigh_upload_packed(sb_neigh* a, const uint32_t* words_host, int64_t row0, int64_t row1) {
    SB_API_BEGIN
    check_rows(a, row0, row1, "sb_neigh_upload_packed");
    SB_CHECK(words_host, "sb_neigh_upload_packed: words_host is NULL");
    a->ctx->bind();
    copy_in(a->ctx, a->words + row0 * a->ld, words_host, (row1 - row0) * a->ld * sizeof(uint32_t));
    SB_API_END
}

int sb_neigh_download_packed(sb_neigh* a, uint32_t* words_host, int64_t row0, int64_t row1) {
    SB_API_BEGIN
    check_rows(a, row0, row1, "sb_neigh_download_packed");
    SB_CHECK(words_host, "sb_neigh_download_packed: words_host is NULL");
    a->ctx->bind();
    copy_out(a->ctx, words_host, a->words + row0 * a->ld, (row1 - row0) * a->ld * sizeof(uint32_t));
    SB_API_END
}

int sb_neigh_rowsums(sb_neigh* a, int64_t* out_host) {
    SB_API_BEGIN
    SB_CHECK(a && out_host, "sb_neigh_rowsums: NULL argument");
    sb_ctx* ctx = a->ctx;
    ctx->bind();
    DevBuf<int64_t> d;
    d.reserve(a->n);
    const int threads = 256;
    k_rowsums<<<static_cast<unsigned>(sb_ceil_div(a->n * 32, threads)), threads, 0, ctx->stream>>>(a->words, a->n,
                                                                                                    a->ld, d.p);
    SB_LAUNCH_CHECK(ctx);
    SB_CUDA(cudaMemcpyAsync(out_host, d.p, a->n * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    SB_CUDA(cudaStreamSynchronize(ctx->stream));
    SB_API_END
}

int sb_neigh_unpack_rows(sb_neigh* a, int64_t r0, int64_t r1, int elem_bytes, void* out_host) {
    SB_API_BEGIN
    check_rows(a, r0, r1, "sb_neigh_unpack_rows");
    SB_CHECK(out_host, "sb_neigh_unpack_rows: out_host is NULL");
    SB_CHECK(elem_bytes == 1 || elem_bytes == 8, "sb_neigh_unpack_rows: elem_bytes must be 1 or 8");
    sb_ctx* ctx = a->ctx;
    ctx->bind();
    if (r0 == r1) return 0;
    // stream the requested rows through a bounded staging buffer
    const int64_t max_elems = 256ll << 20;  // 256 Mi elements per chunk
    const int64_t rows_per_chunk = std::max<int64_t>(1, max_elems / a->n);
    DevBuf<uint8_t> stage;
    stage.reserve(static_cast<size_t>(std::min(rows_per_chunk, r1 - r0)) * a->n * elem_bytes);
    for (int64_t r = r0; r < r1; r += rows_per_chunk) {
        const int64_t re = std::min(r1, r + rows_per_chunk);
        const int64_t total = (re - r) * a->n;
        const unsigned blocks = static_cast<unsigned>(std::min<int64_t>(sb_ceil_div(total, 256), ctx->num_sms * 32));
        if (elem_bytes == 1)
            k_unpack<uint8_t><<<blocks, 256, 0, ctx->stream>>>(a->words, a->n, a->ld, r, re, stage.p);
        else
            k_unpack<int64_t><<<blocks, 256, 0, ctx->stream>>>(a->words, a->n, a->ld, r, re,
                                                               reinterpret_cast<int64_t*>(stage.p));
        SB_LAUNCH_CHECK(ctx);
        SB_CUDA(cudaMemcpyAsync(static_cast<uint8_t*>(out_host) + (r - r0) * a->n * elem_bytes, stage.p,
                                total * elem_bytes, cudaMemcpyDeviceToHost, ctx->stream));
        SB_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    SB_API_END
}

}  // extern "C"
