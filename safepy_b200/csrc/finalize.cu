// What follows the permutation null / the hypergeometric test inside SAFE.compute_pvalues, and the next rows of
// SURVEY.md section 8f that consume its output:
//   streaming null      sb_enrich_null_begin / _add / _counts: the count arrays stay on the device while the caller is
//                       still replaying the reference's sequential RNG stream
//   k_null_tail         counts -> p-values -> NES -> nes_binary -> enriched neighborhoods per attribute, i.e. the tail
//                       of compute_pvalues_by_randomization (reference safepy/safe.py:526-554) and of compute_pvalues
//                       (safe.py:466-472) in one pass; p-values and NES of the P + 1 possible counts come from
//                       host-made tables, so they carry the host libm's bits
//   k_bh_rows           Benjamini-Hochberg adjustment of every row across attributes (multiple_testing=True,
//                       safe.py:536-542 / 599-605: statsmodels fdrcorrection(method='indep') per row) after a
//                       segmented sort of the row
//   k_jaccard           pairwise Jaccard distances between nes_binary columns of the top attributes, the metric
//                       evaluation inside define_domains' linkage(..., metric='jaccard') (safe.py:672-675)
//   copy_out            large device -> host result copies through a pinned ring drained by worker threads
#include <sys/mman.h>

#include <algorithm>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

#include <cub/device/device_segmented_sort.cuh>

#include "enrich.cuh"

namespace sb {

// ================================================================================================ copy_out / copy_in
namespace {

// Pinned ring + worker threads shared by the two directions.  Large transfers between the device and PAGEABLE host
// memory are otherwise bound by one thread: the driver stages them through its own pinned buffer with a
// single-threaded memcpy, and for a freshly allocated destination that thread also takes every first-touch page fault.
struct HostRing {
    static constexpr int kSlots = 32;
    static constexpr size_t kSlot = 4u << 20;
    struct Task {
        int slot;
        int device;      // >= 0: device -> host piece (wait for the slot's event, then copy out of the ring)
        void* host;      //  < 0: host -> device piece (copy into the ring, then flag the slot ready)
        size_t bytes;
    };
    char* pinned = nullptr;
    cudaEvent_t ev[kSlots];
    bool busy[kSlots];   // a worker still owns the slot
    bool used[kSlots];   // the slot's event has been recorded by an upload of this call
    std::mutex mu;       // queue + flags
    std::condition_variable cv_work, cv_done;
    std::deque<Task> queue;
    int pending = 0;
    std::mutex call_mu;  // one transfer at a time
    int workers = 0;

    void start() {
        SB_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&pinned), kSlots * kSlot, cudaHostAllocPortable));
        for (int i = 0; i < kSlots; ++i) {
            SB_CUDA(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
            busy[i] = false;
            used[i] = false;
        }
        const unsigned hw = std::thread::hardware_concurrency();
        workers = static_cast<int>(std::max(2u, std::min(static_cast<unsigned>(kSlots), hw ? hw : 2u)));
        for (int w = 0; w < workers; ++w) std::thread([this] { run(); }).detach();
    }
    void run() {
        for (;;) {
            Task t;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_work.wait(lk, [this] { return !queue.empty(); });
                t = queue.front();
                queue.pop_front();
            }
            char* ring = pinned + static_cast<size_t>(t.slot) * kSlot;
            if (t.device >= 0) {
                cudaSetDevice(t.device);
                cudaEventSynchronize(ev[t.slot]);
                memcpy(t.host, ring, t.bytes);
            } else {
                memcpy(ring, t.host, t.bytes);
            }
            {
                std::lock_guard<std::mutex> lk(mu);
                busy[t.slot] = false;
                --pending;
            }
            cv_done.notify_all();
        }
    }
    void push(const Task& t) {
        {
            std::lock_guard<std::mutex> lk(mu);
            busy[t.slot] = true;
            ++pending;
            queue.push_back(t);
        }
        cv_work.notify_one();
    }
    void wait_slot(int slot) {
        std::unique_lock<std::mutex> lk(mu);
        cv_done.wait(lk, [&] { return !busy[slot]; });
    }
    void wait_all() {
        std::unique_lock<std::mutex> lk(mu);
        cv_done.wait(lk, [&] { return pending == 0; });
    }
};

// one ring per device (its events belong to that device's context); created with the device current
HostRing* host_ring(int device) {
    static HostRing* rings[64] = {};  // leaked on purpose: detached workers may outlive static destruction
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    SB_CHECK(device >= 0 && device < 64, "device ordinal %d out of range", device);
    if (!rings[device]) {
        HostRing* r = new HostRing;
        r->start();
        rings[device] = r;
    }
    return rings[device];
}

bool is_pageable(const void* p) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return attr.type == cudaMemoryTypeUnregistered;
}

}  // namespace

void copy_out(sb_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes) {
    if (bytes == 0) return;
    cudaStream_t st = ctx->stream;
    if (bytes < 2 * HostRing::kSlot || !is_pageable(dst_host)) {
        SB_CUDA(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, st));
        SB_CUDA(cudaStreamSynchronize(st));
        return;
    }
    PhaseTrace tr(ctx, "copy_out");
    {
        // a fresh NumPy buffer is untouched anonymous memory: ask for huge pages on its 2 MB-aligned interior so the
        // first-touch faults below come 512x fewer (no-op where transparent huge pages are off)
        const uintptr_t lo = (reinterpret_cast<uintptr_t>(dst_host) + (1u << 21) - 1) & ~((uintptr_t(1) << 21) - 1);
        const uintptr_t hi = (reinterpret_cast<uintptr_t>(dst_host) + bytes) & ~((uintptr_t(1) << 21) - 1);
        if (hi > lo) madvise(reinterpret_cast<void*>(lo), hi - lo, MADV_HUGEPAGE);
    }
    HostRing* r = host_ring(ctx->device);
    std::lock_guard<std::mutex> call(r->call_mu);
    const char* src = static_cast<const char*>(src_dev);
    char* dst = static_cast<char*>(dst_host);
    int slot = 0;
    cudaError_t err = cudaSuccess;
    for (size_t off = 0; off < bytes && err == cudaSuccess; off += HostRing::kSlot, slot = (slot + 1) % HostRing::kSlots) {
        const size_t len = std::min(HostRing::kSlot, bytes - off);
        r->wait_slot(slot);
        err = cudaMemcpyAsync(r->pinned + static_cast<size_t>(slot) * HostRing::kSlot, src + off, len,
                              cudaMemcpyDeviceToHost, st);
        if (err == cudaSuccess) err = cudaEventRecord(r->ev[slot], st);
        if (err == cudaSuccess) r->push({slot, ctx->device, dst + off, len});
    }
    r->wait_all();
    SB_CUDA(err);
    SB_CUDA(cudaStreamSynchronize(st));
}

void copy_in(sb_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes) {
    if (bytes == 0) return;
    cudaStream_t st = ctx->stream;
    if (bytes < 2 * HostRing::kSlot || !is_pageable(src_host)) {
        SB_CUDA(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, st));
        SB_CUDA(cudaStreamSynchronize(st));
        return;
    }
    PhaseTrace tr(ctx, "copy_in");
    HostRing* r = host_ring(ctx->device);
    std::lock_guard<std::mutex> call(r->call_mu);
    char* src = static_cast<char*>(const_cast<void*>(src_host));
    char* dst = static_cast<char*>(dst_dev);
    const size_t pieces = (bytes + HostRing::kSlot - 1) / HostRing::kSlot;
    for (int i = 0; i < HostRing::kSlots; ++i) r->used[i] = false;
    cudaError_t err = cudaSuccess;
    size_t enq = 0;
    for (size_t iss = 0; iss < pieces && err == cudaSuccess; ++iss) {
        // keep the workers a ring ahead of the uploads
        for (; enq < pieces && enq - iss < static_cast<size_t>(HostRing::kSlots); ++enq) {
            const int slot = static_cast<int>(enq % HostRing::kSlots);
            if (r->used[slot]) cudaEventSynchronize(r->ev[slot]);  // the slot's previous upload has left the ring
            const size_t off = enq * HostRing::kSlot;
            r->push({slot, -1, src + off, std::min(HostRing::kSlot, bytes - off)});
        }
        const int slot = static_cast<int>(iss % HostRing::kSlots);
        const size_t off = iss * HostRing::kSlot;
        r->wait_slot(slot);
        err = cudaMemcpyAsync(dst + off, r->pinned + static_cast<size_t>(slot) * HostRing::kSlot,
                              std::min(HostRing::kSlot, bytes - off), cudaMemcpyHostToDevice, st);
        if (err == cudaSuccess) err = cudaEventRecord(r->ev[slot], st);
        r->used[slot] = true;
    }
    r->wait_all();
    SB_CUDA(err);
    SB_CUDA(cudaStreamSynchronize(st));
}

// ================================================================================================ kernels
__device__ __forceinline__ double qnan() { return __longlong_as_double(0x7FF8000000000000ll); }

constexpr int kTailRows = 16;  // rows per block of the column-strip kernels

// One thread per attribute column, kTailRows consecutive rows per block: loads and stores of a warp are 256 contiguous
// bytes of one row, and the enriched-neighborhood count of a column costs one atomic per thread.
//   FROM_COUNTS   pn / pp / NES from the count tables (bit-identical to the host's division and log10)
//   !FROM_COUNTS  pn / pp already hold (FDR-adjusted) p-values; NES = -log10(p == 0 ? floor : p) on the device
template <bool FROM_COUNTS>
__global__ void __launch_bounds__(256) k_null_tail(const uint32_t* __restrict__ cneg, const uint32_t* __restrict__ cpos,
                                                   const double* __restrict__ ns, const double* __restrict__ ptab,
                                                   const double* __restrict__ nestab, uint32_t tab_len, double floor_p,
                                                   int sign, double thr, int64_t rows, int64_t m,
                                                   double* __restrict__ pn, double* __restrict__ pp,
                                                   double* __restrict__ nes, double* __restrict__ nb,
                                                   int32_t* __restrict__ colcnt) {
    const int64_t j = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
    if (j >= m) return;
    const int64_t r0 = static_cast<int64_t>(blockIdx.y) * kTailRows;
    const int64_t r1 = min(rows, r0 + kTailRows);
    int32_t enriched = 0;
    for (int64_t r = r0; r < r1; ++r) {
        const int64_t at = r * m + j;
        double nes_pos, nes_neg;
        if (FROM_COUNTS) {
            const uint32_t cn = min(cneg[at], tab_len - 1), cp = min(cpos[at], tab_len - 1);
            const bool dead = isnan(ns[at]);  // safe.py:528-530
            const double vn = dead ? qnan() : ptab[cn], vp = dead ? qnan() : ptab[cp];
            pn[at] = vn;
            pp[at] = vp;
            nes_neg = dead ? qnan() : nestab[cn];
            nes_pos = dead ? qnan() : nestab[cp];
        } else {
            const double vn = pn[at], vp = pp[at];
            nes_neg = -log10(vn == 0.0 ? floor_p : vn);  // safe.py:546-547
            nes_pos = -log10(vp == 0.0 ? floor_p : vp);
        }
        const double v = sign == 0 ? nes_pos : (sign == 1 ? nes_neg : __dsub_rn(nes_pos, nes_neg));
        nes[at] = v;
        const bool hit = fabs(v) > thr;  // NaN compares false: nes_binary stays 0 (safe.py:466-468)
        nb[at] = hit ? 1.0 : 0.0;
        enriched += hit;
    }
    if (enriched) atomicAdd(&colcnt[j], enriched);
}

// hypergeometric tail: optional nes = -log10(p) (after FDR), then nes_binary and the column counts
template <bool NES_FROM_P>
__global__ void __launch_bounds__(256) k_binarize(const double* __restrict__ p, double* __restrict__ nes, double thr,
                                                  int64_t rows, int64_t m, double* __restrict__ nb,
                                                  int32_t* __restrict__ colcnt) {
    const int64_t j = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
    if (j >= m) return;
    const int64_t r0 = static_cast<int64_t>(blockIdx.y) * kTailRows;
    const int64_t r1 = min(rows, r0 + kTailRows);
    int32_t enriched = 0;
    for (int64_t r = r0; r < r1; ++r) {
        const int64_t at = r * m + j;
        double v;
        if (NES_FROM_P) {
            v = -log10(p[at]);
            nes[at] = v;
        } else {
            v = nes[at];
        }
        const bool hit = fabs(v) > thr;
        nb[at] = hit ? 1.0 : 0.0;
        enriched += hit;
    }
    if (enriched) atomicAdd(&colcnt[j], enriched);
}

// compute_pvalues' look at the attribute matrix (safe.py:453-458): NaNs per column and the number of values that
// are neither 0, 1 nor NaN.  Same column-strip layout as the tail kernels.
template <class T>
__global__ void __launch_bounds__(256) k_attr_summary(const T* __restrict__ b, int64_t n, int64_t m,
                                                      unsigned long long* __restrict__ nan_per_col,
                                                      unsigned long long* __restrict__ other) {
    const int64_t j = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
    const int64_t r0 = static_cast<int64_t>(blockIdx.y) * 64;
    const int64_t r1 = min(n, r0 + 64);
    unsigned int nans = 0, oth = 0;
    if (j < m) {
        for (int64_t r = r0; r < r1; ++r) {
            const T v = b[r * m + j];
            const bool is_nan = v != v;
            nans += is_nan;
            oth += !is_nan && v != T(0) && v != T(1);
        }
        if (nans) atomicAdd(&nan_per_col[j], static_cast<unsigned long long>(nans));
    }
    oth = __reduce_add_sync(0xffffffffu, oth);
    if ((threadIdx.x & 31) == 0 && oth) atomicAdd(other, static_cast<unsigned long long>(oth));
}

__global__ void k_colcnt_to_f64(const int32_t* __restrict__ c, int64_t m, double* __restrict__ out) {
    const int64_t j = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (j < m) out[j] = static_cast<double>(c[j]);
}

__global__ void k_bh_prepare(int64_t rows, int64_t m, int32_t* __restrict__ idx, int64_t* __restrict__ offsets) {
    const int64_t cells = rows * m;
    int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int64_t step = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t r = i; r <= rows; r += step) offsets[r] = r * m;
    for (; i < cells; i += step) idx[i] = static_cast<int32_t>(i % m);
}

// One CTA per row.  keys = the row's p-values ascending, idx = their columns, out = the row itself (overwritten).
// statsmodels' fdrcorrection:
//   ecdf[k] = (k + 1) / m;  raw[k] = p_sorted[k] / ecdf[k];  adj = reverse running minimum of raw, capped at 1,
// scattered back to the original columns.  A NaN anywhere in the row propagates through np.minimum.accumulate from
// the end (argsort puts NaNs last), so the whole row becomes NaN.  Tied p-values all receive the value of the last
// of them (the quotient is monotone in k), so the order a sort leaves ties in does not matter.
__global__ void __launch_bounds__(256) k_bh_rows(const double* __restrict__ keys, const int32_t* __restrict__ idx,
                                                 int64_t m, double* __restrict__ out) {
    const int64_t row = blockIdx.x;
    const double* k = keys + row * m;
    const int32_t* ix = idx + row * m;
    double* o = out + row * m;
    const int t = threadIdx.x;
    const double dm = static_cast<double>(m);
    const int64_t seg = (m + 255) / 256;
    const int64_t b = min(m, t * seg), e = min(m, b + seg);
    __shared__ double part[256];
    double mn = __longlong_as_double(0x7FF0000000000000ll);  // +inf
    // NaNs are looked for in the unsorted row (`out` still holds it: the adjustment is in place): the segmented
    // sort's comparison path for short rows can push a NaN behind its padding keys and lose it
    int nan_here = 0;
    for (int64_t i = t; i < m; i += 256) nan_here |= isnan(o[i]);
    for (int64_t i = b; i < e; ++i) mn = fmin(mn, __ddiv_rn(k[i], __ddiv_rn(static_cast<double>(i + 1), dm)));
    part[t] = mn;
    const bool has_nan = __syncthreads_or(nan_here);
    if (has_nan) {  // not through idx: with a NaN among the keys the sort may not even return a permutation
        for (int64_t i = t; i < m; i += 256) o[i] = qnan();
        return;
    }
    for (int off = 1; off < 256; off <<= 1) {  // inclusive suffix minimum
        const double v = t + off < 256 ? part[t + off] : __longlong_as_double(0x7FF0000000000000ll);
        __syncthreads();
        part[t] = fmin(part[t], v);
        __syncthreads();
    }
    double run = t + 1 < 256 ? part[t + 1] : __longlong_as_double(0x7FF0000000000000ll);
    for (int64_t i = e - 1; i >= b; --i) {
        run = fmin(run, __ddiv_rn(k[i], __ddiv_rn(static_cast<double>(i + 1), dm)));
        o[ix[i]] = fmin(run, 1.0);
    }
}

// ---------------------------------------------------------------------------------------------- Jaccard
// bits[k][w]: bit b of word w = member[(32 w + b) * m + cols[k]] != 0
__global__ void k_pack_columns(const uint8_t* __restrict__ member, int64_t n, int64_t m,
                               const int32_t* __restrict__ cols, int64_t ncols, int64_t words,
                               uint32_t* __restrict__ bits) {
    const int64_t k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int64_t w = blockIdx.y;
    if (k >= ncols) return;
    const int64_t c = cols[k];
    uint32_t v = 0;
    const int64_t t0 = w * 32;
#pragma unroll 8
    for (int b = 0; b < 32; ++b) {
        const int64_t t = t0 + b;
        if (t < n && member[t * m + c]) v |= 1u << b;
    }
    bits[k * words + w] = v;
}

// One warp per pair (i < j): scipy's boolean Jaccard, d = |a xor b| / |a or b| (0 when both are empty), written at
// the pair's position in pdist's condensed order.
__global__ void __launch_bounds__(256) k_jaccard(const uint32_t* __restrict__ bits, int64_t ncols, int64_t words,
                                                 double* __restrict__ out) {
    const int64_t i = blockIdx.x;
    const int64_t j = static_cast<int64_t>(blockIdx.y) * 8 + (threadIdx.x >> 5);
    if (j <= i || j >= ncols) return;
    const int lane = threadIdx.x & 31;
    const uint32_t* a = bits + i * words;
    const uint32_t* b = bits + j * words;
    unsigned int nx = 0, no = 0;
    for (int64_t w = lane; w < words; w += 32) {
        const uint32_t x = a[w], y = b[w];
        nx += __popc(x ^ y);
        no += __popc(x | y);
    }
    for (int o = 16; o; o >>= 1) {
        nx += __shfl_xor_sync(0xffffffffu, nx, o);
        no += __shfl_xor_sync(0xffffffffu, no, o);
    }
    if (lane == 0) {
        const int64_t at = i * ncols - i * (i + 1) / 2 + (j - i - 1);
        out[at] = no ? __ddiv_rn(static_cast<double>(nx), static_cast<double>(no)) : 0.0;
    }
}

// ================================================================================================ host side
namespace {

// Benjamini-Hochberg over the rows of p [rows][m], in place; scratch is sized by the caller for rows * m cells
struct BhScratch {
    DevBuf<double> keys;
    DevBuf<int32_t> idx_in, idx_out;
    DevBuf<int64_t> offsets;
    DevBuf<char> temp;
};

void bh_rows(sb_ctx* ctx, BhScratch& s, double* p, int64_t rows, int64_t m) {
    const int64_t cells = rows * m;
    SB_CHECK(cells < (1ll << 31), "internal error: FDR chunk too large");
    cudaStream_t st = ctx->stream;
    KernelTimer kt(ctx, SB_K_FDR);
    s.keys.reserve(cells);
    s.idx_in.reserve(cells);
    s.idx_out.reserve(cells);
    s.offsets.reserve(rows + 1);
    k_bh_prepare<<<static_cast<unsigned>(std::min<int64_t>(sb_ceil_div(cells, 256), ctx->num_sms * 8)), 256, 0, st>>>(
        rows, m, s.idx_in.p, s.offsets.p);
    SB_LAUNCH_CHECK(ctx);
    size_t temp_bytes = 0;
    SB_CUDA(cub::DeviceSegmentedSort::SortPairs(nullptr, temp_bytes, p, s.keys.p, s.idx_in.p, s.idx_out.p,
                                                static_cast<int>(cells), static_cast<int>(rows), s.offsets.p,
                                                s.offsets.p + 1, st));
    s.temp.reserve(std::max<size_t>(temp_bytes, 1));
    SB_CUDA(cub::DeviceSegmentedSort::SortPairs(s.temp.p, temp_bytes, p, s.keys.p, s.idx_in.p, s.idx_out.p,
                                                static_cast<int>(cells), static_cast<int>(rows), s.offsets.p,
                                                s.offsets.p + 1, st));
    ctx->launches += 3;  // the segmented sort's partition + large/small segment kernels
    k_bh_rows<<<static_cast<unsigned>(rows), 256, 0, st>>>(s.keys.p, s.idx_out.p, m, p);
    SB_LAUNCH_CHECK(ctx);
}

// rows per pass: <= 32 Mi cells of staging per output array, and a grid.y the launch accepts
int64_t chunk_rows(int64_t n, int64_t m) {
    return std::max<int64_t>(1, std::min<int64_t>({n, (32ll << 20) / m, 65535ll * kTailRows}));
}

dim3 tail_grid(int64_t rows, int64_t m) {
    return dim3(static_cast<unsigned>(sb_ceil_div(m, 256)), static_cast<unsigned>(sb_ceil_div(rows, kTailRows)));
}

void colcnt_out(sb_ctx* ctx, const int32_t* colcnt, int64_t m, double* num_enriched_host) {
    if (!num_enriched_host) return;
    DevBuf<double> f;
    f.reserve(m);
    k_colcnt_to_f64<<<static_cast<unsigned>(sb_ceil_div(m, 256)), 256, 0, ctx->stream>>>(colcnt, m, f.p);
    SB_LAUNCH_CHECK(ctx);
    SB_CUDA(cudaMemcpyAsync(num_enriched_host, f.p, m * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    SB_CUDA(cudaStreamSynchronize(ctx->stream));
}

}  // namespace
}  // namespace sb

namespace sb {
// Permutations of an open streaming null go into ONE packed word per cell (pos << 16 | neg) -- no memset / unpack
// per piece, and the array that crosses NVLink when the permutations are sharded over several GPUs -- and are folded
// into the two count arrays before the 16-bit fields could overflow and whenever the counts are read.
void null_flush(sb_enrich* e) {
    if (e->null_pk_perms == 0) return;
    const size_t cells = static_cast<size_t>(e->n) * e->m;
    unpack_add_counts(e->ctx, e->null_pk.p, static_cast<int64_t>(cells), e->null_cnt.p, e->null_cnt.p + cells);
    e->null_pk_perms = 0;
}

void null_count_dev(sb_enrich* e, const int32_t* perm_dev, int64_t num_perm) {
    for (int64_t p0 = 0; p0 < num_perm;) {
        if (e->null_pk_perms >= 60000) null_flush(e);
        const int64_t np = std::min<int64_t>(num_perm - p0, 60000 - e->null_pk_perms);
        int rc = sb_enrich_perm_counts_packed_dev(e, e->null_score, e->null_engine, perm_dev + p0 * e->n, np,
                                                  e->null_pk.p);
        if (rc) fail("%s", sb_last_error());
        e->null_pk_perms += np;
        p0 += np;
    }
}
}  // namespace sb

using namespace sb;

extern "C" {

// ------------------------------------------------------------------------------------------------ streaming null
int sb_enrich_null_begin(sb_enrich* e, int score_type, int engine) {
    SB_API_BEGIN
    SB_CHECK(e, "sb_enrich_null_begin: NULL handle");
    SB_CHECK(score_type == SB_SCORE_SUM || score_type == SB_SCORE_ZSCORE, "unknown neighborhood_score_type %d",
             score_type);
    SB_CHECK(engine == SB_ENGINE_AUTO || engine == SB_ENGINE_SIMT || engine == SB_ENGINE_TC, "unknown engine %d",
             engine);
    sb_ctx* ctx = e->ctx;
    ctx->bind();
    const size_t cells = static_cast<size_t>(e->n) * e->m;
    e->null_cnt.reserve(2 * cells);
    SB_CUDA(cudaMemsetAsync(e->null_cnt.p, 0, 2 * cells * sizeof(uint32_t), ctx->stream));
    e->null_pk.reserve(cells);
    SB_CUDA(cudaMemsetAsync(e->null_pk.p, 0, cells * sizeof(uint32_t), ctx->stream));
    e->null_pk_perms = 0;
    e->null_score = score_type;
    e->null_engine = engine;
    e->null_perms = 0;
    for (int i = 0; i < 7; ++i) e->null_stats[i] = 0;
    SB_API_END
}

int sb_enrich_null_add(sb_enrich* e, const int32_t* perm_rows_host, int64_t num_perm) {
    SB_API_BEGIN
    SB_CHECK(e && perm_rows_host, "sb_enrich_null_add: NULL argument");
    SB_CHECK(e->null_score >= 0, "sb_enrich_null_add: call sb_enrich_null_begin first");
    SB_CHECK(num_perm >= 0, "sb_enrich_null_add: num_perm < 0");
    sb_ctx* ctx = e->ctx;
    ctx->bind();
    const int64_t piece = std::max<int64_t>(1, (256ll << 20) / e->n);  // <= 1 GiB of indices on the device at a time
    for (int64_t p0 = 0; p0 < num_perm; p0 += piece) {
        const int64_t np = std::min(piece, num_perm - p0);
        e->null_perm.reserve(static_cast<size_t>(np) * e->n);
        SB_CUDA(cudaMemcpyAsync(e->null_perm.p, perm_rows_host + p0 * e->n, static_cast<size_t>(np) * e->n * sizeof(int32_t),
                                cudaMemcpyHostToDevice, ctx->stream));
        null_count_dev(e, e->null_perm.p, np);
        SB_CUDA(cudaStreamSynchronize(ctx->stream));  // the staging buffer is reused by the next piece
        for (int i = 0; i < 7; ++i) {
            if (i == 2 || i == 3 || i == 4)
                e->null_stats[i] = e->stats[i];
            else
                e->null_stats[i] += e->stats[i];
        }
        e->null_perms += np;
    }
    for (int i = 0; i < 7; ++i) e->stats[i] = e->null_stats[i];
    SB_API_END
}

int sb_enrich_null_counts(sb_enrich* e, int64_t* num_perm_out, uint32_t* counts_neg_host, uint32_t* counts_pos_host) {
    SB_API_BEGIN
    SB_CHECK(e, "sb_enrich_null_counts: NULL handle");
    SB_CHECK(e->null_score >= 0, "sb_enrich_null_counts: no null has been started on this plan");
    sb_ctx* ctx = e->ctx;
    ctx->bind();
    const size_t cells = static_cast<size_t>(e->n) * e->m;
    null_flush(e);
    if (num_perm_out) *num_perm_out = e->null_perms;
    if (counts_neg_host) copy_out(ctx, counts_neg_host, e->null_cnt.p, cells * sizeof(uint32_t));
    if (counts_pos_host) copy_out(ctx, counts_pos_host, e->null_cnt.p + cells, cells * sizeof(uint32_t));
    SB_CUDA(cudaStreamSynchronize(ctx->stream));
    SB_API_END
}

int sb_enrich_null_counts_dev(sb_enrich* e, uint32_t** counts_neg_dev, uint32_t** counts_pos_dev) {
    SB_API_BEGIN
    SB_CHECK(e && counts_neg_dev && counts_pos_dev, "sb_enrich_null_counts_dev: NULL argument");
    SB_CHECK(e->null_score >= 0, "sb_enrich_null_counts_dev: no null has been started on this plan");
    e->ctx->bind();
    null_flush(e);
    *counts_neg_dev = e->null_cnt.p;
    *counts_pos_dev = e->null_cnt.p + static_cast<size_t>(e->n) * e->m;
    SB_API_END
}

int sb_enrich_null_packed_dev(sb_enrich* e, uint32_t** counts_packed_dev, int64_t* perms_in_packed) {
    SB_API_BEGIN
    SB_CHECK(e && counts_packed_dev, "sb_enrich_null_packed_dev: NULL argument");
    SB_CHECK(e->null_score >= 0, "sb_enrich_null_packed_dev: no null has been started on this plan");
    *counts_packed_dev = e->null_pk.p;
    if (perms_in_packed) *perms_in_packed = e->null_pk_perms;
    SB_API_END
}

int sb_enrich_null_set_perms(sb_enrich* e, int64_t num_perm) {
    SB_API_BEGIN
    SB_CHECK(e && num_perm >= 0, "sb_enrich_null_set_perms: bad argument");
    SB_CHECK(e->null_score >= 0, "sb_enrich_null_set_perms: no null has been started on this plan");
    SB_CHECK(num_perm < 65536 || e->null_pk_perms == 0 || num_perm == e->null_perms,
             "sb_enrich_null_set_perms: %lld permutations do not fit the packed counters (sum the unpacked arrays)",
             (long long)num_perm);
    if (num_perm != e->null_perms) e->null_pk_perms = std::min<int64_t>(num_perm, 65535);  // summed over the ranks
    e->null_perms = num_perm;
    SB_API_END
}

int sb_enrich_null_finalize(sb_enrich* e, const double* pvalue_of_count_host, const double* nes_of_count_host,
                            int64_t table_len, int multiple_testing, double zero_pvalue_floor, int attribute_sign,
                            double nes_threshold, double* ns_host, double* pvalues_neg_host, double* pvalues_pos_host,
                            double* nes_host, double* nes_binary_host, double* num_enriched_host) {
    SB_API_BEGIN
    SB_CHECK(e && pvalue_of_count_host && nes_of_count_host, "sb_enrich_null_finalize: NULL argument");
    SB_CHECK(e->null_score >= 0, "sb_enrich_null_finalize: no null has been started on this plan");
    SB_CHECK(table_len > e->null_perms && table_len < (1ll << 31),
             "sb_enrich_null_finalize: tables hold %lld entries but %lld permutations were counted (need P + 1)",
             (long long)table_len, (long long)e->null_perms);
    SB_CHECK(attribute_sign >= 0 && attribute_sign <= 2, "sb_enrich_null_finalize: attribute_sign must be 0, 1 or 2");
    sb_ctx* ctx = e->ctx;
    ctx->bind();
    cudaStream_t st = ctx->stream;
    PhaseTrace tr(ctx, "null.finalize");
    const int64_t n = e->n, m = e->m;
    const size_t cells = static_cast<size_t>(n) * m;
    null_flush(e);
    const double* ns = enrich_observed(e, e->null_score);
    DevBuf<double> ptab, nestab, pn, pp, nes, nb;
    DevBuf<int32_t> colcnt;
    ptab.reserve(table_len);
    nestab.reserve(table_len);
    SB_CUDA(cudaMemcpyAsync(ptab.p, pvalue_of_count_host, table_len * sizeof(double), cudaMemcpyHostToDevice, st));
    SB_CUDA(cudaMemcpyAsync(nestab.p, nes_of_count_host, table_len * sizeof(double), cudaMemcpyHostToDevice, st));
    colcnt.reserve(m);
    SB_CUDA(cudaMemsetAsync(colcnt.p, 0, m * sizeof(int32_t), st));
    const int64_t rows_per = chunk_rows(n, m);
    pn.reserve(rows_per * m);
    pp.reserve(rows_per * m);
    nes.reserve(rows_per * m);
    nb.reserve(rows_per * m);
    BhScratch bh;
    // with FDR the first pass must not count enriched neighborhoods: give it a scratch counter
    DevBuf<int32_t> colcnt_scratch;
    if (multiple_testing) colcnt_scratch.reserve(m);
    for (int64_t r0 = 0; r0 < n; r0 += rows_per) {
        const int64_t rows = std::min(rows_per, n - r0);
        const size_t at = static_cast<size_t>(r0) * m, len = static_cast<size_t>(rows) * m;
        {
            KernelTimer kt(ctx, SB_K_TAIL);
            k_null_tail<true><<<tail_grid(rows, m), 256, 0, st>>>(
                e->null_cnt.p + at, e->null_cnt.p + cells + at, ns + at, ptab.p, nestab.p,
                static_cast<uint32_t>(table_len), zero_pvalue_floor, attribute_sign, nes_threshold, rows, m, pn.p, pp.p,
                nes.p, nb.p, multiple_testing ? colcnt_scratch.p : colcnt.p);
            SB_LAUNCH_CHECK(ctx);
        }
        if (multiple_testing) {
            bh_rows(ctx, bh, pn.p, rows, m);
            bh_rows(ctx, bh, pp.p, rows, m);
            k_null_tail<false><<<tail_grid(rows, m), 256, 0, st>>>(nullptr, nullptr, nullptr, nullptr, nullptr, 0,
                                                                    zero_pvalue_floor, attribute_sign, nes_threshold,
                                                                    rows, m, pn.p, pp.p, nes.p, nb.p, colcnt.p);
            SB_LAUNCH_CHECK(ctx);
        }
        if (ns_host) copy_out(ctx, ns_host + at, ns + at, len * sizeof(double));
        if (pvalues_neg_host) copy_out(ctx, pvalues_neg_host + at, pn.p, len * sizeof(double));
        if (pvalues_pos_host) copy_out(ctx, pvalues_pos_host + at, pp.p, len * sizeof(double));
        if (nes_host) copy_out(ctx, nes_host + at, nes.p, len * sizeof(double));
        if (nes_binary_host) copy_out(ctx, nes_binary_host + at, nb.p, len * sizeof(double));
    }
    colcnt_out(ctx, colcnt.p, m, num_enriched_host);
    SB_CUDA(cudaStreamSynchronize(st));
    SB_API_END
}

int sb_enrich_attr_summary(sb_enrich* e, int64_t* nan_per_column_host, int64_t* other_values_out) {
    SB_API_BEGIN
    SB_CHECK(e && nan_per_column_host && other_values_out, "sb_enrich_attr_summary: NULL argument");
    sb_ctx* ctx = e->ctx;
    ctx->bind();
    cudaStream_t st = ctx->stream;
    const int64_t n = e->n, m = e->m;
    DevBuf<unsigned long long> acc;
    acc.reserve(m + 1);
    SB_CUDA(cudaMemsetAsync(acc.p, 0, (m + 1) * sizeof(unsigned long long), st));
    const dim3 grid(static_cast<unsigned>(sb_ceil_div(m, 256)), static_cast<unsigned>(sb_ceil_div(n, 64)));
    SB_CHECK(grid.y <= 65535, "sb_enrich_attr_summary: n=%lld too large for one launch", (long long)n);
    if (e->dtype == SB_F32)
        k_attr_summary<float><<<grid, 256, 0, st>>>(static_cast<const float*>(e->b), n, m, acc.p, acc.p + m);
    else
        k_attr_summary<double><<<grid, 256, 0, st>>>(static_cast<const double*>(e->b), n, m, acc.p, acc.p + m);
    SB_LAUNCH_CHECK(ctx);
    std::vector<unsigned long long> h(m + 1);
    SB_CUDA(cudaMemcpyAsync(h.data(), acc.p, (m + 1) * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaStreamSynchronize(st));
    for (int64_t j = 0; j < m; ++j) nan_per_column_host[j] = static_cast<int64_t>(h[j]);
    *other_values_out = static_cast<int64_t>(h[m]);
    SB_API_END
}

// ------------------------------------------------------------------------------------------------ hypergeometric
int sb_enrich_hypergeom_finalize(sb_enrich* e, int multiple_testing, double nes_threshold, double* pvalues_host,
                                 double* nes_host, double* nes_binary_host, double* num_enriched_host) {
    SB_API_BEGIN
    SB_CHECK(e, "sb_enrich_hypergeom_finalize: NULL handle");
    sb_ctx* ctx = e->ctx;
    ctx->bind();
    cudaStream_t st = ctx->stream;
    const int64_t n = e->n, m = e->m;
    const size_t cells = static_cast<size_t>(n) * m;
    DevBuf<double> pv, nes, nb;
    DevBuf<int32_t> colcnt;
    pv.reserve(cells);
    nes.reserve(cells);
    colcnt.reserve(m);
    SB_CUDA(cudaMemsetAsync(colcnt.p, 0, m * sizeof(int32_t), st));
    int rc = sb_enrich_hypergeom_dev(e, pv.p, nes.p);
    if (rc) fail("%s", sb_last_error());
    PhaseTrace tr(ctx, "hypergeom.finalize");
    const int64_t rows_per = chunk_rows(n, m);
    nb.reserve(rows_per * m);
    BhScratch bh;
    for (int64_t r0 = 0; r0 < n; r0 += rows_per) {
        const int64_t rows = std::min(rows_per, n - r0);
        const size_t at = static_cast<size_t>(r0) * m, len = static_cast<size_t>(rows) * m;
        if (multiple_testing) {
            bh_rows(ctx, bh, pv.p + at, rows, m);
            k_binarize<true><<<tail_grid(rows, m), 256, 0, st>>>(pv.p + at, nes.p + at, nes_threshold, rows, m, nb.p,
                                                                  colcnt.p);
        } else {
            k_binarize<false><<<tail_grid(rows, m), 256, 0, st>>>(pv.p + at, nes.p + at, nes_threshold, rows, m, nb.p,
                                                                   colcnt.p);
        }
        SB_LAUNCH_CHECK(ctx);
        if (pvalues_host) copy_out(ctx, pvalues_host + at, pv.p + at, len * sizeof(double));
        if (nes_host) copy_out(ctx, nes_host + at, nes.p + at, len * sizeof(double));
        if (nes_binary_host) copy_out(ctx, nes_binary_host + at, nb.p, len * sizeof(double));
    }
    colcnt_out(ctx, colcnt.p, m, num_enriched_host);
    SB_CUDA(cudaStreamSynchronize(st));
    SB_API_END
}

// ------------------------------------------------------------------------------------------------ stand-alone rows
int sb_fdr_rows(sb_ctx* ctx, int64_t n, int64_t m, const double* pvalues_host, double* adjusted_host) {
    SB_API_BEGIN
    SB_CHECK(ctx && pvalues_host && adjusted_host, "sb_fdr_rows: NULL argument");
    SB_CHECK(n > 0 && m > 0 && m < (1ll << 31), "sb_fdr_rows: bad shape %lld x %lld", (long long)n, (long long)m);
    ctx->bind();
    cudaStream_t st = ctx->stream;
    const int64_t rows_per = chunk_rows(n, m);
    DevBuf<double> p;
    p.reserve(rows_per * m);
    BhScratch bh;
    for (int64_t r0 = 0; r0 < n; r0 += rows_per) {
        const int64_t rows = std::min(rows_per, n - r0);
        const size_t at = static_cast<size_t>(r0) * m, len = static_cast<size_t>(rows) * m;
        copy_in(ctx, p.p, pvalues_host + at, len * sizeof(double));
        bh_rows(ctx, bh, p.p, rows, m);
        copy_out(ctx, adjusted_host + at, p.p, len * sizeof(double));
    }
    SB_CUDA(cudaStreamSynchronize(st));
    SB_API_END
}

int sb_attr_jaccard(sb_ctx* ctx, int64_t n, const uint8_t* member_host, int64_t m, const int32_t* cols_host,
                    int64_t n_cols, double* condensed_out_host) {
    SB_API_BEGIN
    SB_CHECK(ctx && member_host && cols_host, "sb_attr_jaccard: NULL argument");
    SB_CHECK(n > 0 && m > 0 && n_cols >= 0 && n_cols < (1ll << 31), "sb_attr_jaccard: bad shape");
    for (int64_t k = 0; k < n_cols; ++k)
        SB_CHECK(cols_host[k] >= 0 && cols_host[k] < m, "sb_attr_jaccard: column %d out of range", cols_host[k]);
    if (n_cols < 2) return 0;
    SB_CHECK(condensed_out_host, "sb_attr_jaccard: NULL output");
    ctx->bind();
    cudaStream_t st = ctx->stream;
    const int64_t words = sb_ceil_div(n, 32);
    SB_CHECK(words <= 65535, "sb_attr_jaccard: n=%lld too large", (long long)n);
    const int64_t pairs = n_cols * (n_cols - 1) / 2;
    DevBuf<uint8_t> member;
    DevBuf<int32_t> cols;
    DevBuf<uint32_t> bits;
    DevBuf<double> out;
    member.reserve(static_cast<size_t>(n) * m);
    cols.reserve(n_cols);
    bits.reserve(static_cast<size_t>(n_cols) * words);
    out.reserve(pairs);
    copy_in(ctx, member.p, member_host, static_cast<size_t>(n) * m);
    SB_CUDA(cudaMemcpyAsync(cols.p, cols_host, n_cols * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    k_pack_columns<<<dim3(static_cast<unsigned>(sb_ceil_div(n_cols, 128)), static_cast<unsigned>(words)), 128, 0, st>>>(
        member.p, n, m, cols.p, n_cols, words, bits.p);
    SB_LAUNCH_CHECK(ctx);
    const int64_t jgroups = sb_ceil_div(n_cols, 8);
    SB_CHECK(jgroups <= 65535, "sb_attr_jaccard: %lld attributes are too many for one launch", (long long)n_cols);
    {
        KernelTimer kt(ctx, SB_K_JACCARD);
        k_jaccard<<<dim3(static_cast<unsigned>(n_cols), static_cast<unsigned>(jgroups)), 256, 0, st>>>(bits.p, n_cols,
                                                                                                      words, out.p);
        SB_LAUNCH_CHECK(ctx);
    }
    copy_out(ctx, condensed_out_host, out.p, pairs * sizeof(double));
    SB_CUDA(cudaStreamSynchronize(st));
    SB_API_END
}

}  // extern "C"
