// Graph-side helpers around the hot path:
//   k_edge_len      safe_io.calculate_edge_lengths (reference safepy/safe_io.py:311-333): layout distance x adjacency
//                   weight per edge, O(E) instead of the reference's dense N x N product
//   CSR build       symmetric CSR of an undirected edge list (what networkx's Dijkstra iterates), columns ascending
//   k_components    SAFE.define_top_attributes' unimodality test (reference safepy/safe.py:632-658): connected
//                   components of the subgraph induced by the enriched nodes of every candidate attribute, batched
//                   over attributes (one CTA per attribute, min-label propagation with pointer jumping)
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace sb {

// length = sqrt_rn(dx*dx + dy*dy) * w with every operation rounded separately (scipy's pdist does not fuse, and the
// reference multiplies the distance matrix by the adjacency matrix afterwards).  A zero weight gives NaN: the
// reference turns zeros of the adjacency matrix into NaN and sets no 'length' attribute for those edges.
__global__ void k_edge_len(const double* __restrict__ x, const double* __restrict__ y, const int32_t* __restrict__ eu,
                           const int32_t* __restrict__ ev, const double* __restrict__ w, int64_t n_edges,
                           double* __restrict__ out) {
    int64_t e = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    const int64_t step = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (; e < n_edges; e += step) {
        const int32_t a = eu[e], b = ev[e];
        const double dx = __dsub_rn(x[a], x[b]), dy = __dsub_rn(y[a], y[b]);
        const double d = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
        const double wt = w ? w[e] : 1.0;
        out[e] = wt == 0.0 ? __longlong_as_double(0x7FF8000000000000ll) : __dmul_rn(d, wt);
    }
}

__global__ void k_degree(const int32_t* __restrict__ eu, const int32_t* __restrict__ ev, int64_t n_edges,
                         unsigned long long* __restrict__ deg) {
    int64_t e = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    const int64_t step = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (; e < n_edges; e += step) {
        const int32_t a = eu[e], b = ev[e];
        atomicAdd(&deg[a], 1ull);
        if (a != b) atomicAdd(&deg[b], 1ull);
    }
}

// single-block exclusive scan (n + 1 outputs)
__global__ void __launch_bounds__(1024) k_scan_u64(const unsigned long long* __restrict__ cnt, int64_t n,
                                                    long long* __restrict__ ptr) {
    __shared__ long long part[1024];
    const int t = threadIdx.x;
    const int64_t chunk = (n + 1023) / 1024;
    const int64_t b = t * chunk, e = min(n, b + chunk);
    long long s = 0;
    for (int64_t i = b; i < e; ++i) s += static_cast<long long>(cnt[i]);
    part[t] = s;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        long long v = t >= o ? part[t - o] : 0;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    long long run = t ? part[t - 1] : 0;
    for (int64_t i = b; i < e; ++i) {
        ptr[i] = run;
        run += static_cast<long long>(cnt[i]);
    }
    if (t == 1023) ptr[n] = part[1023];
}

__global__ void k_csr_fill(const int32_t* __restrict__ eu, const int32_t* __restrict__ ev,
                           const double* __restrict__ val, int64_t n_edges, const long long* __restrict__ ptr,
                           unsigned long long* __restrict__ cursor, int32_t* __restrict__ indices,
                           double* __restrict__ vout) {
    int64_t e = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    const int64_t step = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (; e < n_edges; e += step) {
        const int32_t a = eu[e], b = ev[e];
        long long at = ptr[a] + static_cast<long long>(atomicAdd(&cursor[a], 1ull));
        indices[at] = b;
        if (vout) vout[at] = val[e];
        if (a != b) {
            at = ptr[b] + static_cast<long long>(atomicAdd(&cursor[b], 1ull));
            indices[at] = a;
            if (vout) vout[at] = val[e];
        }
    }
}

// columns ascending inside every row (rows are short: insertion sort by one thread per row)
__global__ void k_csr_sort_rows(const long long* __restrict__ ptr, int64_t n, int32_t* __restrict__ indices,
                                double* __restrict__ val) {
    const int64_t row = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (row >= n) return;
    const long long b = ptr[row], e = ptr[row + 1];
    for (long long i = b + 1; i < e; ++i) {
        const int32_t c = indices[i];
        const double v = val ? val[i] : 0.0;
        long long j = i - 1;
        while (j >= b && indices[j] > c) {
            indices[j + 1] = indices[j];
            if (val) val[j + 1] = val[j];
            --j;
        }
        indices[j + 1] = c;
        if (val) val[j + 1] = v;
    }
}

// One CTA per candidate attribute.  label[v] = smallest node id of v's component among the member nodes, -1 outside.
__global__ void __launch_bounds__(512) k_components(const int64_t* __restrict__ indptr,
                                                    const int32_t* __restrict__ indices,
                                                    const uint8_t* __restrict__ member /* [n][m] */, int64_t n,
                                                    int64_t m, const int32_t* __restrict__ cand, int32_t min_size,
                                                    int32_t* __restrict__ labels /* [k][n] */,
                                                    int32_t* __restrict__ sizes /* [k][n] scratch */,
                                                    int32_t* __restrict__ num_cc, int32_t* __restrict__ num_large) {
    const int k = blockIdx.x;
    const int64_t j = cand[k];
    int32_t* lab = labels + static_cast<int64_t>(k) * n;
    int32_t* sz = sizes + static_cast<int64_t>(k) * n;
    __shared__ int s_changed;
    __shared__ int s_cc, s_large;
    for (int64_t v = threadIdx.x; v < n; v += blockDim.x) {
        lab[v] = member[v * m + j] ? static_cast<int32_t>(v) : -1;
        sz[v] = 0;
    }
    __syncthreads();
    while (true) {
        if (threadIdx.x == 0) s_changed = 0;
        __syncthreads();
        // hook: every member node takes the smallest label in its closed neighborhood (members only)
        for (int64_t v = threadIdx.x; v < n; v += blockDim.x) {
            int32_t lv = lab[v];
            if (lv < 0) continue;
            int32_t best = lv;
            for (int64_t e = indptr[v]; e < indptr[v + 1]; ++e) {
                const int32_t lu = lab[indices[e]];
                if (lu >= 0 && lu < best) best = lu;
            }
            if (best < lv) {
                atomicMin(&lab[v], best);
                atomicMin(&lab[lv], best);  // pull the old representative down as well
                s_changed = 1;
            }
        }
        __syncthreads();
        // pointer jumping
        for (int64_t v = threadIdx.x; v < n; v += blockDim.x) {
            int32_t lv = lab[v];
            if (lv < 0) continue;
            int32_t r = lab[lv];
            while (r != lv) {
                lv = r;
                r = lab[lv];
            }
            lab[v] = lv;
        }
        __syncthreads();
        if (!s_changed) break;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        s_cc = 0;
        s_large = 0;
    }
    __syncthreads();
    for (int64_t v = threadIdx.x; v < n; v += blockDim.x) {
        const int32_t lv = lab[v];
        if (lv >= 0) atomicAdd(&sz[lv], 1);
    }
    __syncthreads();
    for (int64_t v = threadIdx.x; v < n; v += blockDim.x) {
        const int32_t s = sz[v];
        if (s > 0) {
            atomicAdd(&s_cc, 1);
            if (s >= min_size) atomicAdd(&s_large, 1);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        num_cc[k] = s_cc;
        num_large[k] = s_large;
    }
}

}  // namespace sb

using namespace sb;

extern "C" {

int sb_graph_edge_lengths(sb_ctx* ctx, int64_t n, const double* x_host, const double* y_host, int64_t n_edges,
                          const int32_t* eu_host, const int32_t* ev_host, const double* weight_host,
                          double* length_out_host) {
    SB_API_BEGIN
    SB_CHECK(ctx && x_host && y_host && (n_edges == 0 || (eu_host && ev_host && length_out_host)),
             "sb_graph_edge_lengths: NULL argument");
    SB_CHECK(n > 0 && n_edges >= 0, "sb_graph_edge_lengths: bad sizes");
    ctx->bind();
    if (n_edges == 0) return 0;
    for (int64_t e = 0; e < n_edges; ++e)
        SB_CHECK(eu_host[e] >= 0 && eu_host[e] < n && ev_host[e] >= 0 && ev_host[e] < n,
                 "sb_graph_edge_lengths: edge %lld has an endpoint outside 0..n-1", (long long)e);
    cudaStream_t st = ctx->stream;
    DevBuf<double> xy, w, out;
    DevBuf<int32_t> uv;
    xy.reserve(2 * n);
    uv.reserve(2 * n_edges);
    out.reserve(n_edges);
    SB_CUDA(cudaMemcpyAsync(xy.p, x_host, n * sizeof(double), cudaMemcpyHostToDevice, st));
    SB_CUDA(cudaMemcpyAsync(xy.p + n, y_host, n * sizeof(double), cudaMemcpyHostToDevice, st));
    SB_CUDA(cudaMemcpyAsync(uv.p, eu_host, n_edges * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    SB_CUDA(cudaMemcpyAsync(uv.p + n_edges, ev_host, n_edges * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    if (weight_host) {
        w.reserve(n_edges);
        SB_CUDA(cudaMemcpyAsync(w.p, weight_host, n_edges * sizeof(double), cudaMemcpyHostToDevice, st));
    }
    const unsigned blocks = static_cast<unsigned>(std::min<int64_t>(sb_ceil_div(n_edges, 256), ctx->num_sms * 16));
    k_edge_len<<<blocks, 256, 0, st>>>(xy.p, xy.p + n, uv.p, uv.p + n_edges, weight_host ? w.p : nullptr, n_edges, out.p);
    SB_LAUNCH_CHECK(ctx);
    SB_CUDA(cudaMemcpyAsync(length_out_host, out.p, n_edges * sizeof(double), cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaStreamSynchronize(st));
    SB_API_END
}

int sb_graph_csr(sb_ctx* ctx, int64_t n, int64_t n_edges, const int32_t* eu_host, const int32_t* ev_host,
                 const double* value_host, int64_t* indptr_out_host, int32_t* indices_out_host,
                 double* value_out_host, int64_t* nnz_out) {
    SB_API_BEGIN
    SB_CHECK(ctx && indptr_out_host && nnz_out && (n_edges == 0 || (eu_host && ev_host && indices_out_host)),
             "sb_graph_csr: NULL argument");
    SB_CHECK(n > 0 && n_edges >= 0, "sb_graph_csr: bad sizes");
    SB_CHECK(!value_host || value_out_host || n_edges == 0, "sb_graph_csr: value_out is NULL");
    ctx->bind();
    for (int64_t e = 0; e < n_edges; ++e)
        SB_CHECK(eu_host[e] >= 0 && eu_host[e] < n && ev_host[e] >= 0 && ev_host[e] < n,
                 "sb_graph_csr: edge %lld has an endpoint outside 0..n-1", (long long)e);
    cudaStream_t st = ctx->stream;
    DevBuf<unsigned long long> deg, cursor;
    DevBuf<long long> ptr;
    DevBuf<int32_t> uv, idx;
    DevBuf<double> val, vout;
    deg.reserve(n);
    cursor.reserve(n);
    ptr.reserve(n + 1);
    uv.reserve(std::max<int64_t>(2 * n_edges, 1));
    idx.reserve(std::max<int64_t>(2 * n_edges, 1));
    SB_CUDA(cudaMemsetAsync(deg.p, 0, n * sizeof(unsigned long long), st));
    SB_CUDA(cudaMemsetAsync(cursor.p, 0, n * sizeof(unsigned long long), st));
    if (n_edges) {
        SB_CUDA(cudaMemcpyAsync(uv.p, eu_host, n_edges * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        SB_CUDA(cudaMemcpyAsync(uv.p + n_edges, ev_host, n_edges * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    }
    if (value_host && n_edges) {
        val.reserve(n_edges);
        vout.reserve(2 * n_edges);
        SB_CUDA(cudaMemcpyAsync(val.p, value_host, n_edges * sizeof(double), cudaMemcpyHostToDevice, st));
    }
    const unsigned blocks =
        static_cast<unsigned>(std::max<int64_t>(1, std::min<int64_t>(sb_ceil_div(n_edges, 256), ctx->num_sms * 16)));
    if (n_edges) {
        k_degree<<<blocks, 256, 0, st>>>(uv.p, uv.p + n_edges, n_edges, deg.p);
        SB_LAUNCH_CHECK(ctx);
    }
    k_scan_u64<<<1, 1024, 0, st>>>(deg.p, n, ptr.p);
    SB_LAUNCH_CHECK(ctx);
    if (n_edges) {
        k_csr_fill<<<blocks, 256, 0, st>>>(uv.p, uv.p + n_edges, value_host ? val.p : nullptr, n_edges, ptr.p, cursor.p,
                                           idx.p, value_host ? vout.p : nullptr);
        SB_LAUNCH_CHECK(ctx);
        k_csr_sort_rows<<<static_cast<unsigned>(sb_ceil_div(n, 128)), 128, 0, st>>>(ptr.p, n, idx.p,
                                                                                   value_host ? vout.p : nullptr);
        SB_LAUNCH_CHECK(ctx);
    }
    static_assert(sizeof(long long) == sizeof(int64_t), "indptr type");
    SB_CUDA(cudaMemcpyAsync(indptr_out_host, ptr.p, (n + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaStreamSynchronize(st));
    const int64_t nnz = indptr_out_host[n];
    *nnz_out = nnz;
    if (nnz) {
        SB_CUDA(cudaMemcpyAsync(indices_out_host, idx.p, nnz * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        if (value_host)
            SB_CUDA(cudaMemcpyAsync(value_out_host, vout.p, nnz * sizeof(double), cudaMemcpyDeviceToHost, st));
        SB_CUDA(cudaStreamSynchronize(st));
    }
    SB_API_END
}

int sb_graph_components(sb_ctx* ctx, int64_t n, const int64_t* indptr_host, const int32_t* indices_host,
                        const uint8_t* member_host, int64_t m, const int32_t* cand_host, int64_t n_cand,
                        int32_t min_size, int32_t* labels_out_host, int32_t* num_cc_out_host,
                        int32_t* num_large_out_host) {
    SB_API_BEGIN
    SB_CHECK(ctx && indptr_host && member_host && (n_cand == 0 || (cand_host && num_cc_out_host && num_large_out_host)),
             "sb_graph_components: NULL argument");
    SB_CHECK(n > 0 && m > 0 && n_cand >= 0, "sb_graph_components: bad sizes");
    ctx->bind();
    if (n_cand == 0) return 0;
    const int64_t nnz = indptr_host[n];
    SB_CHECK(nnz == 0 || indices_host, "sb_graph_components: indices is NULL");
    for (int64_t k = 0; k < n_cand; ++k)
        SB_CHECK(cand_host[k] >= 0 && cand_host[k] < m, "sb_graph_components: attribute index %d out of range", cand_host[k]);
    cudaStream_t st = ctx->stream;
    DevBuf<int64_t> ptr;
    DevBuf<int32_t> idx, cand, labels, sizes, ncc, nlarge;
    DevBuf<uint8_t> mem;
    ptr.reserve(n + 1);
    idx.reserve(std::max<int64_t>(nnz, 1));
    mem.reserve(static_cast<size_t>(n) * m);
    SB_CUDA(cudaMemcpyAsync(ptr.p, indptr_host, (n + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    if (nnz) SB_CUDA(cudaMemcpyAsync(idx.p, indices_host, nnz * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    copy_in(ctx, mem.p, member_host, static_cast<size_t>(n) * m);
    // candidates are processed in chunks so that the label / size scratch stays below ~2 GiB
    const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(n_cand, (1ll << 28) / n));
    labels.reserve(static_cast<size_t>(chunk) * n);
    sizes.reserve(static_cast<size_t>(chunk) * n);
    cand.reserve(chunk);
    ncc.reserve(chunk);
    nlarge.reserve(chunk);
    for (int64_t k0 = 0; k0 < n_cand; k0 += chunk) {
        const int64_t kc = std::min(chunk, n_cand - k0);
        SB_CUDA(cudaMemcpyAsync(cand.p, cand_host + k0, kc * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        k_components<<<static_cast<unsigned>(kc), 512, 0, st>>>(ptr.p, idx.p, mem.p, n, m, cand.p, min_size, labels.p,
                                                                 sizes.p, ncc.p, nlarge.p);
        SB_LAUNCH_CHECK(ctx);
        SB_CUDA(cudaMemcpyAsync(num_cc_out_host + k0, ncc.p, kc * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        SB_CUDA(cudaMemcpyAsync(num_large_out_host + k0, nlarge.p, kc * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        if (labels_out_host)
            SB_CUDA(cudaMemcpyAsync(labels_out_host + k0 * n, labels.p, static_cast<size_t>(kc) * n * sizeof(int32_t),
                                    cudaMemcpyDeviceToHost, st));
        SB_CUDA(cudaStreamSynchronize(st));
    }
    SB_API_END
}

}  // extern "C"
