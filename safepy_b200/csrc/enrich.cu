// Stage 2 on CUDA cores: CSR view of the packed matrix, fp64 neighborhood scores, the exact (fp64,
// ascending-neighbor order) permutation-count kernel used for validation / z-score / GEMM fix-ups, and the
// fused hypergeometric survival-function kernel.
// Replaces compute_neighborhood_score + run_permutations (reference safepy/safe_extras.py:6-70) and the body of
// SAFE.compute_pvalues_by_hypergeom (safepy/safe.py:573-608).
#include <cmath>
#include <vector>

#include "enrich.cuh"

namespace sb {

// ------------------------------------------------------------------------------------------------ CSR build
__global__ void k_rowcount(const uint32_t* __restrict__ words, int64_t n, int64_t ld, int64_t* __restrict__ cnt) {
    const int64_t row = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    int64_t acc = 0;
    for (int64_t w = lane; w < ld; w += 32) acc += __popc(words[row * ld + w]);
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) cnt[row] = acc;
}

// single-block exclusive scan: row_ptr[0..n] from cnt[0..n-1] (cnt aliases row_ptr + 1 is NOT allowed)
__global__ void __launch_bounds__(1024) k_exscan(const int64_t* __restrict__ cnt, int64_t n,
                                                  int64_t* __restrict__ row_ptr) {
    __shared__ int64_t part[1024];
    const int t = threadIdx.x;
    const int64_t chunk = (n + 1023) / 1024;
    const int64_t b = t * chunk, e = min(n, b + chunk);
    int64_t s = 0;
    for (int64_t i = b; i < e; ++i) s += cnt[i];
    part[t] = s;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        int64_t v = t >= o ? part[t - o] : 0;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    int64_t run = t ? part[t - 1] : 0;
    for (int64_t i = b; i < e; ++i) {
        row_ptr[i] = run;
        run += cnt[i];
    }
    if (t == 1023) row_ptr[n] = part[1023];
}

__global__ void k_fill_csr(const uint32_t* __restrict__ words, int64_t n, int64_t ld,
                           const int64_t* __restrict__ row_ptr, int32_t* __restrict__ col_idx) {
    const int64_t row = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    int64_t pos = row_ptr[row];
    for (int64_t w0 = 0; w0 < ld; w0 += 32) {
        const int64_t w = w0 + lane;
        uint32_t bits = w < ld ? words[row * ld + w] : 0u;
        int c = __popc(bits);
        int incl = c;
        for (int o = 1; o < 32; o <<= 1) {
            int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        int64_t at = pos + incl - c;
        while (bits) {
            int b = __ffs(bits) - 1;
            bits &= bits - 1;
            col_idx[at++] = static_cast<int32_t>(w * 32 + b);
        }
        pos += __shfl_sync(0xffffffffu, incl, 31);
    }
}

// ------------------------------------------------------------------------------------------------ scores

template <class T, bool ZS>
__global__ void __launch_bounds__(128) k_score(const int64_t* __restrict__ row_ptr,
                                               const int32_t* __restrict__ col_idx, const T* __restrict__ b,
                                               int64_t n, int64_t m, int64_t row0, double* __restrict__ out) {
    const int64_t i = row0 + blockIdx.x;
    const int64_t j = blockIdx.y * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (j >= m) return;
    out[i * m + j] = score_one<T, ZS>(row_ptr, col_idx, b, nullptr, n, m, i, 0, j);
}

template <class T, bool ZS>
__global__ void __launch_bounds__(128) k_perm_count(const int64_t* __restrict__ row_ptr,
                                                    const int32_t* __restrict__ col_idx, const T* __restrict__ b,
                                                    const int32_t* __restrict__ perm, int64_t n, int64_t m,
                                                    int64_t ncols, const double* __restrict__ s0,
                                                    uint32_t* __restrict__ cneg, uint32_t* __restrict__ cpos,
                                                    uint32_t* __restrict__ packed) {
    const int64_t i = blockIdx.x;
    const int64_t c = blockIdx.y * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (c >= ncols) return;
    const int64_t p = c / m, j = c % m;
    const double s = score_one<T, ZS>(row_ptr, col_idx, b, perm, n, m, i, p, j);
    const double o = s0[i * m + j];
    // safe_extras.py:65-66 (NaN compares false on both sides)
    if (packed) {
        const uint32_t inc = (s <= o ? 1u : 0u) + (s >= o ? 0x10000u : 0u);
        if (inc) atomicAdd(&packed[i * m + j], inc);
        return;
    }
    if (s <= o) atomicAdd(&cneg[i * m + j], 1u);
    if (s >= o) atomicAdd(&cpos[i * m + j], 1u);
}

// Exact re-evaluation of the comparisons the fixed-point GEMM could not decide.  One warp per flagged
// (node i, attribute j, permutation p): the lanes stride over the neighborhood, accumulate the permuted AND the
// observed score in fp64 and combine the 32 partial sums in a fixed butterfly order.  (For the input classes whose
// counts are pinned against the reference -- binary, integer, dyadic, float32-valued data -- these sums are exact in
// fp64, so the summation order is immaterial; for the rest the reference's own BLAS order is unspecified.)
// b_t is the attribute matrix TRANSPOSED ([m][n]): the scattered reads of one comparison then fall into one
// n-element column (a bucket of 64 columns stays L2-resident even at 100k nodes x 5000 attributes).
// blockIdx.y selects a bucket of the list when count_dev is given.
template <class T>
__global__ void __launch_bounds__(256) k_fixup(const int64_t* __restrict__ row_ptr, const int32_t* __restrict__ col_idx,
                                               const T* __restrict__ b_t, const int32_t* __restrict__ perm, int64_t n,
                                               int64_t m, const uint64_t* __restrict__ flag_ij,
                                               const uint32_t* __restrict__ flag_p,
                                               unsigned int total, const unsigned int* __restrict__ count_dev,
                                               int n_bucket_lists, unsigned int cap, uint32_t* __restrict__ cneg,
                                               uint32_t* __restrict__ cpos, uint32_t* __restrict__ packed) {
    const int lane = threadIdx.x & 31;
    // Bucketed list: the whole grid walks the buckets one after the other, so that at any time the scattered reads
    // fall into one column group's slab of b_t (64 columns x n values, L2-resident) -- with all buckets in flight at
    // once the kernel was DRAM-bound on 32-byte sectors fetched for 4-byte values (ncu: 11 GB per C3 batch).
    const int n_buckets = count_dev ? n_bucket_lists : 1;
    for (int bucket = 0; bucket < n_buckets; ++bucket) {
    const uint64_t* fij = flag_ij;
    const uint32_t* fp = flag_p;
    if (count_dev) {
        total = count_dev[bucket];
        if (total > cap) continue;  // overflowed bucket: redone by the caller
        fij += static_cast<size_t>(bucket) * cap;
        fp += static_cast<size_t>(bucket) * cap;
    }
    unsigned int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned int step = (gridDim.x * blockDim.x) >> 5;
    for (; k < total; k += step) {
        const uint64_t ij = fij[k];
        const int64_t i = static_cast<int64_t>(ij >> 32), j = static_cast<int64_t>(ij & 0xffffffffu);
        const int32_t* pr = perm + static_cast<int64_t>(fp[k]) * n;
        const T* col = b_t + j * n;
        double sp = 0.0, so = 0.0;
        // four neighbors per lane and step: the index -> permutation -> value loads of a neighbor depend on each
        // other, so the independent chains in flight are what hides the latency (a bucket of a batch holds only a
        // few hundred flags: the walk is latency-bound, not throughput-bound)
        const int64_t e1 = row_ptr[i + 1];
        for (int64_t e = row_ptr[i] + lane; e < e1; e += 128) {
            int32_t t[4];
            T vo[4], vp[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) t[u] = e + 32 * u < e1 ? col_idx[e + 32 * u] : -1;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                vo[u] = t[u] >= 0 ? col[t[u]] : static_cast<T>(0);
                vp[u] = t[u] >= 0 ? col[pr[t[u]]] : static_cast<T>(0);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (vo[u] == vo[u]) so += static_cast<double>(vo[u]);
                if (vp[u] == vp[u]) sp += static_cast<double>(vp[u]);
            }
        }
        for (int o = 16; o; o >>= 1) {
            sp += __shfl_xor_sync(0xffffffffu, sp, o);
            so += __shfl_xor_sync(0xffffffffu, so, o);
        }
        if (lane == 0) {
            if (packed) {
                const uint32_t inc = (sp <= so ? 1u : 0u) + (sp >= so ? 0x10000u : 0u);
                if (inc) atomicAdd(&packed[i * m + j], inc);
            } else {
                if (sp <= so) atomicAdd(&cneg[i * m + j], 1u);
                if (sp >= so) atomicAdd(&cpos[i * m + j], 1u);
            }
        }
    }
    }
}

// ------------------------------------------------------------------------------------------------ hypergeometric
template <class T>
__global__ void k_has_data(const T* __restrict__ b, int64_t n, int64_t m, int32_t* __restrict__ has) {
    const int64_t row = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    int any = 0;
    for (int64_t j = lane; j < m; j += 32) {
        T v = b[row * m + j];
        any |= (v == v);
    }
    any = __any_sync(0xffffffffu, any);
    if (lane == 0) has[row] = any;
}

// per-column nansum in fp64 (np.nansum(node2attribute, axis=0), safe.py:583); rows are split over blockIdx.y
template <class T>
__global__ void k_colsum(const T* __restrict__ b, int64_t n, int64_t m, double* __restrict__ colsum) {
    const int64_t j = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (j >= m) return;
    const int64_t rows_per = (n + gridDim.y - 1) / gridDim.y;
    const int64_t r0 = blockIdx.y * rows_per, r1 = min(n, r0 + rows_per);
    double s = 0.0;
    for (int64_t r = r0; r < r1; ++r) {
        T v = b[r * m + j];
        if (v == v) s += static_cast<double>(v);
    }
    atomicAdd(&colsum[j], s);
}

__global__ void k_nneigh(const int64_t* __restrict__ row_ptr, const int32_t* __restrict__ col_idx,
                         const int32_t* __restrict__ has, int64_t n, double* __restrict__ nneigh) {
    const int64_t row = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    int64_t acc = 0;
    for (int64_t e = row_ptr[row] + lane; e < row_ptr[row + 1]; e += 32) acc += has[col_idx[e]];
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) nneigh[row] = static_cast<double>(acc);
}

__device__ __forceinline__ bool is_integral(double v) { return v == floor(v) && isfinite(v); }

// log pmf of hypergeom(M_total = N, n_group = r, N_draws = n) at x through a log-factorial table
__device__ __forceinline__ double hg_pmf(const double* __restrict__ lf, int64_t x, int64_t r, int64_t n, int64_t N) {
    const double lp = (lf[r] - lf[x] - lf[r - x]) + (lf[N - r] - lf[n - x] - lf[N - r - n + x]) -
                      (lf[N] - lf[n] - lf[N - n]);
    return exp(lp);
}

// scipy.stats.hypergeom.sf(k, M, n, N) = rv_discrete.sf masking + Boost.Math cdf(complement(...)):
// the smaller tail is summed with the pmf ratio recurrence from a single seeded pmf value.
__device__ double hg_sf(const double* __restrict__ lf, double k, double Mt, double nK, double Nn) {
    const double nan = __longlong_as_double(0x7FF8000000000000ll);
    const bool ok = (Mt > 0) && (nK >= 0) && (Nn >= 0) && (nK <= Mt) && (Nn <= Mt) && is_integral(Mt) &&
                    is_integral(nK) && is_integral(Nn);
    if (!ok || k != k) return nan;
    const double a = fmax(Nn - (Mt - nK), 0.0), bsup = fmin(nK, Nn);
    if (k < a) return 1.0;
    if (!(k < bsup) || !isfinite(k)) return 0.0;
    int64_t x = static_cast<int64_t>(floor(k));
    const int64_t r = static_cast<int64_t>(nK), n = static_cast<int64_t>(Nn), N = static_cast<int64_t>(Mt);
    const double eps = 2.220446049250313e-16;
    const double mode = floor(static_cast<double>(r + 1) * static_cast<double>(n + 1) / static_cast<double>(N + 2));
    double result;
    if (static_cast<double>(x) < mode) {
        result = hg_pmf(lf, x, r, n, N);
        double diff = result;
        const int64_t lower = max(static_cast<int64_t>(0), n + r - N);
        // the four factors are carried as doubles (exact below 2^53) and the quotient of a step does not depend on
        // the running term, so the loop-carried chain is one multiply and one add
        double fa = static_cast<double>(x), fb = static_cast<double>((N + x) - n - r);
        double fc = static_cast<double>(1 + n - x), fd = static_cast<double>(1 + r - x);
        while (diff > eps) {
            diff *= (fa * fb) * __drcp_rn(fc * fd);
            result += diff;
            if (x == lower) break;
            --x;
            fa -= 1.0, fb -= 1.0, fc += 1.0, fd += 1.0;
        }
        result = 1.0 - result;
    } else {
        const int64_t upper = min(r, n);
        result = 0.0;
        if (x != upper) {
            ++x;
            result = hg_pmf(lf, x, r, n, N);
            double diff = result;
            double fa = static_cast<double>(n - x), fb = static_cast<double>(r - x);
            double fc = static_cast<double>(x + 1), fd = static_cast<double>((N + x + 1) - n - r);
            while (x <= upper && diff > result * eps) {
                diff *= (fa * fb) * __drcp_rn(fc * fd);
                result += diff;
                ++x;
                fa -= 1.0, fb -= 1.0, fc += 1.0, fd += 1.0;
            }
        }
    }
    return fmin(fmax(result, 0.0), 1.0);
}

// X: the observed counts either as fp64 scores (k_score) or, FIXED, straight from the tensor-core plan's exact
// fixed-point scores (row order of the plan, per-column binary exponent)
template <bool FIXED>
__global__ void __launch_bounds__(256) k_hypergeom(const double* __restrict__ X, const int64_t* __restrict__ xfix,
                                                   const int32_t* __restrict__ shift,
                                                   const int32_t* __restrict__ row_of_node, int64_t mpad,
                                                   const double* __restrict__ colsum,
                                                   const double* __restrict__ nneigh, const double* __restrict__ lf,
                                                   double n_total, int64_t n, int64_t m, double* __restrict__ pv,
                                                   double* __restrict__ nes) {
    const int64_t total = n * m;
    int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    const int64_t step = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (; idx < total; idx += step) {
        const int64_t i = idx / m, j = idx % m;
        double x;
        if (FIXED) {
            const int64_t row = row_of_node ? row_of_node[i] : i;
            x = ldexp(static_cast<double>(xfix[row * mpad + j]), -shift[j]);
        } else {
            x = X[idx];
        }
        const double p = hg_sf(lf, x - 1.0, n_total, colsum[j], nneigh[i]);
        if (pv) pv[idx] = p;
        if (nes) nes[idx] = -log10(p);
    }
}

__global__ void k_sum_i32(const int32_t* __restrict__ v, int64_t n, unsigned long long* __restrict__ out) {
    int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    const int64_t step = static_cast<int64_t>(gridDim.x) * blockDim.x;
    unsigned long long s = 0;
    for (; i < n; i += step) s += static_cast<unsigned long long>(v[i]);
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(out, s);
}

// packed word per cell (pos << 16 | neg) -> the two count arrays (stored, not added)
__global__ void k_unpack_store(const uint32_t* __restrict__ packed, int64_t cells, uint32_t* __restrict__ cneg,
                               uint32_t* __restrict__ cpos) {
    int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    const int64_t step = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (; i < cells; i += step) {
        const uint32_t v = packed[i];
        cneg[i] = v & 0xffffu;
        cpos[i] = v >> 16;
    }
}

// ------------------------------------------------------------------------------------------------ host side
static void build_csr(sb_enrich* e) {
    sb_ctx* ctx = e->ctx;
    PhaseTrace tr(ctx, "enrich.build_csr");
    const int64_t n = e->n;
    DevBuf<int64_t> cnt;
    cnt.reserve(n);
    e->row_ptr.reserve(n + 1);
    const unsigned wblocks = static_cast<unsigned>(sb_ceil_div(n * 32, 256));
    k_rowcount<<<wblocks, 256, 0, ctx->stream>>>(e->a->words, n, e->a->ld, cnt.p);
    SB_LAUNCH_CHECK(ctx);
    k_exscan<<<1, 1024, 0, ctx->stream>>>(cnt.p, n, e->row_ptr.p);
    SB_LAUNCH_CHECK(ctx);
    int64_t nnz = 0;
    SB_CUDA(cudaMemcpyAsync(&nnz, e->row_ptr.p + n, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    SB_CUDA(cudaStreamSynchronize(ctx->stream));
    e->nnz = nnz;
    e->col_idx.reserve(std::max<int64_t>(nnz, 1));
    k_fill_csr<<<wblocks, 256, 0, ctx->stream>>>(e->a->words, n, e->a->ld, e->row_ptr.p, e->col_idx.p);
    SB_LAUNCH_CHECK(ctx);
}

void enrich_score_into(sb_enrich* e, int score_type, double* out_dev) { enrich_score_rows(e, score_type, out_dev, 0, e->n); }

void enrich_score_rows(sb_enrich* e, int score_type, double* out_dev, int64_t row0, int64_t row1) {
    sb_ctx* ctx = e->ctx;
    SB_CHECK(score_type == SB_SCORE_SUM || score_type == SB_SCORE_ZSCORE, "unknown neighborhood_score_type %d",
             score_type);
    SB_CHECK(row0 >= 0 && row0 <= row1 && row1 <= e->n, "score rows [%lld, %lld) out of range", (long long)row0,
             (long long)row1);
    if (row1 == row0) return;
    dim3 grid(static_cast<unsigned>(row1 - row0), static_cast<unsigned>(sb_ceil_div(e->m, 128)));
    SB_CHECK(grid.y <= 65535, "attribute count %lld too large for one launch", (long long)e->m);
    const bool zs = score_type == SB_SCORE_ZSCORE;
    KernelTimer kt(ctx, SB_K_SCORE);
    if (e->dtype == SB_F32) {
        const float* b = static_cast<const float*>(e->b);
        if (zs)
            k_score<float, true><<<grid, 128, 0, ctx->stream>>>(e->row_ptr.p, e->col_idx.p, b, e->n, e->m, row0,
                                                                out_dev);
        else
            k_score<float, false><<<grid, 128, 0, ctx->stream>>>(e->row_ptr.p, e->col_idx.p, b, e->n, e->m, row0,
                                                                 out_dev);
    } else {
        const double* b = static_cast<const double*>(e->b);
        if (zs)
            k_score<double, true><<<grid, 128, 0, ctx->stream>>>(e->row_ptr.p, e->col_idx.p, b, e->n, e->m, row0,
                                                                 out_dev);
        else
            k_score<double, false><<<grid, 128, 0, ctx->stream>>>(e->row_ptr.p, e->col_idx.p, b, e->n, e->m, row0,
                                                                  out_dev);
    }
    SB_LAUNCH_CHECK(ctx);
}

const double* enrich_observed(sb_enrich* e, int score_type) {
    if (score_type == SB_SCORE_SUM) {
        if (!e->have_s0_sum) {
            e->s0_sum.reserve(static_cast<size_t>(e->n) * e->m);
            enrich_score_into(e, SB_SCORE_SUM, e->s0_sum.p);
            e->have_s0_sum = true;
        }
        return e->s0_sum.p;
    }
    if (!e->have_s0_z) {
        e->s0_z.reserve(static_cast<size_t>(e->n) * e->m);
        enrich_score_into(e, SB_SCORE_ZSCORE, e->s0_z.p);
        e->have_s0_z = true;
    }
    return e->s0_z.p;
}

void simt_perm_counts(sb_enrich* e, int score_type, const int32_t* perm_dev, int64_t num_perm, uint32_t* cneg,
                      uint32_t* cpos, uint32_t* packed) {
    sb_ctx* ctx = e->ctx;
    const double* s0 = enrich_observed(e, score_type);
    const bool zs = score_type == SB_SCORE_ZSCORE;
    const int64_t max_cols = 65535ll * 128;
    const int64_t pb = std::max<int64_t>(1, std::min(num_perm, max_cols / e->m));
    SB_CHECK(e->m <= max_cols, "attribute count %lld too large", (long long)e->m);
    for (int64_t p0 = 0; p0 < num_perm; p0 += pb) {
        const int64_t np = std::min(pb, num_perm - p0);
        const int64_t ncols = np * e->m;
        dim3 grid(static_cast<unsigned>(e->n), static_cast<unsigned>(sb_ceil_div(ncols, 128)));
        const int32_t* perm = perm_dev + p0 * e->n;
#define SB_LAUNCH_PC(T, Z)                                                                                     \
    k_perm_count<T, Z><<<grid, 128, 0, ctx->stream>>>(e->row_ptr.p, e->col_idx.p, static_cast<const T*>(e->b), \
                                                      perm, e->n, e->m, ncols, s0, cneg, cpos, packed)
        if (e->dtype == SB_F32) {
            if (zs)
                SB_LAUNCH_PC(float, true);
            else
                SB_LAUNCH_PC(float, false);
        } else {
            if (zs)
                SB_LAUNCH_PC(double, true);
            else
                SB_LAUNCH_PC(double, false);
        }
#undef SB_LAUNCH_PC
        SB_LAUNCH_CHECK(ctx);
    }
}

// tiled transpose [n][m] -> [m][n]
template <class T>
__global__ void __launch_bounds__(256) k_transpose(const T* __restrict__ in, int64_t n, int64_t m, T* __restrict__ out) {
    __shared__ T tile[32][33];
    const int64_t r0 = static_cast<int64_t>(blockIdx.y) * 32, c0 = static_cast<int64_t>(blockIdx.x) * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int k = ty; k < 32; k += 8) {
        const int64_t r = r0 + k, c = c0 + tx;
        if (r < n && c < m) tile[k][tx] = in[r * m + c];
    }
    __syncthreads();
    for (int k = ty; k < 32; k += 8) {
        const int64_t c = c0 + k, r = r0 + tx;
        if (r < n && c < m) out[c * n + r] = tile[tx][k];
    }
}

const void* enrich_transposed(sb_enrich* e) {
    if (e->b_t) return e->b_t;
    sb_ctx* ctx = e->ctx;
    const size_t bytes = static_cast<size_t>(e->n) * e->m * (e->dtype == SB_F32 ? 4 : 8);
    e->b_t = dev_alloc(bytes);
    dim3 grid(static_cast<unsigned>(sb_ceil_div(e->m, 32)), static_cast<unsigned>(sb_ceil_div(e->n, 32)));
    SB_CHECK(grid.y <= 65535, "attribute matrix has too many rows for the transpose kernel");
    if (e->dtype == SB_F32)
        k_transpose<float><<<grid, 256, 0, ctx->stream>>>(static_cast<const float*>(e->b), e->n, e->m,
                                                          static_cast<float*>(e->b_t));
    else
        k_transpose<double><<<grid, 256, 0, ctx->stream>>>(static_cast<const double*>(e->b), e->n, e->m,
                                                           static_cast<double*>(e->b_t));
    SB_LAUNCH_CHECK(ctx);
    return e->b_t;
}

void fixup_flags(sb_enrich* e, const int32_t* perm_dev, const uint64_t* flag_ij, const uint32_t* flag_p,
                 unsigned int count, uint32_t* cneg, uint32_t* cpos, uint32_t* packed) {
    sb_ctx* ctx = e->ctx;
    const void* bt = enrich_transposed(e);
    const unsigned blocks = static_cast<unsigned>(
        std::min<int64_t>(sb_ceil_div(static_cast<int64_t>(count), 8), static_cast<int64_t>(ctx->num_sms) * 8));
    KernelTimer kt(ctx, SB_K_FIXUP);
    if (e->dtype == SB_F32)
        k_fixup<float><<<blocks, 256, 0, ctx->stream>>>(e->row_ptr.p, e->col_idx.p, static_cast<const float*>(bt),
                                                        perm_dev, e->n, e->m, flag_ij, flag_p, count, nullptr, 0, 0,
                                                        cneg, cpos, packed);
    else
        k_fixup<double><<<blocks, 256, 0, ctx->stream>>>(e->row_ptr.p, e->col_idx.p, static_cast<const double*>(bt),
                                                         perm_dev, e->n, e->m, flag_ij, flag_p, count, nullptr, 0, 0,
                                                         cneg, cpos, packed);
    SB_LAUNCH_CHECK(ctx);
}

void fixup_flag_buckets(sb_enrich* e, cudaStream_t st, const int32_t* perm_dev, const uint64_t* flag_ij,
                        const uint32_t* flag_p, const unsigned int* count_dev, int n_buckets, unsigned int cap,
                        uint32_t* cneg, uint32_t* cpos, uint32_t* packed) {
    sb_ctx* ctx = e->ctx;
    const void* bt = e->b_t;
    SB_CHECK(bt, "internal error: transposed attribute matrix not built");
    const unsigned grid = static_cast<unsigned>(ctx->num_sms * 8);  // the whole device on one bucket at a time
    KernelTimer kt(ctx, SB_K_FIXUP, st);
    if (e->dtype == SB_F32)
        k_fixup<float><<<grid, 256, 0, st>>>(e->row_ptr.p, e->col_idx.p, static_cast<const float*>(bt), perm_dev, e->n,
                                             e->m, flag_ij, flag_p, 0, count_dev, n_buckets, cap, cneg, cpos, packed);
    else
        k_fixup<double><<<grid, 256, 0, st>>>(e->row_ptr.p, e->col_idx.p, static_cast<const double*>(bt), perm_dev,
                                              e->n, e->m, flag_ij, flag_p, 0, count_dev, n_buckets, cap, cneg, cpos,
                                              packed);
    SB_LAUNCH_CHECK(ctx);
}

static void hypergeom_dev(sb_enrich* e, double* pv_dev, double* nes_dev) {
    sb_ctx* ctx = e->ctx;
    const int64_t n = e->n, m = e->m;
    // X = A @ nan0(B): exact integers.  Annotation matrices are binary, i.e. exactly representable in the tensor-core
    // plan's fixed point: X then comes from one pass of the digit GEMM (TCK_STORE); anything else takes the fp64 kernel.
    const int64_t* xfix = nullptr;
    const int32_t *shift = nullptr, *row_of_node = nullptr;
    int64_t mpad = 0;
    const bool fixed = tc_observed_exact(e, &xfix, &shift, &row_of_node, &mpad);
    const double* X = fixed ? nullptr : enrich_observed(e, SB_SCORE_SUM);
    DevBuf<int32_t> has;
    DevBuf<double> colsum, nneigh, lf;
    DevBuf<unsigned long long> total;
    has.reserve(n);
    colsum.reserve(m);
    nneigh.reserve(n);
    total.reserve(1);
    cudaStream_t st = ctx->stream;
    const unsigned wblocks = static_cast<unsigned>(sb_ceil_div(n * 32, 256));
    SB_CUDA(cudaMemsetAsync(colsum.p, 0, m * sizeof(double), st));
    SB_CUDA(cudaMemsetAsync(total.p, 0, sizeof(unsigned long long), st));
    dim3 cgrid(static_cast<unsigned>(sb_ceil_div(m, 128)),
               static_cast<unsigned>(std::max<int64_t>(1, std::min<int64_t>(64, n / 256))));
    if (e->dtype == SB_F32) {
        k_has_data<float><<<wblocks, 256, 0, st>>>(static_cast<const float*>(e->b), n, m, has.p);
        SB_LAUNCH_CHECK(ctx);
        k_colsum<float><<<cgrid, 128, 0, st>>>(static_cast<const float*>(e->b), n, m, colsum.p);
    } else {
        k_has_data<double><<<wblocks, 256, 0, st>>>(static_cast<const double*>(e->b), n, m, has.p);
        SB_LAUNCH_CHECK(ctx);
        k_colsum<double><<<cgrid, 128, 0, st>>>(static_cast<const double*>(e->b), n, m, colsum.p);
    }
    SB_LAUNCH_CHECK(ctx);
    k_sum_i32<<<64, 256, 0, st>>>(has.p, n, total.p);
    SB_LAUNCH_CHECK(ctx);
    k_nneigh<<<wblocks, 256, 0, st>>>(e->row_ptr.p, e->col_idx.p, has.p, n, nneigh.p);
    SB_LAUNCH_CHECK(ctx);
    unsigned long long n_total = 0;
    SB_CUDA(cudaMemcpyAsync(&n_total, total.p, sizeof n_total, cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaStreamSynchronize(st));
    // log-factorial table lf[k] = lgamma(k + 1), k = 0..n_total
    std::vector<double> h_lf(n_total + 2);
    for (unsigned long long k = 0; k < h_lf.size(); ++k) h_lf[k] = std::lgamma(static_cast<double>(k) + 1.0);
    lf.reserve(h_lf.size());
    SB_CUDA(cudaMemcpyAsync(lf.p, h_lf.data(), h_lf.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    const unsigned blocks = static_cast<unsigned>(std::min<int64_t>(sb_ceil_div(n * m, 256), ctx->num_sms * 16));
    {
        KernelTimer kt(ctx, SB_K_HYPERGEOM);
        if (fixed)
            k_hypergeom<true><<<blocks, 256, 0, st>>>(nullptr, xfix, shift, row_of_node, mpad, colsum.p, nneigh.p, lf.p,
                                                      static_cast<double>(n_total), n, m, pv_dev, nes_dev);
        else
            k_hypergeom<false><<<blocks, 256, 0, st>>>(X, nullptr, nullptr, nullptr, 0, colsum.p, nneigh.p, lf.p,
                                                       static_cast<double>(n_total), n, m, pv_dev, nes_dev);
        SB_LAUNCH_CHECK(ctx);
    }
    SB_CUDA(cudaStreamSynchronize(st));
}

}  // namespace sb

using namespace sb;

static size_t elem_size(int dtype) { return dtype == SB_F32 ? 4 : 8; }

static sb_enrich* enrich_new(sb_ctx* ctx, sb_neigh* a, int dtype, int64_t n, int64_t m) {
    SB_CHECK(ctx && a, "sb_enrich_create: NULL handle");
    SB_CHECK(a->ctx == ctx, "sb_enrich_create: neighborhoods belong to another context");
    SB_CHECK(dtype == SB_F32 || dtype == SB_F64, "sb_enrich_create: dtype must be SB_F32 or SB_F64");
    SB_CHECK(n == a->n, "sb_enrich_create: attribute rows (%lld) != nodes (%lld)", (long long)n, (long long)a->n);
    SB_CHECK(m > 0 && m < (1ll << 31), "sb_enrich_create: m=%lld out of range", (long long)m);
    sb_enrich* e = new sb_enrich;
    e->ctx = ctx;
    e->a = a;
    e->n = n;
    e->m = m;
    e->dtype = dtype;
    return e;
}

// engine selection: 'sum' -> digit GEMM with the fused comparison; 'z-score' -> digit GEMM (three sums per permutation)
// + fp64 comparison kernel when the plan can serve it (64+ attributes, finite values), else the exact SIMT engine
static void perm_counts_dispatch(sb_enrich* e, int score_type, int engine, const int32_t* perm_dev, int64_t num_perm,
                                 uint32_t* cneg, uint32_t* cpos, uint32_t* packed) {
    if (engine == SB_ENGINE_SIMT) {
        simt_perm_counts(e, score_type, perm_dev, num_perm, cneg, cpos, packed);
    } else if (score_type == SB_SCORE_SUM) {
        tc_perm_counts(e, perm_dev, num_perm, cneg, cpos, packed);
    } else if (!tc_perm_counts_z(e, perm_dev, num_perm, cneg, cpos, packed)) {
        SB_CHECK(engine != SB_ENGINE_TC,
                 "the tensor-core z-score null needs 64+ attributes, finite values and neighborhoods below 65536 nodes");
        simt_perm_counts(e, score_type, perm_dev, num_perm, cneg, cpos, packed);
    }
}

extern "C" {

int sb_enrich_create(sb_ctx* ctx, sb_neigh* a, const void* b_host, int dtype, int64_t n, int64_t m,
                     sb_enrich** out) {
    SB_API_BEGIN
    SB_CHECK(out && b_host, "sb_enrich_create: NULL argument");
    sb_enrich* e = enrich_new(ctx, a, dtype, n, m);
    try {
        ctx->bind();
        const size_t bytes = static_cast<size_t>(n) * m * elem_size(dtype);
        void* d = dev_alloc(bytes);
        e->b = d;
        e->b_owned = true;
        copy_in(ctx, d, b_host, bytes);
        build_csr(e);
        SB_CUDA(cudaStreamSynchronize(ctx->stream));
    } catch (...) {
        sb_enrich_destroy(e);
        throw;
    }
    *out = e;
    SB_API_END
}

int sb_enrich_create_dev(sb_ctx* ctx, sb_neigh* a, const void* b_dev, int dtype, int64_t n, int64_t m,
                         sb_enrich** out) {
    SB_API_BEGIN
    SB_CHECK(out && b_dev, "sb_enrich_create_dev: NULL argument");
    sb_enrich* e = enrich_new(ctx, a, dtype, n, m);
    try {
        ctx->bind();
        e->b = b_dev;
        e->b_owned = false;
        build_csr(e);
        SB_CUDA(cudaStreamSynchronize(ctx->stream));
    } catch (...) {
        sb_enrich_destroy(e);
        throw;
    }
    *out = e;
    SB_API_END
}

int sb_enrich_destroy(sb_enrich* e) {
    SB_API_BEGIN
    if (e) {
        e->ctx->bind();
        PhaseTrace tr(e->ctx, "enrich.destroy");
        if (e->tc) tc_plan_destroy(e->tc);
        if (e->b_owned && e->b) dev_free(const_cast<void*>(e->b));
        if (e->b_t) dev_free(e->b_t);
        delete e;
    }
    SB_API_END
}

int sb_enrich_score_dev(sb_enrich* e, int score_type, double* out_dev) {
    SB_API_BEGIN
    SB_CHECK(e && out_dev, "sb_enrich_score_dev: NULL argument");
    e->ctx->bind();
    enrich_score_into(e, score_type, out_dev);
    SB_API_END
}

int sb_enrich_observed_rows_dev(sb_enrich* e, int score_type, int64_t row0, int64_t row1, double** scores_dev) {
    SB_API_BEGIN
    SB_CHECK(e && scores_dev, "sb_enrich_observed_rows_dev: NULL argument");
    SB_CHECK(score_type == SB_SCORE_SUM || score_type == SB_SCORE_ZSCORE, "unknown neighborhood_score_type %d",
             score_type);
    e->ctx->bind();
    DevBuf<double>& buf = score_type == SB_SCORE_SUM ? e->s0_sum : e->s0_z;
    buf.reserve(static_cast<size_t>(e->n) * e->m);
    enrich_score_rows(e, score_type, buf.p, row0, row1);
    *scores_dev = buf.p;
    SB_API_END
}

int sb_enrich_observed_set_ready(sb_enrich* e, int score_type) {
    SB_API_BEGIN
    SB_CHECK(e, "sb_enrich_observed_set_ready: NULL handle");
    SB_CHECK(score_type == SB_SCORE_SUM || score_type == SB_SCORE_ZSCORE, "unknown neighborhood_score_type %d",
             score_type);
    DevBuf<double>& buf = score_type == SB_SCORE_SUM ? e->s0_sum : e->s0_z;
    SB_CHECK(buf.p && buf.n >= static_cast<size_t>(e->n) * e->m,
             "sb_enrich_observed_set_ready: call sb_enrich_observed_rows_dev first");
    (score_type == SB_SCORE_SUM ? e->have_s0_sum : e->have_s0_z) = true;
    SB_API_END
}

int sb_enrich_score(sb_enrich* e, int score_type, double* out_host) {
    SB_API_BEGIN
    SB_CHECK(e && out_host, "sb_enrich_score: NULL argument");
    e->ctx->bind();
    const double* s = enrich_observed(e, score_type);
    copy_out(e->ctx, out_host, s, static_cast<size_t>(e->n) * e->m * sizeof(double));
    SB_API_END
}

int sb_enrich_perm_counts_dev(sb_enrich* e, int score_type, int engine, const int32_t* perm_rows_dev,
                              int64_t num_perm, uint32_t* counts_neg_dev, uint32_t* counts_pos_dev) {
    SB_API_BEGIN
    SB_CHECK(e && perm_rows_dev && counts_neg_dev && counts_pos_dev, "sb_enrich_perm_counts_dev: NULL argument");
    SB_CHECK(num_perm >= 0, "sb_enrich_perm_counts_dev: num_perm < 0");
    SB_CHECK(score_type == SB_SCORE_SUM || score_type == SB_SCORE_ZSCORE, "unknown neighborhood_score_type %d",
             score_type);
    SB_CHECK(engine == SB_ENGINE_AUTO || engine == SB_ENGINE_SIMT || engine == SB_ENGINE_TC, "unknown engine %d",
             engine);
    e->ctx->bind();
    if (num_perm == 0) return 0;
    perm_counts_dispatch(e, score_type, engine, perm_rows_dev, num_perm, counts_neg_dev, counts_pos_dev, nullptr);
    SB_API_END
}

int sb_enrich_perm_counts_packed_dev(sb_enrich* e, int score_type, int engine, const int32_t* perm_rows_dev,
                                     int64_t num_perm, uint32_t* counts_packed_dev) {
    SB_API_BEGIN
    SB_CHECK(e && perm_rows_dev && counts_packed_dev, "sb_enrich_perm_counts_packed_dev: NULL argument");
    SB_CHECK(num_perm >= 0 && num_perm < 65536,
             "sb_enrich_perm_counts_packed_dev: num_perm must be in [0, 65536) (16-bit fields)");
    SB_CHECK(score_type == SB_SCORE_SUM || score_type == SB_SCORE_ZSCORE, "unknown neighborhood_score_type %d",
             score_type);
    SB_CHECK(engine == SB_ENGINE_AUTO || engine == SB_ENGINE_SIMT || engine == SB_ENGINE_TC, "unknown engine %d",
             engine);
    e->ctx->bind();
    if (num_perm == 0) return 0;
    perm_counts_dispatch(e, score_type, engine, perm_rows_dev, num_perm, nullptr, nullptr, counts_packed_dev);
    SB_API_END
}

int sb_counts_unpack_dev(sb_ctx* ctx, const uint32_t* counts_packed_dev, int64_t cells, uint32_t* counts_neg_dev,
                         uint32_t* counts_pos_dev) {
    SB_API_BEGIN
    SB_CHECK(ctx && counts_packed_dev && counts_neg_dev && counts_pos_dev && cells >= 0,
             "sb_counts_unpack_dev: bad argument");
    ctx->bind();
    if (cells == 0) return 0;
    const unsigned blocks = static_cast<unsigned>(std::min<int64_t>(sb_ceil_div(cells, 256), ctx->num_sms * 16));
    k_unpack_store<<<blocks, 256, 0, ctx->stream>>>(counts_packed_dev, cells, counts_neg_dev, counts_pos_dev);
    SB_LAUNCH_CHECK(ctx);
    SB_API_END
}

int sb_enrich_perm_counts(sb_enrich* e, int score_type, int engine, const int32_t* perm_rows_host, int64_t num_perm,
                          uint32_t* counts_neg_host, uint32_t* counts_pos_host) {
    SB_API_BEGIN
    SB_CHECK(e && perm_rows_host && counts_neg_host && counts_pos_host, "sb_enrich_perm_counts: NULL argument");
    SB_CHECK(num_perm >= 0, "sb_enrich_perm_counts: num_perm < 0");
    sb_ctx* ctx = e->ctx;
    ctx->bind();
    const size_t cells = static_cast<size_t>(e->n) * e->m;
    DevBuf<uint32_t> cnt;
    cnt.reserve(2 * cells);
    SB_CUDA(cudaMemsetAsync(cnt.p, 0, 2 * cells * sizeof(uint32_t), ctx->stream));
    // stream permutation indices in chunks of <= 1 GiB
    const int64_t chunk = std::max<int64_t>(1, (256ll << 20) / e->n);
    DevBuf<int32_t> perm;
    perm.reserve(static_cast<size_t>(std::min(chunk, std::max<int64_t>(num_perm, 1))) * e->n);
    for (int64_t p0 = 0; p0 < num_perm; p0 += chunk) {
        const int64_t np = std::min(chunk, num_perm - p0);
        SB_CUDA(cudaMemcpyAsync(perm.p, perm_rows_host + p0 * e->n, static_cast<size_t>(np) * e->n * sizeof(int32_t),
                                cudaMemcpyHostToDevice, ctx->stream));
        int rc = sb_enrich_perm_counts_dev(e, score_type, engine, perm.p, np, cnt.p, cnt.p + cells);
        if (rc) fail("%s", sb_last_error());
        SB_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    copy_out(ctx, counts_neg_host, cnt.p, cells * sizeof(uint32_t));
    copy_out(ctx, counts_pos_host, cnt.p + cells, cells * sizeof(uint32_t));
    SB_API_END
}

int sb_enrich_set_node_order(sb_enrich* e, const int32_t* order_host) {
    SB_API_BEGIN
    SB_CHECK(e, "sb_enrich_set_node_order: NULL handle");
    sb_ctx* ctx = e->ctx;
    ctx->bind();
    if (e->tc) {  // operands were built for another order
        SB_CUDA(cudaStreamSynchronize(ctx->stream));
        tc_plan_destroy(e->tc);
        e->tc = nullptr;
    }
    if (!order_host) {
        e->have_order = false;
        return 0;
    }
    const int64_t n = e->n;
    std::vector<int32_t> inv(n, -1);
    for (int64_t i = 0; i < n; ++i) {
        const int32_t v = order_host[i];
        SB_CHECK(v >= 0 && v < n && inv[v] < 0, "sb_enrich_set_node_order: order is not a permutation of 0..n-1 (entry %lld)",
                 (long long)i);
        inv[v] = static_cast<int32_t>(i);
    }
    e->order.reserve(n);
    e->order_inv.reserve(n);
    SB_CUDA(cudaMemcpyAsync(e->order.p, order_host, n * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    SB_CUDA(cudaMemcpyAsync(e->order_inv.p, inv.data(), n * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    SB_CUDA(cudaStreamSynchronize(ctx->stream));
    e->have_order = true;
    SB_API_END
}

int sb_enrich_stats(sb_enrich* e, int64_t* out7_host) {
    SB_API_BEGIN
    SB_CHECK(e && out7_host, "sb_enrich_stats: NULL argument");
    for (int i = 0; i < 7; ++i) out7_host[i] = e->stats[i];
    SB_API_END
}

int sb_enrich_hypergeom_dev(sb_enrich* e, double* pvalues_dev, double* nes_dev) {
    SB_API_BEGIN
    SB_CHECK(e, "sb_enrich_hypergeom_dev: NULL handle");
    e->ctx->bind();
    hypergeom_dev(e, pvalues_dev, nes_dev);
    SB_API_END
}

int sb_enrich_hypergeom(sb_enrich* e, double* pvalues_host, double* nes_host) {
    SB_API_BEGIN
    SB_CHECK(e, "sb_enrich_hypergeom: NULL handle");
    sb_ctx* ctx = e->ctx;
    ctx->bind();
    const size_t cells = static_cast<size_t>(e->n) * e->m;
    DevBuf<double> pv, nes;
    if (pvalues_host) pv.reserve(cells);
    if (nes_host) nes.reserve(cells);
    hypergeom_dev(e, pv.p, nes.p);
    if (pvalues_host) copy_out(ctx, pvalues_host, pv.p, cells * sizeof(double));
    if (nes_host) copy_out(ctx, nes_host, nes.p, cells * sizeof(double));
    SB_CUDA(cudaStreamSynchronize(ctx->stream));
    SB_API_END
}

}  // extern "C"
