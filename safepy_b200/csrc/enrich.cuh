// Stage-2 plan shared between the SIMT kernels (enrich.cu) and the tensor-core path (gemm_tc.cu).
#pragma once
#include "common.cuh"

namespace sb {
struct TcPlan;  // gemm_tc.cu
}

struct sb_enrich {
    sb_ctx* ctx = nullptr;
    sb_neigh* a = nullptr;
    int64_t n = 0, m = 0;
    int dtype = SB_F32;
    const void* b = nullptr;  // [n x m] row-major on the device, NaN = no data
    bool b_owned = false;
    void* b_t = nullptr;      // [m x n] transposed copy for the fix-up kernel, built on first use

    // CSR view of the packed matrix (ascending column order inside a row)
    sb::DevBuf<int64_t> row_ptr;   // n + 1
    sb::DevBuf<int32_t> col_idx;   // nnz
    int64_t nnz = 0;

    // observed 'sum' score in fp64 (ascending-t accumulation), built on demand
    sb::DevBuf<double> s0_sum;
    bool have_s0_sum = false;
    sb::DevBuf<double> s0_z;
    bool have_s0_z = false;

    // optional internal node order for the tensor-core path: order[i] = caller's node at internal position i
    sb::DevBuf<int32_t> order, order_inv;
    bool have_order = false;

    sb::TcPlan* tc = nullptr;
    int64_t stats[7] = {0, 0, 0, 0, 0, 0, 0};

    // streaming null (sb_enrich_null_*, finalize.cu): the two count arrays stay on the device between calls
    sb::DevBuf<uint32_t> null_cnt;   // [2][n * m]: neg, pos
    sb::DevBuf<uint32_t> null_pk;    // [n * m]: pos << 16 | neg of the permutations not yet moved into null_cnt
    int64_t null_pk_perms = 0;       // permutations summed into null_pk (16-bit fields: kept below 60000)
    sb::DevBuf<int32_t> null_perm;   // staging for one piece of permutation indices
    int null_score = -1;             // -1: no null open
    int null_engine = 0;
    int64_t null_perms = 0;
    int64_t null_stats[7] = {0, 0, 0, 0, 0, 0, 0};
};

namespace sb {

#ifdef __CUDACC__
// np.power(B, 2) keeps B's dtype: float32 squares are rounded to float32 before the fp64 dot (safe_extras.py:24)
template <class T>
__device__ __forceinline__ double sq_like_numpy(T v);
template <>
__device__ __forceinline__ double sq_like_numpy<float>(float v) {
    return static_cast<double>(__fmul_rn(v, v));
}
template <>
__device__ __forceinline__ double sq_like_numpy<double>(double v) {
    return __dmul_rn(v, v);
}
// z-score of a neighborhood from the sum, the sum of squares and the number of its non-NaN values, safe_extras.py:19-31
// (one definition: the exact kernels and the tensor-core z-score path must round identically)
__device__ __forceinline__ double zscore_from_sums(double sum, double sq, int64_t cnt) {
    const double N = static_cast<double>(cnt);
    const double M = sum / N;
    const double EXX = sq / N;
    const double EEX = __dmul_rn(M, M);
    const double sd = sqrt(__dsub_rn(EXX, EEX));
    double z = M / sd;
    if (sd == 0.0 || cnt < 3) z = __longlong_as_double(0x7FF8000000000000ll);
    return z;
}
// One (node i, virtual column c) score; c = p * m + j selects permutation p (perm == nullptr: identity) and
// attribute j.  Accumulation is fp64 in ascending neighbor order -- the order the oracle uses.
template <class T, bool ZS>
__device__ __forceinline__ double score_one(const int64_t* __restrict__ row_ptr, const int32_t* __restrict__ col_idx,
                                            const T* __restrict__ b, const int32_t* __restrict__ perm, int64_t n,
                                            int64_t m, int64_t i, int64_t p, int64_t j) {
    const int64_t e0 = row_ptr[i], e1 = row_ptr[i + 1];
    const int32_t* pr = perm ? perm + p * n : nullptr;
    double sum = 0.0, sq = 0.0;
    int64_t cnt = 0;
    for (int64_t e = e0; e < e1; ++e) {
        const int32_t t = col_idx[e];
        const int64_t r = pr ? pr[t] : t;
        const T v = b[r * m + j];
        if (v == v) {
            sum += static_cast<double>(v);
            if (ZS) {
                sq += sq_like_numpy<T>(v);
                ++cnt;
            }
        }
    }
    if (!ZS) return sum;
    return zscore_from_sums(sum, sq, cnt);
}

#endif

// enrich.cu
void enrich_score_into(sb_enrich* e, int score_type, double* out_dev);
// rows [row0, row1) only (out_dev is the whole [n x m] array)
void enrich_score_rows(sb_enrich* e, int score_type, double* out_dev, int64_t row0, int64_t row1);
const double* enrich_observed(sb_enrich* e, int score_type);
// counts are ADDED to cneg / cpos, or (cneg == cpos == nullptr) to one packed word per cell, pos << 16 | neg
void simt_perm_counts(sb_enrich* e, int score_type, const int32_t* perm_dev, int64_t num_perm, uint32_t* cneg,
                      uint32_t* cpos, uint32_t* packed = nullptr);
// exact fp64 re-evaluation of flagged (i, j, p) comparisons; entries are (i << 32 | j) , p pairs
void fixup_flags(sb_enrich* e, const int32_t* perm_dev, const uint64_t* flag_ij, const uint32_t* flag_p,
                 unsigned int count, uint32_t* cneg, uint32_t* cpos, uint32_t* packed = nullptr);
// [m x n] transposed copy of the attribute matrix (built on the context's stream on first use)
const void* enrich_transposed(sb_enrich* e);
// The same for a list bucketed by column group ([n_buckets][cap] entries, bucket b holding count_dev[b] of them), in
// ONE launch on `st` with the counts read on the device; a bucket whose count exceeds `cap` overflowed and is skipped
// as a whole (the caller redoes it).
void fixup_flag_buckets(sb_enrich* e, cudaStream_t st, const int32_t* perm_dev, const uint64_t* flag_ij,
                        const uint32_t* flag_p, const unsigned int* count_dev, int n_buckets, unsigned int cap,
                        uint32_t* cneg, uint32_t* cpos, uint32_t* packed);

// gemm_tc.cu
void tc_plan_destroy(TcPlan* p);
// cneg += packed & 0xffff, cpos += packed >> 16, packed = 0
void unpack_add_counts(sb_ctx* ctx, uint32_t* packed, int64_t cells, uint32_t* cneg, uint32_t* cpos);
// finalize.cu: permutations of an open streaming null (sb_enrich_null_*), counted into the packed accumulator
void null_count_dev(sb_enrich* e, const int32_t* perm_dev, int64_t num_perm);
void null_flush(sb_enrich* e);
void tc_perm_counts(sb_enrich* e, const int32_t* perm_dev, int64_t num_perm, uint32_t* cneg, uint32_t* cpos,
                    uint32_t* packed = nullptr);
// z-score null on the tensor cores; false: the plan cannot serve (the caller takes the exact SIMT engine)
bool tc_perm_counts_z(sb_enrich* e, const int32_t* perm_dev, int64_t num_perm, uint32_t* cneg, uint32_t* cpos,
                      uint32_t* packed = nullptr);
bool tc_observed_exact(sb_enrich* e, const int64_t** s0fix, const int32_t** shift, const int32_t** row_of_node,
                       int64_t* mpad);

}  // namespace sb
