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
    sb::DevBuf<int32_t> null_perm;   // staging for one piece of permutation indices
    int null_score = -1;             // -1: no null open
    int null_engine = 0;
    int64_t null_perms = 0;
    int64_t null_stats[7] = {0, 0, 0, 0, 0, 0, 0};
};

namespace sb {

// enrich.cu
void enrich_score_into(sb_enrich* e, int score_type, double* out_dev);
const double* enrich_observed(sb_enrich* e, int score_type);
void simt_perm_counts(sb_enrich* e, int score_type, const int32_t* perm_dev, int64_t num_perm, uint32_t* cneg,
                      uint32_t* cpos);
// exact fp64 re-evaluation of flagged (i, j, p) comparisons; entries are (i << 32 | j) , p pairs
void fixup_flags(sb_enrich* e, const int32_t* perm_dev, const uint64_t* flag_ij, const uint32_t* flag_p,
                 unsigned int count, uint32_t* cneg, uint32_t* cpos);

// gemm_tc.cu
void tc_plan_destroy(TcPlan* p);
void tc_perm_counts(sb_enrich* e, const int32_t* perm_dev, int64_t num_perm, uint32_t* cneg, uint32_t* cpos);

}  // namespace sb
