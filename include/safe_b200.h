/*
 * safe_b200.h -- C ABI of libsafe_b200.so: SAFE's two data-parallel stages on one B200 (sm_100a).
 *
 * The reference (baryshnikova-lab/safepy) has no FFI; its boundary for this path is the Python method
 * surface of class SAFE plus two free functions.  Each entry point below names the reference code it
 * replaces (paths are relative to the reference checkout):
 *
 *   stage 1  SAFE.define_neighborhoods            safepy/safe.py:369-430
 *   stage 2  compute_neighborhood_score           safepy/safe_extras.py:6-33
 *            run_permutations                     safepy/safe_extras.py:36-70
 *            SAFE.compute_pvalues_by_hypergeom    safepy/safe.py:556-608
 *
 * Conventions
 *   - plain C: pointers + sizes only; no Python / torch / C++ types cross the boundary.
 *   - every function returns 0 on success, non-zero on failure; sb_last_error() then describes the
 *     failure (thread-local, valid until the next call on that thread).
 *   - "_host" pointers are ordinary host memory owned by the caller; "_dev" pointers are device memory on
 *     the context's device (e.g. torch tensor data_ptr()).  The library owns whatever it allocates behind the
 *     opaque handles and frees it in the matching *_destroy.
 *   - one context = one CUDA device = one caller thread at a time.  Multi-GPU = one process (and one context)
 *     per device; the only exchanges (all-gather of packed rows, all-reduce of counts) are done by the host
 *     program over NCCL on the *_dev buffers.
 *   - there is NO CPU fallback: without a usable sm_100 device sb_ctx_create fails.
 *
 * Packed neighborhood matrix: row-major uint32 words, element (s,t) is bit (t & 31) of word [s*ld + (t >> 5)],
 * ld = sb_neigh_ld(n) words per row (multiple of 4 -> 16-byte aligned rows), padding bits are zero.
 */
#ifndef SAFE_B200_H
#define SAFE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SB_ABI_VERSION 1

/* dtype codes for attribute matrices */
#define SB_F32 0
#define SB_F64 1

/* neighborhood_score_type (safepy/safe_extras.py:6) */
#define SB_SCORE_SUM 0
#define SB_SCORE_ZSCORE 1

/* engine selection for the permutation null */
#define SB_ENGINE_AUTO 0   /* tcgen05 int8 digit GEMM + exact fix-up ('z-score' with fewer than 64 attributes: SIMT fp64) */
#define SB_ENGINE_SIMT 1   /* fp64 CUDA-core sparse kernel (exact by construction; validation / z-score) */
#define SB_ENGINE_TC 2     /* force the tensor-core path (z-score: needs 64+ attributes and finite values) */

typedef struct sb_ctx sb_ctx;       /* one device + stream + workspaces */
typedef struct sb_neigh sb_neigh;   /* bit-packed N x N neighborhood matrix on the device */
typedef struct sb_enrich sb_enrich; /* stage-2 plan: neighborhoods x attribute matrix, prepared operands */
typedef struct sb_perm_stream sb_perm_stream; /* host replay of run_permutations' RNG stream (no device involved) */

/* ------------------------------------------------------------------ library / context */
int sb_abi_version(void);
const char* sb_last_error(void);

/* device < 0: use the current CUDA device. Fails unless the device is compute capability 10.x. */
int sb_ctx_create(int device, sb_ctx** out);
int sb_ctx_destroy(sb_ctx* ctx);
/* ordinal of the calling thread's current CUDA device (what device < 0 resolves to), -1 without a usable device */
int sb_current_device(void);
/* CUDA device ordinal of the context (-1 for a NULL handle) */
int sb_ctx_device(sb_ctx* ctx);
/* run all subsequent work of this context on an existing cudaStream_t (e.g. torch's current stream) */
int sb_ctx_set_stream(sb_ctx* ctx, void* cuda_stream);
int sb_ctx_synchronize(sb_ctx* ctx);
/* Device memory is taken from a cached, stream-ordered pool (freed blocks are kept for reuse).  This returns every
 * cached block and the context's scratch buffers to the driver. */
int sb_ctx_release_memory(sb_ctx* ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
int64_t sb_ctx_launch_count(sb_ctx* ctx);
/* Device-time accounting: with profiling on, launches are bracketed by CUDA events on the context's stream.
 * sb_ctx_kernel_ms synchronizes, returns the summed milliseconds and the number of brackets of one kernel class
 * (0 score GEMM, 1 operand gather, 2 fp64 fix-up, 3 SSSP, 4 euclid, 5 hypergeom, 6 fp64 scores, 7 operand prep,
 * 8 p-value/NES tail, 9 row-wise FDR, 10 Jaccard)
 * and clears that class. */
int sb_ctx_profile(sb_ctx* ctx, int enable);
int sb_ctx_kernel_ms(sb_ctx* ctx, int kernel_class, double* ms_out, int64_t* count_out);
/* pin / unpin a caller-owned host buffer so the *_host entry points copy at full PCIe speed */
int sb_host_register(void* ptr, int64_t bytes);
int sb_host_unregister(void* ptr);

/* ------------------------------------------------------------------ stage 1: neighborhoods
 * replaces SAFE.define_neighborhoods, safepy/safe.py:369-430 */

int64_t sb_neigh_ld(int64_t n); /* words per packed row */

/* empty (all-zero) N x N packed matrix owned by the library */
int sb_neigh_create(sb_ctx* ctx, int64_t n, sb_neigh** out);
/* wrap a caller-owned device buffer of n*sb_neigh_ld(n) words (not freed by destroy) */
int sb_neigh_wrap_dev(sb_ctx* ctx, int64_t n, uint32_t* words_dev, sb_neigh** out);
int sb_neigh_destroy(sb_neigh* a);
int64_t sb_neigh_n(const sb_neigh* a);
uint32_t* sb_neigh_words_dev(sb_neigh* a);

/* Shortest-path neighborhoods (safe.py:403-415; networkx all_pairs_dijkstra_path_length semantics):
 * rows row0..row1-1 get A[s,t] = 1 iff dist(s,t) <= cutoff, distances are fp64 left-to-right path sums,
 * the source itself is always a member.  CSR is the symmetric adjacency (both directions present).
 * length_host == NULL means every edge has length 1 (metric 'shortpath' without a 'weight' attribute). */
int sb_neigh_shortpath(sb_neigh* a, const int64_t* indptr_host, const int32_t* indices_host,
                       const double* length_host, double cutoff, int64_t row0, int64_t row1);

/* Euclidean neighborhoods (safe.py:389-399; scipy pdist semantics): A[i,j] = 1 iff sqrt(dx*dx+dy*dy) < nr,
 * evaluated without FMA contraction.  Rows row0..row1-1 are written. */
int sb_neigh_euclid(sb_neigh* a, const double* x_host, const double* y_host, double nr, int64_t row0, int64_t row1);

/* upload an already packed matrix (rows row0..row1-1, ld words each) */
int sb_neigh_upload_packed(sb_neigh* a, const uint32_t* words_host, int64_t row0, int64_t row1);
/* download */
int sb_neigh_download_packed(sb_neigh* a, uint32_t* words_host, int64_t row0, int64_t row1);
/* neighbors per row (np.sum(neighborhoods, axis=1), safe.py:423) */
int sb_neigh_rowsums(sb_neigh* a, int64_t* out_host);
/* dense rows r0..r1-1 as 0/1: elem_bytes 1 -> uint8, 8 -> int64 (safe.py:387 dtype) */
int sb_neigh_unpack_rows(sb_neigh* a, int64_t r0, int64_t r1, int elem_bytes, void* out_host);

/* ------------------------------------------------------------------ stage 2: enrichment */

/* Prepare neighborhoods x attributes: uploads B (n x m row-major, NaN = no data), builds the CSR view of A,
 * and (lazily, on first permutation call) the tensor-core operands. */
int sb_enrich_create(sb_ctx* ctx, sb_neigh* a, const void* b_host, int dtype, int64_t n, int64_t m, sb_enrich** out);
/* same, B already on the device */
int sb_enrich_create_dev(sb_ctx* ctx, sb_neigh* a, const void* b_dev, int dtype, int64_t n, int64_t m,
                         sb_enrich** out);
int sb_enrich_destroy(sb_enrich* e);

/* Optional locality hint for the tensor-core null: order[i] = node placed at internal position i (a permutation of
 * 0..n-1, e.g. nodes sorted along a space-filling curve of the layout).  The null skips all-zero 256 x 64 tiles of
 * the neighborhood matrix, so an order in which neighborhoods are contiguous cuts its work by an order of
 * magnitude; results are identical for any order (inputs and outputs stay in the caller's node numbering).
 * NULL restores the identity.  Call before sb_enrich_perm_counts*. */
int sb_enrich_set_node_order(sb_enrich* e, const int32_t* order_host);

/* compute_neighborhood_score(A, B, type), safe_extras.py:6-33 -> fp64 [n x m] */
int sb_enrich_score(sb_enrich* e, int score_type, double* out_host);
int sb_enrich_score_dev(sb_enrich* e, int score_type, double* out_dev);
/* Observed scores with the node rows sharded over several GPUs: compute rows [row0, row1) into the plan's own [n x m]
 * score array (returned), let the host program exchange the row blocks in place (all-gather / broadcasts), then declare
 * the array complete; sb_enrich_null_finalize and sb_enrich_score then use it instead of recomputing every row. */
int sb_enrich_observed_rows_dev(sb_enrich* e, int score_type, int64_t row0, int64_t row1, double** scores_dev);
int sb_enrich_observed_set_ready(sb_enrich* e, int score_type);

/* run_permutations core, safe_extras.py:56-66.
 * perm_rows[p*n + t] = row of B that sits at node t during permutation p (the caller composes the reference's
 * cumulative in-place shuffles into gather indices).  counts_neg[i*m+j] += #{p : S_p[i,j] <= S_0[i,j]},
 * counts_pos likewise with >=.  The _dev variant ACCUMULATES into caller-zeroed device buffers so that
 * permutation shards can be streamed; the _host variant overwrites. */
int sb_enrich_perm_counts(sb_enrich* e, int score_type, int engine, const int32_t* perm_rows_host, int64_t num_perm,
                          uint32_t* counts_neg_host, uint32_t* counts_pos_host);
int sb_enrich_perm_counts_dev(sb_enrich* e, int score_type, int engine, const int32_t* perm_rows_dev,
                              int64_t num_perm, uint32_t* counts_neg_dev, uint32_t* counts_pos_dev);
/* The same counts ACCUMULATED into ONE word per cell, (counts_pos << 16) | counts_neg -- the array a permutation-
 * sharded run sums over the ranks with a single all-reduce (the reduction safe.py:518-519 does with np.sum over its
 * worker results; half the bytes of the two separate arrays).  The caller zeroes the buffer and keeps the number of
 * permutations summed into a word (over all calls and ranks) below 65536.  sb_counts_unpack_dev STORES the two
 * halves into separate arrays ([cells] each), e.g. after the all-reduce. */
int sb_enrich_perm_counts_packed_dev(sb_enrich* e, int score_type, int engine, const int32_t* perm_rows_dev,
                                     int64_t num_perm, uint32_t* counts_packed_dev);
int sb_counts_unpack_dev(sb_ctx* ctx, const uint32_t* counts_packed_dev, int64_t cells, uint32_t* counts_neg_dev,
                         uint32_t* counts_pos_dev);

/* statistics of the last perm_counts call on this plan:
 * [0] comparisons decided by the GEMM, [1] comparisons sent to the exact fix-up, [2] A tiles stored,
 * [3] A tiles possible, [4] digits used, [5] MMA k-tile iterations issued, [6] flag-list overflow batches */
int sb_enrich_stats(sb_enrich* e, int64_t* out7_host);

/* SAFE.compute_pvalues_by_hypergeom core, safe.py:573-608: p = hypergeom.sf(X-1, n_total, K_j, n_i), nes = -log10 p.
 * Either output may be NULL. */
int sb_enrich_hypergeom(sb_enrich* e, double* pvalues_host, double* nes_host);
int sb_enrich_hypergeom_dev(sb_enrich* e, double* pvalues_dev, double* nes_dev);

/* What SAFE.compute_pvalues inspects before choosing a test (safe.py:453-458): the number of NaNs in every attribute
 * column ([m]) and the number of values that are neither 0, 1 nor NaN (0 => 'auto' picks the hypergeometric test). */
int sb_enrich_attr_summary(sb_enrich* e, int64_t* nan_per_column_host, int64_t* other_values_out);

/* ------------------------------------------------------------------ stage 2: streaming null + fused tail
 * Same arithmetic as sb_enrich_perm_counts, but the two count arrays stay on the device between calls: the caller feeds
 * permutation indices in pieces while it is still replaying the reference's sequential RNG stream
 * (safe_extras.py:46-58), and sb_enrich_null_finalize turns the counts into everything compute_pvalues leaves on the
 * SAFE object without the counts visiting the host. */
int sb_enrich_null_begin(sb_enrich* e, int score_type, int engine);
/* perm_rows_host: [num_perm][n] gather rows as for sb_enrich_perm_counts; may be called any number of times */
int sb_enrich_null_add(sb_enrich* e, const int32_t* perm_rows_host, int64_t num_perm);
/* permutations counted so far and (optionally) the raw counts; any pointer may be NULL */
int sb_enrich_null_counts(sb_enrich* e, int64_t* num_perm_out, uint32_t* counts_neg_host, uint32_t* counts_pos_host);

/* Permutation shards on several GPUs (one process per GPU): the device addresses of the two count arrays of the open
 * null ([n x m] uint32 each, neg immediately followed by pos), so that the host program can sum them across ranks in
 * place (one NCCL all-reduce over 2 * n * m words), and the number of permutations they hold afterwards.  The caller
 * orders the streams: sb_ctx_synchronize before the collective, its own stream before the next library call. */
int sb_enrich_null_counts_dev(sb_enrich* e, uint32_t** counts_neg_dev, uint32_t** counts_pos_dev);
/* The cheaper exchange while the total number of permutations stays below 65536: the open null keeps the permutations
 * it has counted since its last fold in ONE packed word per cell, (counts_pos << 16) | counts_neg ([n x m] uint32, half
 * the bytes of the two arrays).  Returns that array and how many permutations it holds; after summing it across the
 * ranks in place (one all-reduce over n * m words), sb_enrich_null_set_perms(total) declares what it holds now. */
int sb_enrich_null_packed_dev(sb_enrich* e, uint32_t** counts_packed_dev, int64_t* perms_in_packed);
int sb_enrich_null_set_perms(sb_enrich* e, int64_t num_perm);

/* run_permutations' index stream (safe_extras.py:46-58) replayed natively on the host: np.random.seed(seed) of
 * NumPy's legacy MT19937 generator, then per permutation np.random.permutation(rows_with_data) applied in place to
 * the already permuted matrix; what comes out are the composed gather rows sb_enrich_perm_counts takes.
 * has_seed == 0 mirrors np.random.seed(None) (OS entropy).  A stream is pure host state. */
int sb_perm_stream_create(int64_t n, const int64_t* rows_with_data_host, int64_t n_with_data, int has_seed,
                          uint32_t seed, sb_perm_stream** out);
int sb_perm_stream_destroy(sb_perm_stream* s);
/* the next num_perm permutations as gather rows [num_perm][n]; rows_out_host == NULL skips them (a rank that owns a
 * later shard still has to draw the earlier ones: the stream cannot jump) */
int sb_perm_stream_next(sb_perm_stream* s, int64_t num_perm, int32_t* rows_out_host);
/* generator state (624 key words + position, as np.random.get_state() reports it) and permutations drawn so far, so
 * that the caller can leave NumPy's global generator where the reference would have left it */
int sb_perm_stream_state(sb_perm_stream* s, uint32_t* key624_out, int32_t* pos_out, int64_t* drawn_out);
/* sb_enrich_null_add for the next num_perm permutations of a stream, in one call: a producer thread replays piece
 * k + 1 into pinned memory while the device counts piece k */
int sb_enrich_null_add_stream(sb_enrich* e, sb_perm_stream* s, int64_t num_perm);
/* The same for rank `rank` of `world` processes that each hold the same stream: the next num_perm permutations are
 * dealt round-robin in equal pieces, this rank counts its pieces and draws-and-drops the others (the RNG cannot jump),
 * so the foreign draws overlap with its own device work.  The ranks' counts add up to the full null
 * (sb_enrich_null_counts_dev + one all-reduce, then sb_enrich_null_set_perms). */
int sb_enrich_null_add_stream_shard(sb_enrich* e, sb_perm_stream* s, int64_t num_perm, int world, int rank);
/* Optional read-ahead for the call above (the reference draws its permutations inside the loop, safe_extras.py:56-58;
 * here the draws may start before the plan exists): a background thread starts drawing rank `rank`'s share of the next
 * num_perm permutations at once -- while the caller uploads the attribute matrix, say -- and the following
 * sb_enrich_null_add_stream[_shard] with the same (num_perm, world, rank) consumes the rows as they become ready.
 * (sb_perm_stream_next with the same num_perm does too, for world == 1).  Shares above 2 GB of gather rows are not
 * read ahead (the call is then a no-op). */
int sb_perm_stream_prefetch(sb_perm_stream* s, int64_t num_perm, int world, int rank);

/* Tail of SAFE.compute_pvalues_by_randomization (safe.py:526-554) and of SAFE.compute_pvalues (safe.py:466-472):
 *   p = counts / P (NaN where the observed score is NaN); optional Benjamini-Hochberg adjustment of every row across
 *   attributes (multiple_testing, safe.py:536-542); NES+- = -log10(p == 0 ? zero_pvalue_floor : p);
 *   nes = NES+ (attribute_sign 0 'highest'), NES- (1 'lowest') or NES+ - NES- (2 'both');
 *   nes_binary = |nes| > nes_threshold; num_enriched[j] = sum_i nes_binary[i][j].
 * The caller supplies pvalue_of_count[c] = c / P and nes_of_count[c] = -log10(c == 0 ? 1/P : c / P) for c = 0..P
 * (table_len = P + 1) computed with its own libm, so that without FDR every output carries exactly the bits the
 * reference's NumPy expressions produce.  Outputs are [n x m] fp64 (num_enriched: [m]); any of them may be NULL. */
int sb_enrich_null_finalize(sb_enrich* e, const double* pvalue_of_count_host, const double* nes_of_count_host,
                            int64_t table_len, int multiple_testing, double zero_pvalue_floor, int attribute_sign,
                            double nes_threshold, double* ns_host, double* pvalues_neg_host, double* pvalues_pos_host,
                            double* nes_host, double* nes_binary_host, double* num_enriched_host);

/* sb_enrich_hypergeom followed by the optional row-wise FDR (safe.py:599-605), nes = -log10 p and the same
 * nes_binary / num_enriched tail (safe.py:466-472).  Any output may be NULL. */
int sb_enrich_hypergeom_finalize(sb_enrich* e, int multiple_testing, double nes_threshold, double* pvalues_host,
                                 double* nes_host, double* nes_binary_host, double* num_enriched_host);

/* Benjamini-Hochberg adjustment of every row of pvalues [n x m] across its m entries -- what
 * np.apply_along_axis(statsmodels.stats.multitest.fdrcorrection, 1, p)[:, 1, :] yields in safe.py:536-542 / 599-605
 * (method 'indep': sorted p / ((k + 1) / m), reverse running minimum, capped at 1; a NaN makes its whole row NaN). */
int sb_fdr_rows(sb_ctx* ctx, int64_t n, int64_t m, const double* pvalues_host, double* adjusted_host);

/* ------------------------------------------------------------------ graph-side helpers (next rows) */

/* safe_io.calculate_edge_lengths, safepy/safe_io.py:311-333: length[e] = sqrt(dx*dx + dy*dy) * weight[e], every
 * operation rounded separately (pdist + np.multiply); weight_host == NULL means weight 1.  A zero weight yields NaN:
 * the reference sets no 'length' attribute for such an edge (Dijkstra then charges its default cost 1). */
int sb_graph_edge_lengths(sb_ctx* ctx, int64_t n, const double* x_host, const double* y_host, int64_t n_edges,
                          const int32_t* eu_host, const int32_t* ev_host, const double* weight_host,
                          double* length_out_host);

/* Symmetric CSR of an undirected edge list (both directions stored, a self loop once, columns ascending inside a
 * row) -- the adjacency networkx's all_pairs_dijkstra_path_length walks in safe.py:406-410.
 * indices_out / value_out need room for 2 * n_edges entries; *nnz_out receives the number used. */
int sb_graph_csr(sb_ctx* ctx, int64_t n, int64_t n_edges, const int32_t* eu_host, const int32_t* ev_host,
                 const double* value_host, int64_t* indptr_out_host, int32_t* indices_out_host,
                 double* value_out_host, int64_t* nnz_out);

/* SAFE.define_top_attributes' connectivity test, safepy/safe.py:632-658: for every candidate attribute cand[k], the
 * connected components of the subgraph induced by the nodes with member[v * m + cand[k]] != 0.
 * labels_out (optional) [n_cand][n]: smallest node id of the node's component, -1 for non-members;
 * num_cc_out[k] = number of components, num_large_out[k] = components with at least min_size nodes. */
int sb_graph_components(sb_ctx* ctx, int64_t n, const int64_t* indptr_host, const int32_t* indices_host,
                        const uint8_t* member_host, int64_t m, const int32_t* cand_host, int64_t n_cand,
                        int32_t min_size, int32_t* labels_out_host, int32_t* num_cc_out_host,
                        int32_t* num_large_out_host);

/* The metric evaluation inside SAFE.define_domains' linkage(nes_binary[:, top].T, 'average', metric='jaccard'),
 * safepy/safe.py:672-675 (scipy pdist 'jaccard' on 0/1 rows): for the columns cols[0..n_cols) of member [n x m]
 * (non-zero = enriched), d(a, b) = |a xor b| / |a or b| (0 when both are empty), in pdist's condensed order
 * (n_cols * (n_cols - 1) / 2 doubles). */
int sb_attr_jaccard(sb_ctx* ctx, int64_t n, const uint8_t* member_host, int64_t m, const int32_t* cols_host,
                    int64_t n_cols, double* condensed_out_host);

/* ------------------------------------------------------------------ self-test hook (tests only)
 * Runs one 128 x N x K int8 tcgen05 GEMM from host operands through the production tile layouts and returns the
 * int32 accumulators; used to pin descriptor / layout encodings against a CPU product. */
int sb_selftest_mma_i8(sb_ctx* ctx, int ncols, int ktiles, int variant /*0 = production descriptors*/,
                       const int8_t* a_host /*128 x 64*ktiles, 0/1*/, const int8_t* b_host /*64*ktiles x ncols*/,
                       int32_t* d_host /*128 x ncols*/);

/* Streaming-rate probe of the same kernel: `grid` CTAs x `slots` accumulations over `ktiles` L2-resident tile pairs;
 * returns device milliseconds (bench / profiling only). */
int sb_selftest_mma_rate(sb_ctx* ctx, int ncols, int ktiles, int slots, int grid, int dbg /*1: no MMAs, 2: no copies*/,
                         const uint32_t* desc_override /*NULL, or {b_lbo, b_sbo, b_kstep} for speed-only experiments*/,
                         double* ms_out);

#ifdef __cplusplus
}
#endif
#endif /* SAFE_B200_H */
