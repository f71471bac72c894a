// Stage 2 on the 5th-generation tensor cores: permutation null as a block-sparse int8 digit GEMM with the
// "null vs observed" comparison fused into the TMEM epilogue (permuted scores never reach HBM).
// Replaces the hot loop of run_permutations (reference safepy/safe_extras.py:56-66) for 'sum' scores.
//
// Arithmetic.  A is {0,1}.  Every attribute column j of nan0(B) is turned into a fixed-point integer
// q = rint(v * 2^s_j) and written as D balanced base-256 digits (int8).  S_fix = A @ q is then computed EXACTLY by
// tcgen05.mma.kind::i8 (int32 accumulators in TMEM, one accumulator per digit plane, recombined in the epilogue).
// If a column is exactly representable (binary / integer / dyadic data) S_fix comparisons ARE the reference's
// comparisons.  Otherwise |S_fix - 2^s * S_true| <= n_i / 2, so |S_fix(p) - S_fix(0)| > n_i decides the comparison
// rigorously and the rare remainder is appended to a list that enrich.cu re-evaluates in fp64.
//
// Data movement.  The kernel runs on CTA PAIRS (cluster of two SMs, tcgen05 cta_group::2, M = 256): a pair owns a
// block of 256 rows of A (128 per CTA) and multiplies it with a gathered-operand tile whose 64*D columns are split
// between the two CTAs' shared memories, so every byte of the gathered operand that leaves L2 feeds 256 rows.
// A is kept as bit tiles: for every non-empty 256 x 64 tile one 64-bit mask per row (stored as two words, split_mask).
// The masks ride along with the gathered tiles through the smem ring, expander warps turn them into int8 0 / 1 in
// TMEM with byte permutes (no table), and the MMA takes its A operand from TMEM.  Gathered tiles are stored in HBM in
// the tensor core's canonical no-swizzle MN-major core-matrix order, one half per CTA, so one 1-D bulk async copy
// (TMA engine, UBLKCP) lands a half tile MMA-ready.  The z-score null (neighborhood_score_type = 'z-score',
// safe_extras.py:19-31) runs through the same kernel with six digit planes per 32 attributes and its own epilogue
// (TCK_Z, see "z-score null" below).  (A variant in which the kernel gathered the rows itself with 16-byte LDGSTS was measured at
// 387 ms vs 226 ms for the C3 null -- 16 cache lines per instruction at ~2 cycles per L1TEX wavefront, on a shared
// memory pipe the expanders already kept half busy; see profiles/r2a_bench_c3_fused_ldgsts_gather.json.)
#include <algorithm>
#include <climits>
#include <cmath>
#include <vector>

#include "enrich.cuh"
#include "sm100_ptx.cuh"

namespace sb {

constexpr int TC_ROWS = 128;             // rows of A per CTA (= TMEM lanes)
constexpr int TC_PROWS = 2 * TC_ROWS;    // rows per CTA pair = rows of one work unit's row block
constexpr int TC_KT = 64;                // K extent of one tile (2 MMAs of K=32)
constexpr int TC_STAGES = 5;             // smem ring depth, TC_TPS tiles per stage
constexpr int TC_ASLOTS = 4;             // TMEM ring depth (expanded A tiles), TC_APS tiles per slot
constexpr int TC_APS = 2;                // k-tiles per A slot (half a fill: the MMA warp frees A slots twice per fill)
constexpr int TC_TPS = 4;                // k-tiles per pipeline fill (tile lists are padded to a multiple of it):
                                         // one fill = 8 MMAs = 768 tensor cycles at D = 3, which covers the
                                         // ~550 cycles of barrier / issue latency every role spends per fill
constexpr int TC_EPI_WARPS = 8;
constexpr int TC_EXP_WARPS = 4;          // one per TMEM lane quarter
// The SM's warp arbiter favours high warp ids: the latency-critical single-warp roles (copy issue, MMA issue)
// therefore sit above the epilogue and expander warps.
constexpr int TC_EXP_WARP0 = TC_EPI_WARPS;
constexpr int TC_PROD_WARP = TC_EPI_WARPS + TC_EXP_WARPS;
constexpr int TC_MMA_WARP = TC_PROD_WARP + 1;
constexpr int TC_THREADS = (TC_MMA_WARP + 1) * 32;
// TMEM columns: accumulator buffer b at [256 b, 256 b + 64 D); A slot s (TC_APS tiles x 16 columns) in the gaps
// [192, 256) and [448, 512)
__host__ __device__ constexpr uint32_t tc_acol(uint32_t s) { return (s >> 1) * 256u + 192u + (s & 1u) * 32u; }
static_assert(TC_APS * 16 == 32 && TC_ASLOTS == 4 && TC_TPS == 2 * TC_APS, "A slots must tile the TMEM gaps");
constexpr int TC_SCHED = 4;              // depth of the work-unit ring (scheduler -> all other roles of the pair)
constexpr int TC_KT_SMEM = 1024;         // k-tile ids of the current row block cached in smem (tail: global)

enum : int { TCM_COUNT = 1, TCM_FLAG = 2, TCM_STORE = 4, TCM_RAW = 8, TCM_Z = 16 };
// kernel flavours (compile-time, so the hot epilogue carries no mode tests)
enum : int { TCK_COUNT = 0, TCK_STORE = 1, TCK_RAW = 2, TCK_Z = 3 };

struct GemmParams {
    const ulonglong2* a_bits; // [n_tiles / TPS][2 (CTA rank)][TPS][128]: the 64 membership bits A[row][64 kt + k] of
                              // a row as two words, see split_mask
    const int32_t* tile_ptr;  // [n_rb + 1], every row block holds a multiple of TC_TPS tiles
    const int32_t* tile_kt;   // [n_tiles]
    const int8_t* bcat;       // [slot][kt][2 (CTA rank)][64 x 32*D], K rows in the expanders' order (tc_kpos)
    int32_t n_kt, n_rb, n_cg, q_total, q_chunks, q_per;  // n_rb: blocks of TC_PROWS rows
    int32_t band_rb, n_bands; // unit order: band of row blocks, then q chunk, then column group, then row block
    int32_t rb0;              // first row block of this launch (units cover row blocks [rb0, rb0 + n_rb))
    unsigned int* unit_counter;  // dynamic scheduler (zeroed before the launch)
    int32_t mode;
    int64_t n, m, mpad;
    int32_t log2_mpad, pps, batch_perms;
    int64_t* s0fix;           // [n_rb * 256][mpad] observed fixed-point scores (TCK_STORE: output)
    const int64_t* row_ptr;   // band_i = row_ptr[i+1] - row_ptr[i]
    const int32_t* node_of_row;  // internal row -> caller's node id (nullptr: identity)
    const uint8_t* inexact;   // [mpad]
    uint32_t* cpk;            // packed counts (pos << 16 | neg) per (node, attribute), one atomic per cell
    uint64_t* flag_ij;
    uint32_t* flag_p;
    unsigned int* flag_count;  // [n_cg]: the list is bucketed by column group (locality of the fix-up kernel)
    unsigned int flag_cap;     // capacity of one bucket
    int32_t* raw_out;         // TCM_RAW: [256][64*D]
    uint32_t b_lbo, b_sbo;
    int32_t q_wrap;           // 1: every slot re-reads slot 0 (rate self-test); otherwise unused
    int32_t dbg;              // rate probe only: bit0 = no MMAs, bit1 = no copies, bit2 = MMA warp does not wait for operands
    uint32_t b_kstep;         // descriptor start-address advance per K=32 MMA
    long long* prof;          // per-role cycle counters [16] (see print_prof), or nullptr
    // TCK_Z (z-score null): a column group is 32 attributes x 6 digit planes (3 of the value, 2 of the square, 1 of
    // the non-NaN indicator), so one accumulation holds the three sums of a cell (see the epilogue)
    const double* z0t;        // [32 n_cg][rows_pad] observed z-scores (exact engine), internal row order
    const float4* zcol;       // per attribute {2^-shift1, 2^-shift2, error radius per neighbor of the sum, of the squares}
    const int32_t *zshift1, *zshift2;
    const uint8_t *zinex1, *zinex2;
    int64_t rows_pad;
};

// cycle accounting per role, enabled by a non-null GemmParams::prof (SB_TRACE runs and the rate probe)
#define TC_TIMED(KINDV, acc, stmt)            \
    do {                                      \
        if (PROF) {                           \
            const long long _t0 = clock64();  \
            stmt;                             \
            (acc) += clock64() - _t0;         \
        } else {                              \
            stmt;                             \
        }                                     \
    } while (0)

struct UnitInfo {
    int32_t rb, cg, q0, q1, t0, nfills, pad0, pad1;
};

template <int D>
struct TcCfg {
    static constexpr int NCOLS = 64 * D;
    static constexpr int HALF_B = TC_KT * NCOLS / 2;                // this CTA's half of a gathered tile
    static constexpr int TPS = TC_TPS;
    static constexpr int STAGE_B = TPS * HALF_B;                    // gathered-operand half tiles of one fill
    static constexpr int STAGE_A = TPS * TC_ROWS * 16;              // this CTA's A bit tiles (one split mask per row)
    static constexpr int STAGE = STAGE_B + STAGE_A;
    static constexpr int STAGES = TC_STAGES;
    static constexpr int OFF_S0HI = STAGES * STAGE;                 // int32 [64][128]
    static constexpr int OFF_S0LO = OFF_S0HI + 64 * TC_ROWS * 4;    // uint32 [16][128], 4 columns per word
    static constexpr int OFF_KT = OFF_S0LO + 16 * TC_ROWS * 4;      // int32 [TC_KT_SMEM]
    static constexpr int OFF_UNIT = OFF_KT + TC_KT_SMEM * 4;        // UnitInfo [TC_SCHED]
    static constexpr int OFF_BAR = OFF_UNIT + TC_SCHED * 32;
    static constexpr int SMEM = OFF_BAR + 512;  // 34 mbarriers + the TMEM base address
};

// 64 membership bits -> 16 TMEM words of four int8 values 0 / 1, without a table.
// One byte permute (PRMT) expands four bits.  Its four selector nibbles are four CONSECUTIVE nibbles of a mask word,
// and the 8-byte source it selects from is constant: {00,01,00,01,00,01,00,01} returns 0 / 1 by bit 0 of a nibble,
// {00,00,01,01,00,00,01,01} by bit 1, {00,00,00,00,01,01,01,01} by bit 2.  Bit 3 of a selector nibble would ask for
// the replicated SIGN of the selected byte instead, so the plan stores a row's mask w as two words (split_mask): w with
// bit 3 of every nibble cleared, and those bits moved down to bit 0 of their nibble (read with the first source).
// Word 4 g + j of a row comes from 16-bit group g with source j: 16 permutes + 4 shifts per 64 bits (the lookup-table
// version of round 1 cost 24 instructions, 8 of them shared-memory loads with bank conflicts).
// Byte i of that word is mask bit 16 g + 4 i + j, i.e. the MMA sees the 64 entries of a tile in the order
//   K position 16 g + 4 j + i  <->  tile entry 16 g + 4 i + j        (tc_kpos, an involution),
// and the gathered operand tiles are written with their rows in the same order.
// Why 0 / 1 and not 0 / -1 (which a single mask word gives, the sign replication returning 0x00 / 0xFF for bit 3 as
// well): the chip runs at its power cap and the tensor pipe draws measurably less on 0x01 than on 0xFF -- the same
// pipeline, masks and instruction stream sustain 3050 int8 TOPS against 2770-2890 in the L2-hot rate probe; the C3 null
// itself gains ~1 % (profiles/r2h_gemm_ab_experiments.txt, 10).
__host__ __device__ constexpr int tc_kpos(int p) { return (p & ~15) | ((p & 3) << 2) | ((p >> 2) & 3); }

template <uint32_t LO, uint32_t HI>
__device__ __forceinline__ uint32_t expand4(uint32_t ctl) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(LO), "r"(HI), "r"(ctl));
    return d;
}
// The mask w of a row is stored as two words (split_mask): x = w with bit 3 of every nibble cleared, y = those bits
// moved to bit 0 of their nibble.
__host__ __device__ __forceinline__ ulonglong2 split_mask(uint64_t w) {
    return make_ulonglong2(w & 0x7777777777777777ull, (w >> 3) & 0x1111111111111111ull);
}
__device__ __forceinline__ void expand_bits(ulonglong2 m, uint32_t (&r)[16]) {
    const uint32_t xlo = static_cast<uint32_t>(m.x), xhi = static_cast<uint32_t>(m.x >> 32);
    const uint32_t ylo = static_cast<uint32_t>(m.y), yhi = static_cast<uint32_t>(m.y >> 32);
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        const uint32_t xw = g < 2 ? xlo : xhi, yw = g < 2 ? ylo : yhi;
        const uint32_t x = (g & 1) ? xw >> 16 : xw;  // PRMT reads the low 16 bits of its selector only
        const uint32_t y = (g & 1) ? yw >> 16 : yw;
        r[g * 4 + 0] = expand4<0x01000100u, 0x01000100u>(x);
        r[g * 4 + 1] = expand4<0x01010000u, 0x01010000u>(x);
        r[g * 4 + 2] = expand4<0x00000000u, 0x01010101u>(x);
        r[g * 4 + 3] = expand4<0x01000100u, 0x01000100u>(y);
    }
}

// z-score comparison of one (cell, permutation) in fp64 from the exact fixed-point sums -- what the fp32 filter of the
// TCK_Z epilogue could not settle (~1e-3 of the comparisons).  Columns whose values AND squares are exactly
// representable: the fixed-point sums are the exact sums and the z-score is evaluated by the very function the exact
// engine uses (same bits, ties included).  Otherwise |fixed-point sum - true sum| <= (non-NaN neighbors) / 2 units,
// which bounds the z-score from both sides (it increases with the sum and is monotone in the sum of squares with the
// sign of the mean); an interval that does not contain the observed z-score decides rigorously, the rest is left to the
// exact engine (*undecided).  Returns the packed count increment (pos << 16 | neg).
__device__ __noinline__ uint32_t z_compare_fp64(long long s1, long long s2, int cnt, double z0, int shift1,
                                                int shift2, bool inex1, bool inex2, bool* undecided) {
    const double sc1 = ldexp(1.0, -shift1), sc2 = ldexp(1.0, -shift2);  // fixed point -> value, exact
    const double a = static_cast<double>(s1) * sc1, b = static_cast<double>(s2) * sc2;
    const double N = static_cast<double>(cnt);
    *undecided = false;
    if (!inex1 && !inex2) {
        const double z = zscore_from_sums(a, b, cnt);
        return (z <= z0 ? 1u : 0u) | (z >= z0 ? 0x10000u : 0u);
    }
    // error radii of the two sums per non-NaN neighbor: half a unit, widened by 2^-12 for the exact engine's own fp64
    // accumulation error (< n * 2^-53 * 2^22 units, n < 65536); zero for exactly representable columns
    const double ra = inex1 ? 0.5 * sc1 * (1.0 + 0x1p-12) : 0.0, rb = inex2 ? 0.5 * sc2 * (1.0 + 0x1p-12) : 0.0;
    const double ea = ra * N, eb = rb * N;
    const double alo = a - ea, ahi = a + ea;
    const double mmax = fmax(fabs(alo), fabs(ahi)) / N;
    const double var_min = (b - eb) / N - mmax * mmax;
    // well inside the domain only (no cancellation in EXX - EEX): else the exact engine decides
    if (!(var_min > 0.0) || !((b + eb) / N + mmax * mmax < 1e6 * var_min)) {
        *undecided = true;
        return 0u;
    }
    const double mh = ahi / N, ml = alo / N;
    const double bh = ahi > 0.0 ? b - eb : b + eb;  // the z-score falls with the spread when the mean is > 0
    const double bl = alo > 0.0 ? b + eb : b - eb;
    double zhi = mh / sqrt(bh / N - mh * mh);
    double zlo = ml / sqrt(bl / N - ml * ml);
    zhi += 1e-12 * fabs(zhi) + 1e-300;
    zlo -= 1e-12 * fabs(zlo) + 1e-300;
    if (zlo > z0) return 0x10000u;
    if (zhi < z0) return 1u;
    *undecided = true;
    return 0u;
}

// Work units are numbered so that consecutive units share operands in L2: inside a band of row blocks (whose A
// tiles stay L2-resident) the q chunk is the slowest index, then the column group, then the row block -- CTA pairs
// that fetch neighbouring unit numbers read the same gathered operand slab (q chunk, column group).
__device__ __forceinline__ void decode_unit(const GemmParams& p, int u, int& rb, int& cg, int& q0, int& q1) {
    const int per_full = p.band_rb * p.n_cg * p.q_chunks;
    const int b = min(u / per_full, p.n_bands - 1);
    const int r = u - b * per_full;
    const int rows = min(p.band_rb, p.n_rb - b * p.band_rb);
    rb = b * p.band_rb + r % rows;
    const int rest = r / rows;
    cg = rest % p.n_cg;
    const int qc = rest / p.n_cg;
    q0 = qc * p.q_per;
    q1 = min(p.q_total, q0 + p.q_per);
}

// One cluster = one CTA pair (ranks 0 = leader, 1 = peer); 448 threads per CTA.
// Warp roles in BOTH CTAs: 0..7 = epilogue (own 128 rows), 8..11 = A expanders (own rows, own TMEM), 12 = bulk-copy
// producer (own half of the gathered tiles + own bit tiles), 13 = TMEM alloc; in the leader warp 12 is also the
// scheduler and warp 13 the MMA issuer for the pair.
// Scheduler: draws work units from a global counter and publishes them into the unit rings of both CTAs, so all
// roles of the pair walk the same unit sequence.  Issue loops are warp-uniform with one elected lane issuing.
// Cross-CTA signalling: peer expanders / epilogue warps / ring consumers arrive on the leader's afull / tempty /
// sempty barriers through shared::cluster; the leader's tcgen05.commit multicasts to the aempty / empty / tfull
// barriers of both CTAs.
// Epilogue warp w reads TMEM lane quarter (w & 3) and the 32-column half (w >> 2) of every digit plane.
template <int D, int KIND, bool SMALL_M, bool PROF>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1) k_gemm(const GemmParams p) {
    using C = TcCfg<D>;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sB = smem;
    int32_t* s_hi = reinterpret_cast<int32_t*>(smem + C::OFF_S0HI);
    uint32_t* s_lo = reinterpret_cast<uint32_t*>(smem + C::OFF_S0LO);
    int32_t* s_kt = reinterpret_cast<int32_t*>(smem + C::OFF_KT);
    UnitInfo* s_unit = reinterpret_cast<UnitInfo*>(smem + C::OFF_UNIT);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
    uint64_t* full = bars;                        // [STAGES]  local: this CTA's copies landed
    uint64_t* empty = bars + C::STAGES;           // [STAGES]  local: MMA commit (multicast) + own expanders
    uint64_t* tfull = bars + 2 * C::STAGES;       // [2]       local: MMA commit (multicast)
    uint64_t* tempty = bars + 2 * C::STAGES + 2;  // [2]       leader's is used: epilogue warps of both CTAs
    uint64_t* sfull = bars + 2 * C::STAGES + 4;   // [TC_SCHED] local: written by the leader's scheduler
    uint64_t* sempty = sfull + TC_SCHED;          // [TC_SCHED] leader's is used: ring consumers of both CTAs
    uint64_t* afull = sempty + TC_SCHED;          // [TC_ASLOTS] leader's is used: expanders of both CTAs
    uint64_t* aempty = afull + TC_ASLOTS;         // [TC_ASLOTS] local: MMA commit (multicast)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aempty + TC_ASLOTS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < C::STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1 + TC_EXP_WARPS);  // MMA commit + own expander warps (done reading the bit tiles)
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tfull[b], 1);
            mbar_init(&tempty[b], 2 * TC_EPI_WARPS);
        }
        for (int s = 0; s < TC_SCHED; ++s) {
            mbar_init(&sfull[s], 1);
            // consumers: leader MMA + peer producer + epilogue and expander warps of both CTAs
            mbar_init(&sempty[s], 2 * (TC_EPI_WARPS + TC_EXP_WARPS) + 2);
        }
        for (int s = 0; s < TC_ASLOTS; ++s) {
            mbar_init(&afull[s], 2 * TC_EXP_WARPS);
            mbar_init(&aempty[s], 1);
        }
        mbar_fence_init();
    }
    if (warp == TC_MMA_WARP) tmem_alloc2(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // barrier inits of both CTAs are visible before any remote arrive
    tc_fence_after();
    const uint32_t tbase = *tmem_slot;

    const int n_units = p.n_rb * p.n_cg * p.q_chunks;
    // leader-side addresses (shared::cluster) of the barriers the peer signals remotely
    const uint32_t ld_sempty = mapa_u32(smem_u32(sempty), 0);
    const uint32_t ld_afull = mapa_u32(smem_u32(afull), 0);
    const uint32_t ld_tempty = mapa_u32(smem_u32(tempty), 0);

    if (warp == TC_PROD_WARP) {
        // ------------------------------------------------------------ scheduler (leader) + producer (both CTAs)
        const size_t q_stride = static_cast<size_t>(p.n_cg) * p.n_kt * (2 * C::HALF_B);  // bytes between slots
        uint32_t stage = 0, phase = 0, uit = 0;
        long long pt_wait = 0, pt_fills = 0;
        const long long pt_start = PROF ? clock64() : 0;
        unsigned int next_u = 0;
        if (leader) {
            if (lane == 0) next_u = atomicAdd(p.unit_counter, 1u);
            next_u = __shfl_sync(0xffffffffu, next_u, 0);
        }
        const uint32_t peer_unit = mapa_u32(smem_u32(s_unit), 1), peer_sfull = mapa_u32(smem_u32(sfull), 1);
        while (true) {
            const uint32_t sl = uit % TC_SCHED, spar = (uit / TC_SCHED) & 1u;
            int rb, cg, q0, q1, t0, nfills;
            if (leader) {
                const int u = static_cast<int>(next_u);
                mbar_wait_cluster(&sempty[sl], spar ^ 1u);
                const bool stop = u >= n_units || u < 0;
                if (!stop) {
                    if (lane == 0) next_u = atomicAdd(p.unit_counter, 1u);  // consumed at the top of the next iteration
                    decode_unit(p, u, rb, cg, q0, q1);
                    rb += p.rb0;
                    t0 = p.tile_ptr[rb];
                    nfills = (p.tile_ptr[rb + 1] - t0) / C::TPS;
                } else {
                    rb = -1;
                    cg = q0 = q1 = t0 = nfills = 0;
                }
                if (lane == 0) {
                    UnitInfo inf;
                    inf.rb = rb; inf.cg = cg; inf.q0 = q0; inf.q1 = q1; inf.t0 = t0; inf.nfills = nfills;
                    inf.pad0 = inf.pad1 = 0;
                    s_unit[sl] = inf;
                    const uint32_t pu = peer_unit + sl * static_cast<uint32_t>(sizeof(UnitInfo));
                    st_cluster_u32(pu + 0, static_cast<uint32_t>(rb));
                    st_cluster_u32(pu + 4, static_cast<uint32_t>(cg));
                    st_cluster_u32(pu + 8, static_cast<uint32_t>(q0));
                    st_cluster_u32(pu + 12, static_cast<uint32_t>(q1));
                    st_cluster_u32(pu + 16, static_cast<uint32_t>(t0));
                    st_cluster_u32(pu + 20, static_cast<uint32_t>(nfills));
                    mbar_arrive(&sfull[sl]);
                    mbar_arrive_cluster(peer_sfull + sl * 8);
                }
                if (stop) break;
            } else {
                mbar_wait_cluster(&sfull[sl], spar);
                rb = s_unit[sl].rb; cg = s_unit[sl].cg; q0 = s_unit[sl].q0; q1 = s_unit[sl].q1;
                t0 = s_unit[sl].t0; nfills = s_unit[sl].nfills;
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(ld_sempty + sl * 8);
                if (rb < 0) break;
            }
            const int nk = nfills * C::TPS;
            for (int i = lane; i < min(nk, TC_KT_SMEM); i += 32) s_kt[i] = p.tile_kt[t0 + i];
            __syncwarp();
            // this CTA's half of every gathered tile and its own rows of the A bit tiles
            const int8_t* const b_cg = p.bcat + static_cast<size_t>(cg) * p.n_kt * (2 * C::HALF_B) + rank * C::HALF_B;
            const ulonglong2* const a_unit =
                p.a_bits + (static_cast<size_t>(t0 / C::TPS) * 2 + rank) * (C::TPS * TC_ROWS);
            for (int q = q0; q < q1 && !(p.dbg & 4); ++q) {
                const int8_t* const b_q = b_cg + (p.q_wrap == 1 ? 0 : static_cast<size_t>(q) * q_stride);
                for (int f = 0; f < nfills; ++f) {
                    int kts[C::TPS];
#pragma unroll
                    for (int t = 0; t < C::TPS; ++t) {
                        const int i = f * C::TPS + t;
                        kts[t] = i < TC_KT_SMEM ? s_kt[i] : p.tile_kt[t0 + i];
                    }
                    TC_TIMED(KIND, pt_wait, mbar_wait(&empty[stage], phase ^ 1u));
                    if (PROF) ++pt_fills;
                    if (elect_one()) {
                        if (p.dbg & 2) {
                            mbar_arrive(&full[stage]);
                        } else {
                            mbar_expect_tx(&full[stage], C::STAGE);
                            uint8_t* const st = sB + stage * C::STAGE;
#pragma unroll
                            for (int t = 0; t < C::TPS; ++t)
                                bulk_g2s(st + t * C::HALF_B, b_q + static_cast<size_t>(kts[t]) * (2 * C::HALF_B),
                                         C::HALF_B, &full[stage]);
                            bulk_g2s(st + C::STAGE_B, a_unit + static_cast<size_t>(f) * (2 * C::TPS * TC_ROWS),
                                     C::STAGE_A, &full[stage]);
                        }
                    }
                    __syncwarp();
                    if (++stage == C::STAGES) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
            if (leader) next_u = __shfl_sync(0xffffffffu, next_u, 0);
            ++uit;
        }
        if (PROF && leader && lane == 0) {
            atomicAdd(reinterpret_cast<unsigned long long*>(p.prof + 0), clock64() - pt_start);
            atomicAdd(reinterpret_cast<unsigned long long*>(p.prof + 1), pt_wait);
            atomicAdd(reinterpret_cast<unsigned long long*>(p.prof + 10), pt_fills);
            if (p.dbg & 4) p.prof[10] = static_cast<long long>(p.q_total) * (p.tile_ptr[1] / C::TPS);  // probe: n_rb = 1
        }
    } else if (warp >= TC_EXP_WARP0 && warp < TC_PROD_WARP) {
        // ------------------------------------------------------------ A expanders: bits -> int8 0/1 in TMEM
        const int quarter = warp - TC_EXP_WARP0;
        const int r = quarter * 32 + lane;
        const uint32_t t_lane = tbase + (static_cast<uint32_t>(quarter * 32) << 16);
        uint32_t aslot = 0, aphase = 0, stage = 0, phase = 0, uit = 0;
        long long xt_wait = 0, xt_st = 0, xt_full = 0, xt_alu = 0;
        const long long xt_start = PROF ? clock64() : 0;
        while (true) {
            const uint32_t sl = uit % TC_SCHED, spar = (uit / TC_SCHED) & 1u;
            mbar_wait_cluster(&sfull[sl], spar);
            const int rb = s_unit[sl].rb, q0 = s_unit[sl].q0, q1 = s_unit[sl].q1, nfills = s_unit[sl].nfills;
            __syncwarp();
            if (lane == 0) {
                if (leader)
                    mbar_arrive(&sempty[sl]);
                else
                    mbar_arrive_cluster(ld_sempty + sl * 8);
            }
            ++uit;
            if (rb < 0) break;
            // Software pipeline over half fills: the TMEM store of half h is in flight while half h+1 is expanded
            // (byte permutes), so neither the tcgen05.st latency nor the permutes sit on the critical path alone.
            auto publish = [&](uint32_t slot) {   // the store into `slot` was issued earlier
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (leader)
                        mbar_arrive(&afull[slot]);
                    else
                        mbar_arrive_cluster_relaxed(ld_afull + slot * 8);
                }
            };
            bool pending = false;
            uint32_t pending_slot = 0;
            for (int q = q0; q < q1 && !(p.dbg & 4); ++q) {
                for (int f = 0; f < nfills; ++f) {
                    // the bit tiles of this fill arrive in the smem stage together with the gathered operand
                    TC_TIMED(KIND, xt_full, mbar_wait(&full[stage], phase));
                    const ulonglong2* const sbits =
                        reinterpret_cast<const ulonglong2*>(sB + stage * C::STAGE + C::STAGE_B) + r;
                    ulonglong2 w[C::TPS];
#pragma unroll
                    for (int t = 0; t < C::TPS; ++t) w[t] = sbits[t * TC_ROWS];
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty[stage]);
                    if (++stage == C::STAGES) {
                        stage = 0;
                        phase ^= 1u;
                    }
#pragma unroll
                    for (int h = 0; h < C::TPS / TC_APS; ++h) {
                        uint32_t e[TC_APS][16];
                        const long long xa0 = PROF ? clock64() : 0;
                        if (p.dbg & 8) {  // timing experiment: no expansion work (results are garbage)
#pragma unroll
                            for (int t = 0; t < TC_APS; ++t)
#pragma unroll
                                for (int c = 0; c < 16; ++c) e[t][c] = static_cast<uint32_t>(w[h * TC_APS + t].x);
                        } else {
#pragma unroll
                            for (int t = 0; t < TC_APS; ++t) expand_bits(w[h * TC_APS + t], e[t]);
                        }
                        if (PROF) {  // keep the permutes inside the timed window
#pragma unroll
                            for (int t = 0; t < TC_APS; ++t)
#pragma unroll
                                for (int c = 0; c < 16; ++c) asm volatile("" ::"r"(e[t][c]));
                            xt_alu += clock64() - xa0;
                        }
                        if (pending) TC_TIMED(KIND, xt_st, publish(pending_slot));
                        TC_TIMED(KIND, xt_wait, mbar_wait(&aempty[aslot], aphase ^ 1u));
                        tc_fence_after();
                        const uint32_t ta = t_lane + tc_acol(aslot);
#pragma unroll
                        for (int t = 0; t < TC_APS; ++t) tmem_st16(ta + t * 16, e[t]);
                        pending = true;
                        pending_slot = aslot;
                        if (++aslot == TC_ASLOTS) {
                            aslot = 0;
                            aphase ^= 1u;
                        }
                    }
                }
            }
            if (pending) publish(pending_slot);  // nothing follows in this unit: do not hold the MMA warp back
        }
        if (PROF && leader && warp == TC_EXP_WARP0 && lane == 0) {
            atomicAdd(reinterpret_cast<unsigned long long*>(p.prof + 2), clock64() - xt_start);
            atomicAdd(reinterpret_cast<unsigned long long*>(p.prof + 3), xt_wait);
            atomicAdd(reinterpret_cast<unsigned long long*>(p.prof + 11), xt_st);
            atomicAdd(reinterpret_cast<unsigned long long*>(p.prof + 12), xt_full);
            atomicAdd(reinterpret_cast<unsigned long long*>(p.prof + 13), xt_alu);
        }
    } else if (warp == TC_MMA_WARP) {
        // ------------------------------------------------------------ MMA issuer (leader only; one elected lane)
        if (leader) {
            const uint32_t idesc = idesc_i8(TC_PROWS, C::NCOLS, /*a_signed*/ 1, /*b_signed*/ 1, /*a MN*/ 0, /*b MN*/ 1);
            const uint64_t b_desc0 = smem_desc_noswz(smem_u32(sB), p.b_lbo, p.b_sbo);
            const uint32_t b_ks = p.b_kstep >> 4;
            uint32_t stage = 0, aslot = 0, aphase = 0, acc_it = 0, uit = 0;
            long long mt_wa = 0, mt_wt = 0;
            const long long mt_start = PROF ? clock64() : 0;
            while (true) {
                const uint32_t sl = uit % TC_SCHED, spar = (uit / TC_SCHED) & 1u;
                mbar_wait(&sfull[sl], spar);
                const int rb = s_unit[sl].rb, q0 = s_unit[sl].q0, q1 = s_unit[sl].q1, nfills = s_unit[sl].nfills;
                __syncwarp();
                if (lane == 0) mbar_arrive(&sempty[sl]);
                if (rb < 0) break;
                for (int q = q0; q < q1; ++q) {
                    // accumulations issued so far: buffer = acc_it & 1, barrier phase = (acc_it >> 1) & 1
                    const uint32_t buf = acc_it & 1u;
                    TC_TIMED(KIND, mt_wt, mbar_wait(&tempty[buf], ((acc_it >> 1) & 1u) ^ 1u));
                    const uint32_t d_tmem = tbase + buf * 256;
                    for (int f = 0; f < nfills; ++f) {
                        // no wait on full[stage]: the expanders of both CTAs pass their full[stage] before they
                        // publish the first A slot of the fill
                        const uint64_t b_st = b_desc0 + stage * (C::STAGE >> 4);
#pragma unroll
                        for (int h = 0; h < C::TPS / TC_APS; ++h) {
                            if (!(p.dbg & 4)) TC_TIMED(KIND, mt_wa, mbar_wait(&afull[aslot], aphase));
                            tc_fence_after();
                            if (elect_one()) {
                                if (!(p.dbg & 1)) {
                                    // A: TMEM slot (same address in both CTAs), 8 columns per K=32 MMA.
                                    // B: MN-major smem half tiles, four 8-row K groups per MMA
                                    const uint32_t a_t = tbase + tc_acol(aslot);
#pragma unroll
                                    for (int t = 0; t < TC_APS; ++t) {
#pragma unroll
                                        for (int ks = 0; ks < TC_KT / 32; ++ks)
                                            mma_i8_ts2(d_tmem, a_t + t * 16 + ks * 8,
                                                       b_st + ((h * TC_APS + t) * (C::HALF_B >> 4) + ks * b_ks), idesc,
                                                       (f | h | t | ks) != 0);
                                    }
                                }
                                if (!(p.dbg & 4)) mma_commit2(&aempty[aslot]);
                                if (h == C::TPS / TC_APS - 1) {
                                    if (!(p.dbg & 4)) mma_commit2(&empty[stage]);
                                    if (f == nfills - 1) mma_commit2(&tfull[buf]);
                                }
                            }
                            __syncwarp();
                            if (++aslot == TC_ASLOTS) {
                                aslot = 0;
                                aphase ^= 1u;
                            }
                        }
                        if (++stage == C::STAGES) stage = 0;
                    }
                    ++acc_it;
                }
                ++uit;
            }
            if (PROF && lane == 0) {
                atomicAdd(reinterpret_cast<unsigned long long*>(p.prof + 4), clock64() - mt_start);
                atomicAdd(reinterpret_cast<unsigned long long*>(p.prof + 5), mt_wa);
                atomicAdd(reinterpret_cast<unsigned long long*>(p.prof + 7), mt_wt);
            }
        }
    } else {
        // ------------------------------------------------------------ epilogue: TMEM -> compare -> counts
        const int quarter = warp & 3;                       // TMEM lane quarter this warp may read
        const int half = warp >> 2;                         // which 32 of the slot's 64 columns
        const int row_in_tile = quarter * 32 + lane;
        const int c0 = half * 32;
        uint32_t acc_it = 0, uit = 0;
        long long et_wait = 0;
        const long long et_start = PROF ? clock64() : 0;
        while (true) {
            const uint32_t sl = uit % TC_SCHED, spar = (uit / TC_SCHED) & 1u;
            mbar_wait_cluster(&sfull[sl], spar);
            const int rb = s_unit[sl].rb, cg = s_unit[sl].cg, q0 = s_unit[sl].q0, q1 = s_unit[sl].q1;
            __syncwarp();
            if (lane == 0) {
                if (leader)
                    mbar_arrive(&sempty[sl]);
                else
                    mbar_arrive_cluster(ld_sempty + sl * 8);
            }
            ++uit;
            if (rb < 0) break;
            const int64_t row = static_cast<int64_t>(rb) * TC_PROWS + rank * TC_ROWS + row_in_tile;
            const bool row_ok = row < p.n;
            // counts, bands and fix-ups are addressed by the caller's node id
            const int64_t node = (row_ok && p.node_of_row) ? p.node_of_row[row] : row;
            const int64_t jbase = SMALL_M ? 0 : static_cast<int64_t>(cg) * 64;
            int band = 0, bh = 1;
            uint32_t inexact_mask = 0;
            uint32_t cnt[32];
            if (KIND == TCK_COUNT) {
                band = row_ok ? static_cast<int>(p.row_ptr[node + 1] - p.row_ptr[node]) : 0;
                bh = (band + 255) / 256 + 1;
                // thread-private copy of the observed fixed-point scores, split as S0 = hi * 256 + lo
#pragma unroll 2
                for (int g = 0; g < 8; ++g) {
                    uint32_t lo4 = 0;
#pragma unroll
                    for (int x = 0; x < 4; ++x) {
                        const int cc = g * 4 + x, c = c0 + cc;
                        const int64_t j = SMALL_M ? (c & (p.mpad - 1)) : jbase + c;
                        if (p.inexact[j]) inexact_mask |= 1u << cc;
                        const long long s0 = p.s0fix[row * p.mpad + j];
                        s_hi[c * TC_ROWS + row_in_tile] = static_cast<int>(s0 >> 8);
                        lo4 |= static_cast<uint32_t>(s0 & 255) << (8 * x);
                    }
                    s_lo[((c0 >> 2) + g) * TC_ROWS + row_in_tile] = lo4;
                }
#pragma unroll
                for (int cc = 0; cc < 32; ++cc) cnt[cc] = 0;
            }
            // TCK_Z: this thread owns 16 attributes of the group (its 32-column half of the three plane pairs holds
            // all six planes of them); their observed z-scores and column constants sit in a warp-private smem area
            float* const zw = reinterpret_cast<float*>(smem + C::OFF_S0HI) + warp * (16 * 32 + 16 * 4);
            const int64_t jz0 = static_cast<int64_t>(cg) * 32 + half * 16;
            if (KIND == TCK_Z) {
                __syncwarp();
#pragma unroll 4
                for (int aa = 0; aa < 16; ++aa) {
                    const int64_t j = jz0 + aa;
                    float z = __int_as_float(0x7fc00000);
                    if (row_ok && j < p.m) z = static_cast<float>(p.z0t[j * p.rows_pad + row]);
                    zw[aa * 32 + lane] = z;
                }
                if (lane < 16) {
                    float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (jz0 + lane < p.m) c = p.zcol[jz0 + lane];
                    reinterpret_cast<float4*>(zw + 16 * 32)[lane] = c;
                }
                __syncwarp();
#pragma unroll
                for (int cc = 0; cc < 16; ++cc) cnt[cc] = 0;
            }
            // one band per warp-uniform test: columns of a group are normally all exact or all inexact
            const bool all_exact = __all_sync(0xffffffffu, inexact_mask == 0);
            const bool all_inexact = __all_sync(0xffffffffu, inexact_mask == 0xffffffffu);

            for (int q = q0; q < q1; ++q) {
                const uint32_t buf = acc_it & 1u, tpar = (acc_it >> 1) & 1u;
                TC_TIMED(KIND, et_wait, mbar_wait(&tfull[buf], tpar));
                tc_fence_after();
                const uint32_t t_addr = tbase + (static_cast<uint32_t>(quarter * 32) << 16) + buf * 256 + c0;
                uint32_t flagmask = 0;
                constexpr int CW = 8;  // columns per TMEM load: keeps the live accumulator set at 8 D registers
                if (KIND == TCK_Z) {
                    // TMEM column d * 64 + half * 32 + s * 16 + a = plane 2 d + s of attribute half * 16 + a;
                    // planes 0-2: digits of the sum of the values, 3-4: of the sum of the squares, 5: non-NaN neighbors
                    const float4* const zc4 = reinterpret_cast<const float4*>(zw + 16 * 32);
#pragma unroll
                    for (int it = 0; it < 2; ++it) {
                        uint32_t acc[6][8];
#pragma unroll
                        for (int pl = 0; pl < 6; ++pl) tmem_ld8(t_addr + (pl >> 1) * 64 + (pl & 1) * 16 + it * 8, acc[pl]);
                        tmem_ld_wait();
#pragma unroll
                        for (int x = 0; x < 8; ++x) {
                            const int aa = it * 8 + x;
                            const float z0f = zw[aa * 32 + lane];
                            const int cntv = static_cast<int32_t>(acc[5][x]);
                            if (!(z0f == z0f) || cntv < 3) continue;  // NaN never compares (safe_extras.py:65-66)
                            const int v0 = static_cast<int32_t>(acc[0][x]);
                            const int hi1 = (v0 >> 8) + static_cast<int32_t>(acc[1][x]) +
                                            (static_cast<int32_t>(acc[2][x]) << 8);  // sum = hi1 * 256 + lo1
                            const int lo1 = v0 & 255;
                            const int s2 = static_cast<int32_t>(acc[3][x]) + (static_cast<int32_t>(acc[4][x]) << 8);
                            const float4 zc = zc4[aa];
                            {
                                // fp32 filter: the interval of z_compare_fp64 in single precision, accepted only well
                                // away from cancellation (spread < 64 x variance: every step is then good to ~1e-5
                                // relative) and with a 1e-4 relative margin.  All but ~1e-3 of the comparisons end here.
                                const float Nf = static_cast<float>(cntv), rN = 1.f / Nf;
                                const float af = fmaf(static_cast<float>(hi1), 256.f, static_cast<float>(lo1)) * zc.x;
                                const float bf = static_cast<float>(s2) * zc.y;
                                const float eaf = zc.z * Nf, ebf = zc.w * Nf;
                                const float mh = (af + eaf) * rN, ml = (af - eaf) * rN;
                                const float mmax = fmaxf(fabsf(mh), fabsf(ml));
                                const float vmin = (bf - ebf) * rN - mmax * mmax, spread = (bf + ebf) * rN + mmax * mmax;
                                if (vmin > 0.f && spread < 64.f * vmin) {
                                    const float bh = mh > 0.f ? bf - ebf : bf + ebf, bl = ml > 0.f ? bf + ebf : bf - ebf;
                                    const float zhi = mh * rsqrtf(bh * rN - mh * mh), zlo = ml * rsqrtf(bl * rN - ml * ml);
                                    const float tol = 1e-4f * (fabsf(zhi) + fabsf(zlo) + fabsf(z0f)) + 1e-30f;
                                    if (zlo - tol > z0f) {
                                        cnt[aa] += 0x10000u;
                                        continue;
                                    }
                                    if (zhi + tol < z0f) {
                                        cnt[aa] += 1u;
                                        continue;
                                    }
                                }
                            }
                            const int64_t j = jz0 + aa;
                            bool und;
                            cnt[aa] += z_compare_fp64(static_cast<long long>(hi1) * 256 + lo1, s2, cntv,
                                                      p.z0t[j * p.rows_pad + row], p.zshift1[j], p.zshift2[j],
                                                      p.zinex1[j] != 0, p.zinex2[j] != 0, &und);
                            if (und) flagmask |= 1u << aa;
                        }
                    }
                }
#pragma unroll
                for (int ch = 0; ch < ((KIND == TCK_Z || (p.dbg & 16)) ? 0 : 32 / CW); ++ch) {  // dbg 16: timing experiment, no epilogue work
                    uint32_t acc[D][CW];
#pragma unroll
                    for (int d = 0; d < D; ++d) tmem_ld8(t_addr + d * 64 + ch * CW, acc[d]);
                    tmem_ld_wait();
#pragma unroll
                    for (int x = 0; x < CW; ++x) {
                        const int cc = ch * CW + x;   // column inside this warp's half
                        const int c = c0 + cc;        // column inside the slot
                        if (KIND == TCK_RAW) {
                            if (q == q1 - 1) {  // rate probe: only the last accumulation of a unit is stored
#pragma unroll
                                for (int d = 0; d < D; ++d)
                                    p.raw_out[(rank * TC_ROWS + row_in_tile) * C::NCOLS + d * 64 + c] =
                                        static_cast<int32_t>(acc[d][x]);
                            }
                        } else if (KIND == TCK_STORE && SMALL_M) {
                            long long S = static_cast<int32_t>(acc[0][x]);
                            if (D > 1) S += static_cast<long long>(static_cast<int32_t>(acc[1][x])) << 8;
                            if (D > 2) S += static_cast<long long>(static_cast<int32_t>(acc[D - 1][x])) << 16;
                            if (c < p.mpad) p.s0fix[row * p.mpad + c] = S;
                        }
                    }
                    if (KIND == TCK_STORE && !SMALL_M) {
                        // observed scores: row-major [n_rb * 256][mpad], what the COUNT flavour preloads
#pragma unroll
                        for (int x = 0; x < CW; ++x) {
                            long long S = static_cast<int32_t>(acc[0][x]);
                            if (D > 1) S += static_cast<long long>(static_cast<int32_t>(acc[1][x])) << 8;
                            if (D > 2) S += static_cast<long long>(static_cast<int32_t>(acc[D - 1][x])) << 16;
                            const int64_t col = jbase + c0 + ch * CW + x;
                            p.s0fix[row * p.mpad + col] = S;
                        }
                    }
                    if (KIND == TCK_COUNT) {
                        // Fast pass on the high part only.  S = hi * 256 + lo with lo in [0, 256) (all int32, |S| < 2^38
                        // is checked on the host), so S - S0 = dh * 256 + dl with |dl| <= 255 and
                        //   dh >= bh  =>  S - S0 >  band,      dh <= -bh  =>  S - S0 < -band,     bh = (band + 255) / 256 + 1
                        // (band = 0 for exactly representable columns).  The few comparisons in between are redone
                        // with the low parts in the slow pass below.
                        uint32_t undmask = 0;
#pragma unroll
                        for (int x = 0; x < CW; ++x) {
                            const int cc = ch * CW + x;
                            const int c = c0 + cc;
                            const int a0 = static_cast<int32_t>(acc[0][x]);
                            int hi = a0 >> 8;
                            if (D > 1) hi += static_cast<int32_t>(acc[1][x]);
                            if (D > 2) hi += static_cast<int32_t>(acc[D - 1][x]) << 8;
                            const int dh = hi - s_hi[c * TC_ROWS + row_in_tile];
                            const int bhc = all_exact ? 1 : ((all_inexact || ((inexact_mask >> cc) & 1u)) ? bh : 1);
                            bool gt = dh >= bhc, lt = dh <= -bhc;
                            bool und = !(gt | lt);
                            if (SMALL_M) {
                                const bool live = q * p.pps + (c >> p.log2_mpad) < p.batch_perms;
                                gt &= live;
                                lt &= live;
                                und &= live;
                            }
                            cnt[cc] += (gt ? 0x10000u : 0u) + (lt ? 1u : 0u);
                            if (und) undmask |= 1u << x;
                        }
                        if (__any_sync(0xffffffffu, undmask != 0)) {
#pragma unroll
                            for (int x = 0; x < CW; ++x) {
                                if (!((undmask >> x) & 1u)) continue;
                                const int cc = ch * CW + x;
                                const int c = c0 + cc;
                                const int a0 = static_cast<int32_t>(acc[0][x]);
                                int hi = a0 >> 8;
                                if (D > 1) hi += static_cast<int32_t>(acc[1][x]);
                                if (D > 2) hi += static_cast<int32_t>(acc[D - 1][x]) << 8;
                                const int lo = a0 & 255;
                                const int o_hi = s_hi[c * TC_ROWS + row_in_tile];
                                const int o_lo = (s_lo[(c >> 2) * TC_ROWS + row_in_tile] >> (8 * (c & 3))) & 255;
                                const int diff = (hi - o_hi) * 256 + (lo - o_lo);  // |hi - o_hi| < bh here: no overflow
                                const int b = (!all_exact && (all_inexact || ((inexact_mask >> cc) & 1u))) ? band : 0;
                                const bool gt = diff > b, lt = diff < -b;
                                const bool tie = !(gt | lt) && b == 0;
                                cnt[cc] += ((gt | tie) ? 0x10000u : 0u) + ((lt | tie) ? 1u : 0u);
                                if (!(gt | lt) && b != 0) flagmask |= 1u << cc;
                            }
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (leader)
                        mbar_arrive(&tempty[buf]);
                    else
                        mbar_arrive_cluster_relaxed(ld_tempty + buf * 8);
                }
                ++acc_it;
                if (KIND == TCK_Z) {
                    if (!(p.mode & TCM_COUNT)) {
#pragma unroll
                        for (int cc = 0; cc < 16; ++cc) cnt[cc] = 0;
                    }
                    // rare (~1e-4): comparisons the fixed-point sums cannot decide go to the exact engine (k_zfix)
                    if (flagmask && (p.mode & TCM_FLAG)) {
                        while (flagmask) {
                            const int aa = __ffs(flagmask) - 1;
                            flagmask &= flagmask - 1;
                            const unsigned int k = atomicAdd(p.flag_count, 1u);
                            if (k < p.flag_cap) {
                                p.flag_ij[k] = (static_cast<uint64_t>(node) << 32) | static_cast<uint64_t>(jz0 + aa);
                                p.flag_p[k] = static_cast<uint32_t>(q);
                            }
                        }
                    }
                }
                if (KIND == TCK_COUNT) {
                    if (!(p.mode & TCM_COUNT)) {
#pragma unroll
                        for (int cc = 0; cc < 32; ++cc) cnt[cc] = 0;
                    }
                    // rare: comparisons the fixed-point scores cannot decide go to the exact fp64 fix-up list
                    if (flagmask && (p.mode & TCM_FLAG) && row_ok) {
                        while (flagmask) {
                            const int cc = __ffs(flagmask) - 1;
                            flagmask &= flagmask - 1;
                            const int c = c0 + cc;
                            const int64_t j = SMALL_M ? (c & (p.mpad - 1)) : jbase + c;
                            const int pl = SMALL_M ? q * p.pps + (c >> p.log2_mpad) : q;
                            if (j < p.m) {
                                const unsigned int k = atomicAdd(&p.flag_count[cg], 1u);
                                if (k < p.flag_cap) {
                                    const size_t at = static_cast<size_t>(cg) * p.flag_cap + k;
                                    p.flag_ij[at] = (static_cast<uint64_t>(node) << 32) | static_cast<uint64_t>(j);
                                    p.flag_p[at] = static_cast<uint32_t>(pl);
                                }
                            }
                        }
                    }
                }
            }
            if (KIND == TCK_Z && (p.mode & TCM_COUNT) && row_ok) {
#pragma unroll
                for (int aa = 0; aa < 16; ++aa)
                    if (jz0 + aa < p.m && cnt[aa]) atomicAdd(&p.cpk[node * p.m + jz0 + aa], cnt[aa]);
            }
            if (KIND == TCK_COUNT && (p.mode & TCM_COUNT) && row_ok) {
                if (SMALL_M) {
                    // columns c and c' with c % mpad == c' % mpad carry different permutations of the same attribute
                    for (int j = 0; j < p.mpad && j < p.m; ++j) {
                        uint32_t v = 0;
#pragma unroll
                        for (int cc = 0; cc < 32; ++cc)
                            if (((c0 + cc) & (p.mpad - 1)) == j) v += cnt[cc];
                        if (v) atomicAdd(&p.cpk[node * p.m + j], v);
                    }
                } else {
#pragma unroll
                    for (int cc = 0; cc < 32; ++cc) {
                        const int64_t j = jbase + c0 + cc;
                        if (j < p.m && cnt[cc]) atomicAdd(&p.cpk[node * p.m + j], cnt[cc]);
                    }
                }
            }
        }
        if (PROF && leader && warp == 0 && lane == 0) {
            atomicAdd(reinterpret_cast<unsigned long long*>(p.prof + 8), clock64() - et_start);
            atomicAdd(reinterpret_cast<unsigned long long*>(p.prof + 9), et_wait);
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // no remote arrive or multicast commit may target a CTA that has already exited
    if (warp == TC_MMA_WARP) tmem_dealloc2(tbase, 512);
}

// ------------------------------------------------------------------------------------------------ operand builders
// packed matrix in the internal node order: out[inv[s]][inv[t]] = A[s][t]; one warp per CSR row of A
__global__ void __launch_bounds__(256) k_permute_packed(const int64_t* __restrict__ row_ptr,
                                                        const int32_t* __restrict__ col_idx,
                                                        const int32_t* __restrict__ inv, int64_t n, int64_t ld,
                                                        uint32_t* __restrict__ out) {
    const int64_t row = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    uint32_t* const dst = out + static_cast<int64_t>(inv[row]) * ld;
    for (int64_t e = row_ptr[row] + lane; e < row_ptr[row + 1]; e += 32) {
        const int32_t t = inv[col_idx[e]];
        atomicOr(&dst[t >> 5], 1u << (t & 31));
    }
}


// occupancy of 256 x 64 tiles of the packed matrix; one block per row block (= rows of one CTA pair)
__global__ void __launch_bounds__(256) k_tile_occ(const uint32_t* __restrict__ words, int64_t n, int64_t ld,
                                                  int32_t n_kt, uint8_t* __restrict__ occ,
                                                  int32_t* __restrict__ rb_count) {
    const int rb = blockIdx.x;
    const int64_t r0 = static_cast<int64_t>(rb) * TC_PROWS, r1 = min(n, r0 + TC_PROWS);
    __shared__ int s_total;
    if (threadIdx.x == 0) s_total = 0;
    __syncthreads();
    int mine = 0;
    for (int kt = threadIdx.x; kt < n_kt; kt += blockDim.x) {
        uint32_t any = 0;
        for (int64_t r = r0; r < r1; ++r) {
            const uint2 w = *reinterpret_cast<const uint2*>(words + r * ld + 2 * kt);
            any |= w.x | w.y;
        }
        occ[static_cast<size_t>(rb) * n_kt + kt] = any ? 1 : 0;
        mine += any ? 1 : 0;
    }
    if (mine) atomicAdd(&s_total, mine);
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_total == 0) {  // keep at least one (all-zero) tile so every row block still runs its epilogue
            occ[static_cast<size_t>(rb) * n_kt] = 1;
            s_total = 1;
        }
        rb_count[rb] = s_total;
    }
}

// number of distinct k-tiles that the row blocks of every band touch (one block per band)
__global__ void __launch_bounds__(256) k_band_kt(const uint8_t* __restrict__ occ, int32_t n_rb, int32_t n_kt,
                                                 int32_t band_rb, int32_t* __restrict__ band_cnt) {
    const int b = blockIdx.x;
    const int r0 = b * band_rb, r1 = min(n_rb, r0 + band_rb);
    __shared__ int s_total;
    if (threadIdx.x == 0) s_total = 0;
    __syncthreads();
    int mine = 0;
    for (int kt = threadIdx.x; kt < n_kt; kt += blockDim.x) {
        int any = 0;
        for (int r = r0; r < r1 && !any; ++r) any = occ[static_cast<size_t>(r) * n_kt + kt];
        mine += any;
    }
    if (mine) atomicAdd(&s_total, mine);
    __syncthreads();
    if (threadIdx.x == 0) band_cnt[b] = s_total;
}

// tile_kt / tile_rb lists from the occupancy flags; one block per row block (serial chunks + block scan).
// Row blocks are padded to a multiple of TC_TPS tiles: padding entries get tile_rb = -1 (expanded as all-zero A
// tiles, so they contribute nothing) and repeat the last valid k-tile id.
__global__ void __launch_bounds__(256) k_tile_list(const uint8_t* __restrict__ occ, int32_t n_kt,
                                                   const int32_t* __restrict__ tile_ptr, int32_t* __restrict__ tile_kt,
                                                   int32_t* __restrict__ tile_rb) {
    const int rb = blockIdx.x;
    __shared__ int part[256];
    __shared__ int s_last;
    const int chunk = (n_kt + 255) / 256;
    const int b = threadIdx.x * chunk, e = min(n_kt, b + chunk);
    int c = 0;
    for (int kt = b; kt < e; ++kt) c += occ[static_cast<size_t>(rb) * n_kt + kt];
    part[threadIdx.x] = c;
    if (threadIdx.x == 0) s_last = 0;
    __syncthreads();
    for (int o = 1; o < 256; o <<= 1) {
        int v = threadIdx.x >= o ? part[threadIdx.x - o] : 0;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    const int t0 = tile_ptr[rb];
    int at = t0 + (threadIdx.x ? part[threadIdx.x - 1] : 0);
    int last = -1;
    for (int kt = b; kt < e; ++kt)
        if (occ[static_cast<size_t>(rb) * n_kt + kt]) {
            tile_kt[at] = kt;
            tile_rb[at] = rb;
            ++at;
            last = kt;
        }
    if (last >= 0) atomicMax(&s_last, last);
    __syncthreads();
    const int real = part[255], padded = tile_ptr[rb + 1] - t0;
    for (int i = real + threadIdx.x; i < padded; i += blockDim.x) {
        tile_kt[t0 + i] = s_last;
        tile_rb[t0 + i] = -1;
    }
}

// the 256 x 64 bit block of every stored tile as one split mask per row (padding tiles and rows are zero), laid
// out so that the 128 rows one CTA needs for one fill (TC_TPS consecutive tiles) are contiguous:
//   out[((tile / TPS) * 2 + rank) * TPS + tile % TPS][r],  rank = row / 128, r = row % 128
__global__ void __launch_bounds__(256) k_pack_tiles(const uint32_t* __restrict__ words, int64_t n, int64_t ld,
                                                    const int32_t* __restrict__ tile_kt,
                                                    const int32_t* __restrict__ tile_rb, int64_t n_tiles,
                                                    ulonglong2* __restrict__ out) {
    const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (idx >= n_tiles * TC_PROWS) return;
    const int64_t tile = idx / TC_PROWS;
    const int r256 = static_cast<int>(idx % TC_PROWS);
    const int kt = tile_kt[tile], rb = tile_rb[tile];
    uint64_t w = 0;
    if (rb >= 0) {
        const int64_t row = static_cast<int64_t>(rb) * TC_PROWS + r256;
        if (row < n) {
            const uint2 v = *reinterpret_cast<const uint2*>(words + row * ld + 2 * kt);
            w = static_cast<uint64_t>(v.x) | (static_cast<uint64_t>(v.y) << 32);
        }
    }
    const int64_t g = tile / TC_TPS, t = tile % TC_TPS;
    const int rank = r256 / TC_ROWS, r = r256 % TC_ROWS;
    out[(((g * 2 + rank) * TC_TPS) + t) * TC_ROWS + r] = split_mask(w);
}

// The matrices the digit GEMM multiplies the neighborhoods with are functions of the attribute matrix, evaluated on
// the fly: XF_VALUE nan0(B) (the 'sum' score); and for the z-score (safe_extras.py:19-31) XF_SQUARE nan0(B^2) with
// the square rounded like np.power(B, 2) rounds it (plus the 0/1 matrix of the non-NaN entries, see k_quantize_z).
enum : int { XF_VALUE = 0, XF_SQUARE = 1 };
template <int XF, class T>
__device__ __forceinline__ double xf_value(T v) {  // NaN = "contributes nothing"
    if (XF == XF_VALUE) return static_cast<double>(v);
    return v == v ? sq_like_numpy<T>(v) : static_cast<double>(v);
}

// per-column exponent range of the operand: kmax = exponent of the largest magnitude, lmin = exponent of the lowest
// set mantissa bit over all non-zero values, vmax = the largest magnitude itself (bit pattern of the double; these
// order like the values).  flags: bit0 = some value is +-inf
template <class T, int XF>
__global__ void k_col_range(const T* __restrict__ b, int64_t n, int64_t m, int32_t* __restrict__ kmax,
                            int32_t* __restrict__ lmin, unsigned long long* __restrict__ vmax,
                            int32_t* __restrict__ flags) {
    const int64_t j = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (j >= m) return;
    const int64_t rows_per = (n + gridDim.y - 1) / gridDim.y;
    const int64_t r0 = blockIdx.y * rows_per, r1 = min(n, r0 + rows_per);
    int hi = INT_MIN, lo = INT_MAX, bad = 0;
    double big = 0.0;
    for (int64_t r = r0; r < r1; ++r) {
        const double v = xf_value<XF, T>(b[r * m + j]);
        if (v != v || v == 0.0) continue;
        if (isinf(v)) {
            bad = 1;
            continue;
        }
        int e;
        const double f = frexp(fabs(v), &e);  // |v| = f * 2^e, f in [0.5, 1)
        const long long mant = static_cast<long long>(ldexp(f, 53));
        hi = max(hi, e - 1);
        lo = min(lo, e - 53 + (__ffsll(mant) - 1));
        big = fmax(big, fabs(v));
    }
    if (hi != INT_MIN) {
        atomicMax(&kmax[j], hi);
        atomicMin(&lmin[j], lo);
        atomicMax(&vmax[j], static_cast<unsigned long long>(__double_as_longlong(big)));
    }
    if (bad) atomicOr(flags, 1);
}

// fixed-point digits: q = rint(v * 2^shift[j]); stored are the D balanced base-256 int8 digits of q.
// RECORDS = false (M < 64): plane d at digits + d*n*mpad.
// RECORDS = true: one 64*D-byte record per (column group, node), digits[(cg * n + r) * 64 D + d * 64 + (j & 63)] --
// what k_gather copies (192 contiguous bytes at D = 3), one column group's records being a compact N * 64 D byte
// slab that stays L2-resident while the permutations of a batch are gathered from it.
template <class T, int D, bool RECORDS, int XF>
__global__ void k_quantize(const T* __restrict__ b, int64_t n, int64_t m, int64_t mpad,
                           const int32_t* __restrict__ shift, int8_t* __restrict__ digits) {
    const int64_t total = n * mpad;
    int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    const int64_t step = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (; idx < total; idx += step) {
        const int64_t r = idx / mpad, j = idx % mpad;
        int q = 0;
        if (j < m) {
            const double v = xf_value<XF, T>(b[r * m + j]);
            if (v == v) q = static_cast<int>(rint(ldexp(v, shift[j])));
        }
#pragma unroll
        for (int d = 0; d < D; ++d) {
            int dig;
            if (d == D - 1) {
                dig = q;
            } else {
                dig = static_cast<int>(static_cast<int8_t>(q & 0xff));
                q = (q - dig) >> 8;
            }
            const size_t at = RECORDS ? (static_cast<size_t>(j >> 6) * n + r) * (64 * D) + d * 64 + (j & 63)
                                      : static_cast<size_t>(d) * total + idx;
            digits[at] = static_cast<int8_t>(dig);
        }
    }
}

// Bcat tiles for a batch: tile (slot, kt) = 64 K-rows x 64*D columns, stored as two halves of 32*D columns (one per
// CTA of a pair), each in MN-major core-matrix order.  With k = kg*8 + r and 16-column chunk nc (columns nc*16 ..):
//   half = nc / (NC16/2),   16-byte chunk index inside the half = kg * (NC16/2 * 8) + (nc % (NC16/2)) * 8 + r
// column nc*16+x of the tile = digit plane (nc / 4), attribute/permutation column c = (nc & 3) * 16 + x.
template <int D>
__device__ __forceinline__ int gather_chunk_pos(int k, int nc) {
    constexpr int HC = 2 * D;  // 16-column chunks per half
    return (nc / HC) * (TC_KT * HC) + (k >> 3) * (HC * 8) + (nc % HC) * 8 + (k & 7);
}

// M >= 64: one block builds the tile of one (column group, permutation slot, k-tile) from the 64 source records.
// Tile row p (the MMA's K position) holds the record of tile entry tc_kpos(p).  The grid runs k-tiles fastest, then
// permutation slots, then column groups: while the permutations of a batch are gathered, the only records touched
// are those of one column group (N * 64 D bytes), i.e. the reads are L2 hits and DRAM sees the tile writes only.
template <int D>
__global__ void __launch_bounds__(256) k_gather(const int8_t* __restrict__ records, const int32_t* __restrict__ perm,
                                                const int32_t* __restrict__ order, int64_t n, int32_t n_kt,
                                                int32_t n_cg, int8_t* __restrict__ bcat) {
    constexpr int NC16 = 4 * D;
    constexpr int CHUNKS = TC_KT * NC16;       // 16-byte chunks per tile
    // one pad chunk per 8 (a 128-byte core matrix becomes 144 bytes): the NC16 chunks of a record, which land 128
    // bytes apart in the tile, then fall into different banks
    __shared__ uint4 s_tile[CHUNKS + CHUNKS / 8];
    __shared__ int32_t s_src[TC_KT];
    const int kt = blockIdx.x;
    const int q = blockIdx.y;
    const int cg = blockIdx.z;
    if (threadIdx.x < TC_KT) {
        const int64_t t = static_cast<int64_t>(kt) * TC_KT + tc_kpos(threadIdx.x);
        int32_t src = -1;
        if (t < n) {
            const int64_t node = order ? order[t] : t;  // internal position t holds the caller's node `node`
            src = perm ? perm[static_cast<int64_t>(q) * n + node] : static_cast<int32_t>(node);
        }
        s_src[threadIdx.x] = src;
    }
    __syncthreads();
    const uint4* const rec = reinterpret_cast<const uint4*>(records + static_cast<size_t>(cg) * n * (64 * D));
    for (int idx = threadIdx.x; idx < CHUNKS; idx += blockDim.x) {
        const int k = idx / NC16, nc = idx % NC16;  // NC16 consecutive threads read one contiguous record
        const int32_t src = s_src[k];
        uint4 v = make_uint4(0, 0, 0, 0);
        if (src >= 0) v = __ldg(rec + static_cast<size_t>(src) * NC16 + nc);
        const int at = gather_chunk_pos<D>(k, nc);
        s_tile[at + (at >> 3)] = v;
    }
    __syncthreads();
    // streaming stores: the tiles are read by the GEMM much later; they must not evict the records from L2
    const size_t slot = static_cast<size_t>(q) * n_cg + cg;
    uint4* dst = reinterpret_cast<uint4*>(bcat + (slot * n_kt + kt) * (TC_KT * 64 * D));
    for (int i = threadIdx.x; i < CHUNKS; i += blockDim.x) __stcs(dst + i, s_tile[i + (i >> 3)]);
}

// M < 64: a slot holds pps = 64 / mpad permutations of all attributes; byte-wise assembly (tiny problems only)
template <int D>
__global__ void __launch_bounds__(256) k_gather_small(const int8_t* __restrict__ digits,
                                                      const int32_t* __restrict__ perm,
                                                      const int32_t* __restrict__ order, int64_t n, int64_t mpad,
                                                      int32_t n_kt, int32_t pps, int32_t log2_mpad,
                                                      int32_t batch_perms, int8_t* __restrict__ bcat) {
    constexpr int NC16 = 4 * D;
    constexpr int CHUNKS = TC_KT * NC16;
    const int kt = blockIdx.x;
    const int q = blockIdx.y;  // n_cg == 1: slot == q
    uint4* dst = reinterpret_cast<uint4*>(bcat + (static_cast<size_t>(q) * n_kt + kt) * (TC_KT * 64 * D));
    const size_t plane = static_cast<size_t>(n) * mpad;
    for (int idx = threadIdx.x; idx < CHUNKS; idx += blockDim.x) {
        const int k = idx / NC16, nc = idx % NC16;
        const int64_t t = static_cast<int64_t>(kt) * TC_KT + tc_kpos(k);  // tile row k = the MMA's K position
        const int d = nc >> 2, jc = nc & 3;
        uint32_t w[4] = {0, 0, 0, 0};
        if (t < n) {
            const int64_t node = order ? order[t] : t;
#pragma unroll
            for (int x = 0; x < 16; ++x) {
                const int c = jc * 16 + x;
                const int pl = q * pps + (c >> log2_mpad);
                const int j = c & (static_cast<int>(mpad) - 1);
                if (pl < batch_perms) {
                    const int64_t src = perm ? perm[static_cast<int64_t>(pl) * n + node] : node;
                    const uint32_t byte = static_cast<uint8_t>(digits[d * plane + src * mpad + j]);
                    w[x >> 2] |= byte << ((x & 3) * 8);
                }
            }
        }
        dst[gather_chunk_pos<D>(k, nc)] = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

__global__ void k_unpack_counts(uint32_t* __restrict__ cpk, int64_t cells, uint32_t* __restrict__ cneg,
                                uint32_t* __restrict__ cpos) {
    int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    const int64_t step = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (; i < cells; i += step) {
        const uint32_t v = cpk[i];
        if (v) {
            cpos[i] += v >> 16;
            cneg[i] += v & 0xffffu;
            cpk[i] = 0;
        }
    }
}

// ------------------------------------------------------------------------------------------------ plan
// One matrix the neighborhoods are multiplied with (see XF_*): fixed-point digits of its columns
struct TcOperand {
    int D = 0;                 // digit planes
    bool built = false;
    bool usable = true;        // false: +-inf among the values
    bool any_inexact = false;  // some column is not exactly representable (its comparisons carry an error band)
    DevBuf<int8_t> digits;
    DevBuf<int32_t> shift;     // per-column binary exponent of the fixed point: q = rint(v * 2^shift[j])
    DevBuf<uint8_t> inexact;   // [mpad]
    DevBuf<int64_t> s0fix;     // observed fixed-point scores [n_rb * 256][mpad], internal row order
};

struct TcPlan {
    TcOperand op;              // the 'sum' operand; in the z plan: the z records
    int64_t n = 0, m = 0, mpad = 0;
    int32_t n_rb = 0, n_kt = 0, n_cg = 0, pps = 1, log2_mpad = 0;
    int64_t max_nb = 0;        // largest neighborhood
    int64_t n_tiles = 0;       // stored tiles (row blocks padded to a multiple of TC_TPS)
    int64_t n_tiles_real = 0;  // non-empty tiles
    int32_t band_rb = 0, n_bands = 1, band_kt = 0;  // L2 blocking: row blocks per band, max distinct k-tiles of a band
    bool usable = true;  // false: data contains +-inf -> SIMT engine
    DevBuf<ulonglong2> a_bits;
    const int32_t* order = nullptr;  // e->order.p when the caller supplied a node order (internal row -> node)
    DevBuf<int32_t> tile_ptr, tile_kt, tile_rb;
    DevBuf<unsigned int> flag_count;
    int64_t cpk_perms = 0;  // permutations accumulated in the packed counters since the last unpack (16-bit fields)
    uint32_t* cpk = nullptr;  // packed counters of the call in progress (context scratch or the caller's array)
    unsigned int flag_cap = 0;
    // z-score null: the 'sum' plan owns a second plan whose column groups are 32 attributes x 6 planes (build_plan_z);
    // the fields below belong to that second plan
    TcPlan* z = nullptr;
    DevBuf<int32_t> z_shift2;          // fixed point of the squares
    DevBuf<uint8_t> z_inex2;
    DevBuf<float4> z_col;              // filter constants per attribute (k_zcol)
    const int32_t* z_shift1 = nullptr;  // fixed point of the values: the owning plan's
    const uint8_t* z_inex1 = nullptr;
    const double* z_z0t = nullptr;     // observed z-scores of the null in progress, plan layout
    ~TcPlan() { delete z; }
};

void tc_plan_destroy(TcPlan* p) { delete p; }

static void print_prof(sb_ctx* ctx, const long long* d_prof, const char* what) {
    long long h[16];
    SB_CUDA(cudaMemcpyAsync(h, d_prof, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
    SB_CUDA(cudaStreamSynchronize(ctx->stream));
    const double fills = std::max<double>(1.0, static_cast<double>(h[10]));
    fprintf(stderr,
            "[sb_trace] %s: cycles per fill (%d k-tiles): producer %.0f (wait empty %.0f) | expander %.0f (wait full "
            "%.0f, wait aempty %.0f, publish %.0f, expand alu %.0f) | mma %.0f (wait A %.0f, wait tempty %.0f) | "
            "epilogue %.0f (wait tfull %.0f)\n",
            what, TC_TPS, h[0] / fills, h[1] / fills, h[2] / fills, h[12] / fills, h[3] / fills, h[11] / fills,
            h[13] / fills,
            h[4] / fills, h[5] / fills, h[7] / fills, h[8] / fills, h[9] / fills);
}

template <int D, int KIND, bool SMALL_M, bool PROF>
static void launch_gemm(sb_ctx* ctx, const GemmParams& gp, int grid) {
    using C = TcCfg<D>;
    static bool configured = false;
    if (!configured) {
        SB_CUDA(cudaFuncSetAttribute(k_gemm<D, KIND, SMALL_M, PROF>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     C::SMEM));
        configured = true;
    }
    KernelTimer kt(ctx, SB_K_GEMM);
    // `grid` counts CTA pairs; the kernel carries __cluster_dims__(2, 1, 1)
    k_gemm<D, KIND, SMALL_M, PROF><<<2 * grid, TC_THREADS, C::SMEM, ctx->stream>>>(gp);
    SB_LAUNCH_CHECK(ctx);
}

template <int D, bool PROF>
static void launch_gemm_k(sb_ctx* ctx, int kind, bool small_m, const GemmParams& gp, int grid) {
    if (kind == TCK_Z) {
        if constexpr (D == 3)  // six planes of 32 attributes: always the 192-column configuration
            launch_gemm<3, TCK_Z, false, PROF>(ctx, gp, grid);
    } else if (kind == TCK_RAW)
        launch_gemm<D, TCK_RAW, false, PROF>(ctx, gp, grid);
    else if (kind == TCK_STORE)
        small_m ? launch_gemm<D, TCK_STORE, true, false>(ctx, gp, grid)
                : launch_gemm<D, TCK_STORE, false, false>(ctx, gp, grid);
    else
        small_m ? launch_gemm<D, TCK_COUNT, true, PROF>(ctx, gp, grid)
                : launch_gemm<D, TCK_COUNT, false, PROF>(ctx, gp, grid);
}

static void launch_gemm_d(sb_ctx* ctx, int D, const GemmParams& gp_in, int grid) {
    GemmParams gp = gp_in;
    if (gp.band_rb <= 0) {
        gp.band_rb = gp.n_rb;
        gp.n_bands = 1;
    }
    ctx->ws_counter.reserve(1);
    gp.unit_counter = ctx->ws_counter.p;
    SB_CUDA(cudaMemsetAsync(gp.unit_counter, 0, sizeof(unsigned int), ctx->stream));
    const int kind = (gp.mode & TCM_Z) ? TCK_Z : (gp.mode & TCM_RAW) ? TCK_RAW : (gp.mode & TCM_STORE) ? TCK_STORE : TCK_COUNT;
    const bool small_m = gp.mpad < 64;
    if (gp.prof) {  // cycle accounting compiled in (SB_TRACE / rate probe)
        if (D == 1)
            launch_gemm_k<1, true>(ctx, kind, small_m, gp, grid);
        else if (D == 2)
            launch_gemm_k<2, true>(ctx, kind, small_m, gp, grid);
        else
            launch_gemm_k<3, true>(ctx, kind, small_m, gp, grid);
        return;
    }
    if (D == 1)
        launch_gemm_k<1, false>(ctx, kind, small_m, gp, grid);
    else if (D == 2)
        launch_gemm_k<2, false>(ctx, kind, small_m, gp, grid);
    else
        launch_gemm_k<3, false>(ctx, kind, small_m, gp, grid);
}

// Gathered operand tiles of a batch of permutations -> bcat, on stream st
static void launch_gather(sb_ctx* ctx, const TcPlan* pl, const TcOperand& op, const int32_t* perm, int slots,
                          int batch_perms, int8_t* bcat, cudaStream_t st) {
    KernelTimer kt(ctx, SB_K_GATHER, st);
    if (pl->mpad >= 64) {
        const int nq = slots / pl->n_cg;
        dim3 grid(static_cast<unsigned>(pl->n_kt), static_cast<unsigned>(nq), static_cast<unsigned>(pl->n_cg));
        SB_CHECK(grid.y <= 65535 && grid.z <= 65535, "too many column groups / permutations in one batch");
#define SB_G(DD) \
    k_gather<DD><<<grid, 256, 0, st>>>(op.digits.p, perm, pl->order, pl->n, pl->n_kt, pl->n_cg, bcat)
        if (op.D == 1)
            SB_G(1);
        else if (op.D == 2)
            SB_G(2);
        else
            SB_G(3);
#undef SB_G
    } else {
        dim3 grid(static_cast<unsigned>(pl->n_kt), static_cast<unsigned>(slots));
        SB_CHECK(grid.y <= 65535, "too many column slots in one batch (%d)", slots);
#define SB_G(DD)                                                                                                  \
    k_gather_small<DD><<<grid, 256, 0, st>>>(op.digits.p, perm, pl->order, pl->n, pl->mpad, pl->n_kt, pl->pps,    \
                                             pl->log2_mpad, batch_perms, bcat)
        if (op.D == 1)
            SB_G(1);
        else if (op.D == 2)
            SB_G(2);
        else
            SB_G(3);
#undef SB_G
    }
    SB_LAUNCH_CHECK(ctx);
}

static GemmParams base_params(sb_enrich* e, TcPlan* pl, const TcOperand& op) {
    sb_ctx* ctx = e->ctx;
    GemmParams gp{};
    gp.a_bits = pl->a_bits.p;
    gp.tile_ptr = pl->tile_ptr.p;
    gp.tile_kt = pl->tile_kt.p;
    gp.bcat = ctx->ws_bcat.p;
    gp.n_kt = pl->n_kt;
    gp.n_rb = pl->n_rb;
    gp.n_cg = pl->n_cg;
    gp.n = pl->n;
    gp.m = pl->m;
    gp.mpad = pl->mpad;
    gp.log2_mpad = pl->log2_mpad;
    gp.pps = pl->pps;
    gp.s0fix = op.s0fix.p;
    gp.row_ptr = e->row_ptr.p;
    gp.node_of_row = pl->order;
    gp.inexact = op.inexact.p;
    gp.flag_ij = ctx->ws_flag_ij.p;
    gp.flag_p = ctx->ws_flag_p.p;
    gp.flag_count = pl->flag_count.p;
    gp.flag_cap = pl->flag_cap / static_cast<unsigned int>(pl->n_cg);
    gp.cpk = pl->cpk;
    const uint32_t ncols = 64u * op.D;
    gp.b_lbo = ncols / 2 * 8;  // MN-major B half tile: stride between 8-row K groups
    gp.b_sbo = 128;        //             stride between 16-column chunks
    gp.q_wrap = INT_MAX;
    gp.b_kstep = 4 * (ncols / 2) * 8;
    if (pl->z_col.p) {  // z-score plan
        gp.z0t = pl->z_z0t;
        gp.zcol = pl->z_col.p;
        gp.zshift1 = pl->z_shift1;
        gp.zshift2 = pl->z_shift2.p;
        gp.zinex1 = pl->z_inex1;
        gp.zinex2 = pl->z_inex2.p;
        gp.rows_pad = static_cast<int64_t>(pl->n_rb) * TC_PROWS;
        gp.flag_cap = pl->flag_cap;  // one list, no buckets
    }
    return gp;
}

static int64_t slots_for(const TcPlan* pl, int64_t perms) {
    return pl->mpad >= 64 ? perms * pl->n_cg : sb_ceil_div(perms, pl->pps);
}

// Per-column exponent ranges of one operand (XF_*) on the host; returns k_col_range's flags (1: some value is +-inf)
static int32_t column_ranges(sb_enrich* e, int xf, std::vector<int32_t>& h_kmax, std::vector<int32_t>& h_lmin,
                             std::vector<double>& h_vmax) {
    sb_ctx* ctx = e->ctx;
    cudaStream_t st = ctx->stream;
    const int64_t n = e->n, m = e->m;
    DevBuf<int32_t> kmax, lmin, flags;
    DevBuf<unsigned long long> vmax;
    kmax.reserve(m);
    lmin.reserve(m);
    vmax.reserve(m);
    flags.reserve(1);
    h_kmax.assign(m, INT_MIN);
    h_lmin.assign(m, INT_MAX);
    h_vmax.assign(m, 0.0);
    SB_CUDA(cudaMemsetAsync(vmax.p, 0, m * sizeof(unsigned long long), st));
    SB_CUDA(cudaMemcpyAsync(kmax.p, h_kmax.data(), m * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    SB_CUDA(cudaMemcpyAsync(lmin.p, h_lmin.data(), m * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    SB_CUDA(cudaMemsetAsync(flags.p, 0, sizeof(int32_t), st));
    dim3 cgrid(static_cast<unsigned>(sb_ceil_div(m, 128)),
               static_cast<unsigned>(std::max<int64_t>(1, std::min<int64_t>(64, n / 256))));
#define SB_R(T, X) \
    k_col_range<T, X><<<cgrid, 128, 0, st>>>(static_cast<const T*>(e->b), n, m, kmax.p, lmin.p, vmax.p, flags.p)
#define SB_RX(T)                 \
    do {                         \
        if (xf == XF_VALUE)      \
            SB_R(T, XF_VALUE);   \
        else                     \
            SB_R(T, XF_SQUARE);  \
    } while (0)
    if (e->dtype == SB_F32)
        SB_RX(float);
    else
        SB_RX(double);
#undef SB_RX
#undef SB_R
    SB_LAUNCH_CHECK(ctx);
    int32_t h_flags = 0;
    SB_CUDA(cudaMemcpyAsync(h_kmax.data(), kmax.p, m * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaMemcpyAsync(h_lmin.data(), lmin.p, m * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaMemcpyAsync(&h_flags, flags.p, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    static_assert(sizeof(double) == sizeof(unsigned long long), "bit patterns");
    SB_CUDA(cudaMemcpyAsync(h_vmax.data(), vmax.p, m * sizeof(double), cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaStreamSynchronize(st));
    return h_flags;
}

// Fixed point of a column that D balanced digits cannot hold exactly: the largest power-of-two scale whose largest
// value still fits.  D digits hold |q| <= qmax = 127 (256^D - 1) / 255 -- almost 2^(8D-1) -- so the column's exact
// maximum buys one bit over the exponent-based choice |q| < 2^(8D-2) unless its mantissa sits in the top 0.4 %, and
// every bit halves the error band, i.e. the number of comparisons that go to the exact fix-up.  `extra_bit` is off
// where the wider sums could overflow the epilogue's int32 differences (neighborhoods of 32768+ nodes).
static int inexact_shift(int D, int kmax, double vmax, bool extra_bit) {
    const int shift = (8 * D - 3) - kmax;  // |q| < 2^(8D-2)
    const double qmax = 127.0 * (std::pow(256.0, D) - 1.0) / 255.0;
    if (extra_bit && std::rint(std::ldexp(vmax, shift + 1)) <= qmax) return shift + 1;
    return shift;
}

// Digit planes of the 'sum' operand and its observed fixed-point scores
static void build_operand(sb_enrich* e, TcPlan* pl) {
    sb_ctx* ctx = e->ctx;
    cudaStream_t st = ctx->stream;
    const int xf = XF_VALUE;
    TcOperand& op = pl->op;
    if (op.built) return;
    const int64_t n = e->n, m = e->m;
    PhaseTrace* tr = new PhaseTrace(ctx, "tc.plan.digits");
    std::vector<int32_t> h_kmax, h_lmin;
    std::vector<double> h_vmax;
    const int32_t h_flags = column_ranges(e, xf, h_kmax, h_lmin, h_vmax);
    op.built = true;
    if (h_flags & 1) {
        op.usable = false;
        delete tr;
        return;
    }
    // digits needed for exactness: a column needs bits = kmax - lmin + 1 magnitude bits; D balanced digits hold
    // |q| <= 127 * (256^D - 1) / 255, i.e. bits <= 8 D - 2
    int D = 1;
    for (int64_t j = 0; j < m; ++j)
        if (h_kmax[j] != INT_MIN) {
            const int bits = h_kmax[j] - h_lmin[j] + 1;
            D = std::max(D, std::min(3, (bits + 2 + 7) / 8));
        }
    op.D = D;
    std::vector<int32_t> h_shift(m, 0);
    std::vector<uint8_t> h_inexact(pl->mpad, 0);
    for (int64_t j = 0; j < m; ++j) {
        if (h_kmax[j] == INT_MIN) continue;
        const int bits = h_kmax[j] - h_lmin[j] + 1;
        if (bits <= 8 * D - 2) {
            h_shift[j] = -h_lmin[j];  // lowest set bit lands on 2^0: every value is an exact integer
        } else {
            h_shift[j] = inexact_shift(D, h_kmax[j], h_vmax[j], pl->max_nb < 32768);
            h_inexact[j] = 1;
            op.any_inexact = true;
        }
    }
    op.shift.reserve(m);
    op.inexact.reserve(pl->mpad);
    SB_CUDA(cudaMemcpyAsync(op.shift.p, h_shift.data(), m * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    SB_CUDA(cudaMemcpyAsync(op.inexact.p, h_inexact.data(), pl->mpad, cudaMemcpyHostToDevice, st));
    op.digits.reserve(static_cast<size_t>(D) * n * pl->mpad);
    const unsigned qblocks =
        static_cast<unsigned>(std::min<int64_t>(sb_ceil_div(n * pl->mpad, 256), ctx->num_sms * 32));
#define SB_Q(T, DD, X)                                                                                              \
    do {                                                                                                            \
        if (pl->mpad >= 64)                                                                                         \
            k_quantize<T, DD, true, X><<<qblocks, 256, 0, st>>>(static_cast<const T*>(e->b), n, m, pl->mpad,        \
                                                                op.shift.p, op.digits.p);                           \
        else                                                                                                        \
            k_quantize<T, DD, false, X><<<qblocks, 256, 0, st>>>(static_cast<const T*>(e->b), n, m, pl->mpad,       \
                                                                 op.shift.p, op.digits.p);                          \
    } while (0)
#define SB_QD(T, X)          \
    do {                     \
        if (D == 1)          \
            SB_Q(T, 1, X);   \
        else if (D == 2)     \
            SB_Q(T, 2, X);   \
        else                 \
            SB_Q(T, 3, X);   \
    } while (0)
    if (e->dtype == SB_F32)
        SB_QD(float, XF_VALUE);
    else
        SB_QD(double, XF_VALUE);
#undef SB_QD
#undef SB_Q
    SB_LAUNCH_CHECK(ctx);
    SB_CUDA(cudaStreamSynchronize(st));  // host vectors above go out of scope

    delete tr;
    tr = new PhaseTrace(ctx, "tc.plan.s0fix");
    // ---- observed fixed-point scores: one identity-permutation pass through the same kernel
    ctx->ws_flag_ij.reserve(pl->flag_cap);
    ctx->ws_flag_p.reserve(pl->flag_cap);
    const size_t tile_b = static_cast<size_t>(TC_KT) * 64 * D;
    const int64_t slots1 = slots_for(pl, 1);
    ctx->ws_bcat.reserve(static_cast<size_t>(slots1) * pl->n_kt * tile_b);
    op.s0fix.reserve(static_cast<size_t>(pl->n_rb) * TC_PROWS * pl->mpad);
    launch_gather(ctx, pl, op, nullptr, static_cast<int>(slots1), 1, ctx->ws_bcat.p, st);
    GemmParams gp = base_params(e, pl, op);
    gp.mode = TCM_STORE;
    gp.q_total = 1;
    gp.q_chunks = 1;
    gp.q_per = 1;
    gp.batch_perms = 1;
    const int units = pl->n_rb * pl->n_cg;
    launch_gemm_d(ctx, D, gp, std::min(units, ctx->num_sms / 2));
    SB_CUDA(cudaStreamSynchronize(st));
    delete tr;
}

static TcPlan* build_plan(sb_enrich* e, bool zgroups = false) {
    sb_ctx* ctx = e->ctx;
    cudaStream_t st = ctx->stream;
    TcPlan* pl = new TcPlan;
    try {
        KernelTimer kt_prep(ctx, SB_K_PREP);
        PhaseTrace tr_all(ctx, "tc.build_plan");
        const int64_t n = e->n, m = e->m;
        pl->n = n;
        pl->m = m;
        if (zgroups) {  // z-score plan: a group is 32 attributes (x 6 planes = the 192 columns of an accumulation)
            pl->n_cg = static_cast<int32_t>(sb_ceil_div(m, 32));
            pl->mpad = static_cast<int64_t>(pl->n_cg) * 64;
            pl->pps = 1;
            pl->log2_mpad = 0;
        } else if (m >= 64) {
            pl->mpad = sb_ceil_div(m, 64) * 64;
            pl->n_cg = static_cast<int32_t>(pl->mpad / 64);
            pl->pps = 1;
            pl->log2_mpad = 0;
        } else {
            int lg = 0;
            while ((1 << lg) < m) ++lg;
            pl->mpad = 1 << lg;
            pl->log2_mpad = lg;
            pl->n_cg = 1;
            pl->pps = static_cast<int32_t>(64 / pl->mpad);
        }
        pl->n_rb = static_cast<int32_t>(sb_ceil_div(n, TC_PROWS));
        pl->n_kt = static_cast<int32_t>(sb_ceil_div(n, TC_KT));

        // the epilogue compares in int32 (score = hi * 256 + lo): needs n_i * max|q| < 2^38 (see inexact_shift)
        {
            std::vector<int64_t> h_rp(n + 1);
            SB_CUDA(cudaMemcpyAsync(h_rp.data(), e->row_ptr.p, (n + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
            SB_CUDA(cudaStreamSynchronize(st));
            int64_t max_nb = 0;
            for (int64_t i = 0; i < n; ++i) max_nb = std::max(max_nb, h_rp[i + 1] - h_rp[i]);
            pl->max_nb = max_nb;
            if (max_nb >= 65536) {
                pl->usable = false;  // neighborhoods of 65536+ nodes: exact SIMT engine
                return pl;
            }
        }

        // ---- internal node order (optional): tiles are built from the permuted matrix, everything that addresses
        // the caller's arrays (gather source rows, counts, bands, fix-ups) goes through order[]
        const uint32_t* a_words = e->a->words;
        DevBuf<uint32_t> permuted;
        if (e->have_order) {
            pl->order = e->order.p;
            permuted.reserve(static_cast<size_t>(n) * e->a->ld);
            SB_CUDA(cudaMemsetAsync(permuted.p, 0, static_cast<size_t>(n) * e->a->ld * sizeof(uint32_t), st));
            k_permute_packed<<<static_cast<unsigned>(sb_ceil_div(n * 32, 256)), 256, 0, st>>>(
                e->row_ptr.p, e->col_idx.p, e->order_inv.p, n, e->a->ld, permuted.p);
            SB_LAUNCH_CHECK(ctx);
            a_words = permuted.p;
        }
        // ---- A tiles
        PhaseTrace* tr = new PhaseTrace(ctx, "tc.plan.a_tiles");
        DevBuf<uint8_t> occ;
        DevBuf<int32_t> rb_count;
        occ.reserve(static_cast<size_t>(pl->n_rb) * pl->n_kt);
        rb_count.reserve(pl->n_rb);
        k_tile_occ<<<pl->n_rb, 256, 0, st>>>(a_words, n, e->a->ld, pl->n_kt, occ.p, rb_count.p);
        SB_LAUNCH_CHECK(ctx);
        std::vector<int32_t> h_cnt(pl->n_rb), h_ptr(pl->n_rb + 1);
        SB_CUDA(cudaMemcpyAsync(h_cnt.data(), rb_count.p, pl->n_rb * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        SB_CUDA(cudaStreamSynchronize(st));
        {
            // bands of about one row block per CTA pair: all pairs then work on the same (q chunk, column group) and
            // the gathered slab they share is only what the band's neighborhoods reach
            const int pairs = std::max(1, ctx->num_sms / 2);
            pl->n_bands = std::max(1, static_cast<int>((pl->n_rb + pairs / 2) / pairs));
            pl->band_rb = static_cast<int32_t>(sb_ceil_div(pl->n_rb, pl->n_bands));
            pl->n_bands = static_cast<int32_t>(sb_ceil_div(pl->n_rb, pl->band_rb));
            DevBuf<int32_t> band_cnt;
            band_cnt.reserve(pl->n_bands);
            k_band_kt<<<pl->n_bands, 256, 0, st>>>(occ.p, pl->n_rb, pl->n_kt, pl->band_rb, band_cnt.p);
            SB_LAUNCH_CHECK(ctx);
            std::vector<int32_t> h_band(pl->n_bands);
            SB_CUDA(cudaMemcpyAsync(h_band.data(), band_cnt.p, pl->n_bands * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
            SB_CUDA(cudaStreamSynchronize(st));
            pl->band_kt = *std::max_element(h_band.begin(), h_band.end());
        }
        int64_t run = 0, real = 0;
        for (int i = 0; i < pl->n_rb; ++i) {
            h_ptr[i] = static_cast<int32_t>(run);
            real += h_cnt[i];
            run += sb_ceil_div(h_cnt[i], TC_TPS) * TC_TPS;
        }
        pl->n_tiles_real = real;
        SB_CHECK(run < (1ll << 31), "too many non-empty neighborhood tiles (%lld)", (long long)run);
        h_ptr[pl->n_rb] = static_cast<int32_t>(run);
        pl->n_tiles = run;
        pl->tile_ptr.reserve(pl->n_rb + 1);
        pl->tile_kt.reserve(run);
        pl->tile_rb.reserve(run);
        pl->a_bits.reserve(static_cast<size_t>(run) * TC_PROWS);
        SB_CUDA(cudaMemcpyAsync(pl->tile_ptr.p, h_ptr.data(), (pl->n_rb + 1) * sizeof(int32_t),
                                cudaMemcpyHostToDevice, st));
        k_tile_list<<<pl->n_rb, 256, 0, st>>>(occ.p, pl->n_kt, pl->tile_ptr.p, pl->tile_kt.p, pl->tile_rb.p);
        SB_LAUNCH_CHECK(ctx);
        k_pack_tiles<<<static_cast<unsigned>(sb_ceil_div(run * TC_PROWS, 256)), 256, 0, st>>>(
            a_words, n, e->a->ld, pl->tile_kt.p, pl->tile_rb.p, run, pl->a_bits.p);
        SB_LAUNCH_CHECK(ctx);

        delete tr;
        // ---- flag list.  Only inexact columns can flag (~3e-5 of their comparisons for N(0,1) data); the list is
        // sized for a quarter of the cells per launch (4M .. 128M entries of 12 bytes) instead of the worst case, and
        // an overflowing launch is redone in (slot, row-block range) pieces whose worst case fits (tc_perm_counts).
        const int64_t min_cap = static_cast<int64_t>(TC_PROWS) * 64 * pl->n_cg;  // >= one row block per bucket
        pl->flag_cap = static_cast<unsigned int>(min_cap);
        pl->flag_count.reserve(pl->n_cg);
        if (zgroups) return pl;  // build_plan_z fills in the records
        build_operand(e, pl);
        pl->usable = pl->op.usable;
        if (pl->op.any_inexact) {  // exactly representable data never flags: nothing to reserve
            const int64_t padded_cells = static_cast<int64_t>(pl->n_rb) * TC_PROWS * pl->mpad;
            int64_t cap = std::min<int64_t>(std::max<int64_t>(padded_cells / 4, 4ll << 20), 128ll << 20);
            if (getenv("SB_FLAG_CAP")) cap = atoll(getenv("SB_FLAG_CAP"));  // tests: force the overflow recovery
            cap = std::max<int64_t>(cap, min_cap);
            pl->flag_cap = static_cast<unsigned int>(sb_ceil_div(cap, pl->n_cg) * pl->n_cg);
        }
    } catch (...) {
        delete pl;
        throw;
    }
    return pl;
}

// one GEMM launch over slots [q_first, q_first + q_total) of the prepared batch and row blocks [rb0, rb0 + n_rb)
static void run_batch_gemm(sb_enrich* e, TcPlan* pl, const TcOperand& op, int mode, int q_first, int q_total,
                           int batch_perms, int rb0, int n_rb) {
    sb_ctx* ctx = e->ctx;
    GemmParams gp = base_params(e, pl, op);
    gp.mode = mode;
    gp.q_total = q_total;
    gp.batch_perms = batch_perms;
    gp.rb0 = rb0;
    gp.n_rb = n_rb;
    const size_t tile_b = static_cast<size_t>(TC_KT) * 64 * op.D;
    // slot offset into the gathered operand of the batch
    gp.bcat += static_cast<size_t>(q_first) * pl->n_cg * pl->n_kt * tile_b;
    // L2 blocking (see decode_unit; bands are chosen in build_plan).  The number of slots per unit (q_per) trades
    // per-unit epilogue overhead (observed-score preload, count flush) against the gathered slab a band touches per
    // (q chunk, column group): measured on C3, 13+ slots per unit (64 MB slabs, the default; SB_SLAB_MB overrides)
    // run 20 % faster than 4 and the 5-stage ring still hides the L2 / HBM latency.
    const bool whole = rb0 == 0 && n_rb == pl->n_rb;
    gp.band_rb = whole ? pl->band_rb : n_rb;
    gp.n_bands = whole ? pl->n_bands : 1;
    const double slab = std::max(1, pl->band_kt) * static_cast<double>(tile_b);  // gathered bytes per (slot, group)
    static const double slab_mb = getenv("SB_SLAB_MB") ? atof(getenv("SB_SLAB_MB")) : 64.0;
    int q_per = static_cast<int>(std::max(1.0, std::min(64.0, slab_mb * (1 << 20) / slab)));
    const int base_units = n_rb * pl->n_cg;
    const int want_chunks = static_cast<int>(sb_ceil_div(4 * (ctx->num_sms / 2), base_units));
    q_per = std::min<int>(q_per, std::max<int>(1, q_total / std::max(1, want_chunks)));
    q_per = std::max(1, std::min(q_per, q_total));
    gp.q_chunks = static_cast<int32_t>(sb_ceil_div(q_total, q_per));
    gp.q_per = static_cast<int32_t>(sb_ceil_div(q_total, gp.q_chunks));  // equal chunks: a short last chunk pays the
                                                                          // per-unit overhead for a few slots only
    const int64_t units = static_cast<int64_t>(base_units) * gp.q_chunks;
    SB_CHECK(units < (1ll << 31), "too many work units in one batch (%lld)", (long long)units);
    static const bool trace = getenv("SB_TRACE") != nullptr;
    DevBuf<long long> d_prof;
    if (trace) {
        d_prof.reserve(16);
        SB_CUDA(cudaMemsetAsync(d_prof.p, 0, 16 * sizeof(long long), ctx->stream));
        gp.prof = d_prof.p;
        fprintf(stderr, "[sb_trace] gemm schedule: n_rb %d n_cg %d q_total %d q_per %d band_rb %d n_bands %d units %lld\n",
                n_rb, pl->n_cg, q_total, gp.q_per, gp.band_rb, gp.n_bands, (long long)units);
    }
    static const int dbg = getenv("SB_DBG") ? atoi(getenv("SB_DBG")) : 0;  // timing experiments only (8, 16)
    gp.dbg = dbg;
    launch_gemm_d(ctx, op.D, gp, static_cast<int>(std::min<int64_t>(units, ctx->num_sms / 2)));
    if (trace) print_prof(ctx, d_prof.p, "batch gemm");
}

void unpack_add_counts(sb_ctx* ctx, uint32_t* packed, int64_t cells, uint32_t* cneg, uint32_t* cpos) {
    const unsigned blocks = static_cast<unsigned>(std::min<int64_t>(sb_ceil_div(cells, 256), ctx->num_sms * 16));
    k_unpack_counts<<<blocks, 256, 0, ctx->stream>>>(packed, cells, cneg, cpos);
    SB_LAUNCH_CHECK(ctx);
}

static void flush_counts(sb_enrich* e, TcPlan* pl, uint32_t* cneg, uint32_t* cpos) {
    unpack_add_counts(e->ctx, pl->cpk, e->n * e->m, cneg, cpos);
    pl->cpk_perms = 0;
}

// Counts of `num_perm` permutations.  Either ADDED to the caller's two arrays (cneg / cpos), or, with `packed`
// (cneg == cpos == nullptr), ADDED to one word per cell, pos << 16 | neg -- the form that crosses NVLink in the
// multi-GPU all-reduce; the caller keeps the number of permutations summed into a word below 65536.
// Observed 'sum' scores from the tensor cores when every attribute column is exactly representable in the plan's fixed
// point (binary / integer / dyadic data -- what the hypergeometric test is given): S[i][j] = s0fix[row_of[i]][j] *
// 2^-shift[j], exact.  Returns false when the plan cannot serve (inexact columns, +-inf, giant neighborhoods).
bool tc_observed_exact(sb_enrich* e, const int64_t** s0fix, const int32_t** shift, const int32_t** row_of_node,
                       int64_t* mpad) {
    if (!e->tc) e->tc = build_plan(e);
    TcPlan* pl = e->tc;
    if (!pl->usable || pl->op.any_inexact) return false;
    *s0fix = pl->op.s0fix.p;
    *shift = pl->op.shift.p;
    *row_of_node = e->have_order ? e->order_inv.p : nullptr;
    *mpad = pl->mpad;
    return true;
}

void tc_perm_counts(sb_enrich* e, const int32_t* perm_dev, int64_t num_perm, uint32_t* cneg, uint32_t* cpos,
                    uint32_t* packed) {
    sb_ctx* ctx = e->ctx;
    cudaStream_t st = ctx->stream;
    PhaseTrace tr_all(ctx, "tc.perm_counts(total)");
    if (!e->tc) e->tc = build_plan(e);
    TcPlan* pl = e->tc;
    if (!pl->usable) {  // +-inf in the data (or neighborhoods of 65536+ nodes): fixed point cannot represent it
        simt_perm_counts(e, SB_SCORE_SUM, perm_dev, num_perm, cneg, cpos, packed);
        return;
    }
    SB_CHECK(!packed || num_perm < 65536, "packed counts hold fewer than 65536 permutations per call");
    PhaseTrace* tr_ws = new PhaseTrace(ctx, "tc.workspace");

    // Batches: the gathered tiles of a batch take <= ~1/4 of the free memory seen at the first null of this context (at
    // most 40 GiB -- C5: 5.53 s per null with 16 GiB batches, 5.31 s with 40; C3 indifferent; cudaMemGetInfo was seen to take hundreds of ms on a busy device: ask once per context).  Per batch:
    // gather -> GEMM -> fix-ups, all on the context's stream and without a host synchronisation in between (the fix-up
    // kernel reads the bucket counters on the device; they are logged and looked at once, after the last batch).
    // Measured and dropped: the gather of batch b + 1 and the fix-ups of batch b - 1 on a second stream under the GEMM
    // of batch b (double-buffered operand and lists).  The kernels do overlap -- their small blocks fit next to a
    // k_gemm CTA -- but the step does not get shorter (C3: 290 / 289 / 292 ms for none / fix-ups / both overlapped):
    // the GEMM slows down by what the side work takes, because the chip runs at its 1 kW power cap (SM clock 1.6-1.7
    // of 1.965 GHz, sw_power_cap active) and a step costs the same energy either way.
    const size_t tile_b = static_cast<size_t>(TC_KT) * 64 * pl->op.D;
    const size_t slot_bytes = static_cast<size_t>(pl->n_kt) * tile_b;
    if (ctx->bcat_budget == 0) {
        size_t free_b = 0, total_b = 0;
        SB_CUDA(cudaMemGetInfo(&free_b, &total_b));
        ctx->bcat_budget = std::max<size_t>(free_b / 4, 1);
    }
    const size_t budget = std::max(std::min<size_t>(std::max<size_t>(ctx->bcat_budget, slot_bytes * slots_for(pl, 1)),
                                                    40ull << 30),
                                   ctx->ws_bcat.n);
    int64_t max_slots = std::max<int64_t>(slots_for(pl, 1), static_cast<int64_t>(budget / slot_bytes));
    max_slots = std::min<int64_t>(max_slots, 65535ll * (pl->mpad >= 64 ? pl->n_cg : 1));
    int64_t pb = pl->mpad >= 64 ? max_slots / pl->n_cg : max_slots * pl->pps;
    pb = std::max<int64_t>(1, std::min<int64_t>(pb, 16384));  // 16-bit per-unit counters
    pb = std::min(pb, num_perm);
    const int64_t n_batches = sb_ceil_div(num_perm, pb);
    pb = sb_ceil_div(num_perm, n_batches);  // equal batches
    ctx->ws_bcat.reserve(static_cast<size_t>(slots_for(pl, pb)) * slot_bytes);
    // the fix-up list (and, unless the caller supplies the packed array, the packed counters) live in context scratch
    ctx->ws_flag_ij.reserve(pl->flag_cap);
    ctx->ws_flag_p.reserve(pl->flag_cap);
    DevBuf<unsigned int> count_log;  // bucket counters of every batch, read back once at the end
    count_log.reserve(static_cast<size_t>(n_batches) * pl->n_cg);
    if (packed) {
        pl->cpk = packed;
    } else {
        ctx->ws_cpk.reserve(static_cast<size_t>(e->n) * e->m);
        SB_CUDA(cudaMemsetAsync(ctx->ws_cpk.p, 0, static_cast<size_t>(e->n) * e->m * sizeof(uint32_t), st));
        pl->cpk = ctx->ws_cpk.p;
    }
    pl->cpk_perms = 0;
    if (pl->op.any_inexact) enrich_transposed(e);

    delete tr_ws;
    int64_t flagged = 0, overflow_batches = 0, ktile_iters = 0;
    const int64_t tiles_per_pass = static_cast<int64_t>(pl->n_tiles) * pl->n_cg;  // includes padding tiles
    const unsigned int cap_cg = pl->flag_cap / static_cast<unsigned int>(pl->n_cg);
    auto batch_of = [&](int64_t b, const int32_t*& perm, int64_t& np, int& slots, int& q_total) {
        const int64_t p0 = b * pb;
        np = std::min(pb, num_perm - p0);
        perm = perm_dev + p0 * e->n;
        slots = static_cast<int>(slots_for(pl, np));
        q_total = pl->mpad >= 64 ? static_cast<int>(np) : slots;
    };
    auto gather_batch = [&](int64_t b, cudaStream_t on) {
        const int32_t* perm;
        int64_t np;
        int slots, q_total;
        batch_of(b, perm, np, slots, q_total);
        launch_gather(ctx, pl, pl->op, perm, slots, static_cast<int>(np), ctx->ws_bcat.p, on);
    };
    for (int64_t b = 0; b < n_batches; ++b) {
        const int32_t* perm;
        int64_t np;
        int slots, q_total;
        batch_of(b, perm, np, slots, q_total);
        {
            PhaseTrace tr(ctx, "tc.batch.gather");
            gather_batch(b, st);
        }
        PhaseTrace tr_b(ctx, "tc.batch.gemm+fixup");
        if (!packed && pl->cpk_perms + np > 60000) flush_counts(e, pl, cneg, cpos);  // 16-bit fields about to overflow
        pl->cpk_perms += np;
        SB_CUDA(cudaMemsetAsync(pl->flag_count.p, 0, pl->n_cg * sizeof(unsigned int), st));
        run_batch_gemm(e, pl, pl->op, TCM_COUNT | TCM_FLAG, 0, q_total, static_cast<int>(np), 0, pl->n_rb);
        ktile_iters += tiles_per_pass * q_total;
        if (pl->op.any_inexact) {
            SB_CUDA(cudaMemcpyAsync(count_log.p + b * pl->n_cg, pl->flag_count.p, pl->n_cg * sizeof(unsigned int),
                                    cudaMemcpyDeviceToDevice, st));
            fixup_flag_buckets(e, st, perm, ctx->ws_flag_ij.p, ctx->ws_flag_p.p, pl->flag_count.p, pl->n_cg, cap_cg,
                               cneg, cpos, packed);
        }
    }

    // ---- bucket counters of all batches: statistics, and the (rare) overflowed buckets, which the fix-up kernel skipped
    std::vector<unsigned int> h_log(static_cast<size_t>(n_batches) * pl->n_cg, 0u);
    if (pl->op.any_inexact) {
        SB_CUDA(cudaMemcpyAsync(h_log.data(), count_log.p, h_log.size() * sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
        SB_CUDA(cudaStreamSynchronize(st));
    }
    std::vector<unsigned int> h_flags(pl->n_cg);
    for (int64_t b = 0; b < n_batches; ++b) {
        std::vector<char> redo(pl->n_cg, 0);
        bool any = false;
        for (int cg = 0; cg < pl->n_cg; ++cg) {
            const unsigned int c = h_log[b * pl->n_cg + cg];
            if (c > cap_cg) {
                redo[cg] = 1;
                any = true;
            } else {
                flagged += c;
            }
        }
        if (!any) continue;
        // Overflowed buckets: re-emit the flags of this batch (without re-adding the decided counts) slot by slot and
        // in row-block ranges whose worst case -- every cell of the range flagged -- fits a bucket, and fix up the
        // overflowed buckets only (the others are already done).
        ++overflow_batches;
        const int32_t* perm;
        int64_t np;
        int slots, q_total;
        batch_of(b, perm, np, slots, q_total);
        gather_batch(b, st);
        const int rb_step = std::max<int>(1, static_cast<int>(cap_cg / (TC_PROWS * 64)));
        for (int q = 0; q < q_total; ++q) {
            const int bp = pl->mpad >= 64 ? 1
                                          : static_cast<int>(std::min<int64_t>(pl->pps, np - static_cast<int64_t>(q) * pl->pps));
            // the batch-local permutation index of a flag is relative to slot q: offset the index base instead
            const int32_t* perm_q = perm + static_cast<int64_t>(q) * pl->pps * e->n;
            for (int rb0 = 0; rb0 < pl->n_rb; rb0 += rb_step) {
                const int nrb = std::min(rb_step, pl->n_rb - rb0);
                SB_CUDA(cudaMemsetAsync(pl->flag_count.p, 0, pl->n_cg * sizeof(unsigned int), st));
                run_batch_gemm(e, pl, pl->op, TCM_FLAG, q, 1, bp, rb0, nrb);
                SB_CUDA(cudaMemcpyAsync(h_flags.data(), pl->flag_count.p, pl->n_cg * sizeof(unsigned int),
                                        cudaMemcpyDeviceToHost, st));
                SB_CUDA(cudaStreamSynchronize(st));
                for (int cg = 0; cg < pl->n_cg; ++cg) {
                    if (!redo[cg] || !h_flags[cg]) continue;
                    SB_CHECK(h_flags[cg] <= cap_cg, "internal error: flag list overflow in single-slot recovery");
                    fixup_flags(e, perm_q, ctx->ws_flag_ij.p + static_cast<size_t>(cg) * cap_cg,
                                ctx->ws_flag_p.p + static_cast<size_t>(cg) * cap_cg, h_flags[cg], cneg, cpos, packed);
                    flagged += h_flags[cg];
                }
            }
            ktile_iters += tiles_per_pass;
        }
    }
    if (!packed) flush_counts(e, pl, cneg, cpos);
    pl->cpk = nullptr;
    e->stats[0] = e->n * e->m * num_perm - flagged;
    e->stats[1] = flagged;
    e->stats[2] = pl->n_tiles_real;
    e->stats[3] = static_cast<int64_t>(pl->n_rb) * pl->n_kt;
    e->stats[4] = pl->op.D;
    e->stats[5] = ktile_iters;
    e->stats[6] = overflow_batches;
}

// ================================================================================================ z-score null
// neighborhood_score_type = 'z-score' (safe_extras.py:19-31) on the tensor cores.  Per permutation the score needs
// three contractions with the neighborhood matrix -- the sum of the neighbors' values, the sum of their squares (rounded
// like np.power(B, 2) rounds them) and the number of non-NaN neighbors -- and then M / std in fp64.  All three come out
// of ONE accumulation: a column group of the z plan is 32 attributes x 6 int8 digit planes (3 of the value, 2 of the
// square, 1 of the non-NaN indicator) = the same 192-byte records and 192-column MMAs as the 'sum' null, and the
// epilogue (k_gemm<3, TCK_Z>) holds the three exact fixed-point sums of a cell in registers:
//   * an fp32 filter with a 1e-4 margin settles all but ~1e-3 of the comparisons against the observed z-score;
//   * the rest is evaluated in fp64 (z_compare_fp64): by the exact engine's own formula where values and squares are
//     exactly representable (binary / integer / dyadic data: same bits, no fix-ups), else through a rigorous interval;
//   * comparisons whose interval contains the observed z-score (~1e-4) go to a list that k_zfix re-evaluates with the
//     exact engine's own accumulation (score_one order), one warp per entry.
// No sum ever reaches HBM (round-2 first version: three STORE GEMMs + a comparison kernel, 24 bytes of sums per
// comparison written and read back, 1.16 ms per C3 permutation).

// z records: digits[(cg * n + r) * 192 + d * 64 + h * 32 + s * 16 + a] = plane 2 d + s of attribute cg * 32 + h * 16 + a,
// i.e. the 32-column half h of every plane pair d -- what one epilogue warp reads -- carries all six planes of 16
// attributes.  Planes 0-2: balanced base-256 digits of rint(v 2^shift1), 3-4: of rint(v^2 2^shift2), 5: 1 where v is
// not NaN.
template <class T>
__global__ void k_quantize_z(const T* __restrict__ b, int64_t n, int64_t m, int32_t n_cg,
                             const int32_t* __restrict__ shift1, const int32_t* __restrict__ shift2,
                             int8_t* __restrict__ digits) {
    const int64_t total = static_cast<int64_t>(n_cg) * n * 32;
    int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    const int64_t step = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (; idx < total; idx += step) {
        const int a32 = static_cast<int>(idx & 31);
        const int64_t r = (idx >> 5) % n, cg = (idx >> 5) / n;
        const int64_t j = cg * 32 + a32;
        int q1 = 0, q2 = 0, valid = 0;
        if (j < m) {
            const T v = b[r * m + j];
            if (v == v) {
                valid = 1;
                q1 = static_cast<int>(rint(ldexp(static_cast<double>(v), shift1[j])));
                q2 = static_cast<int>(rint(ldexp(xf_value<XF_SQUARE, T>(v), shift2[j])));
            }
        }
        int8_t pl[6];
        int dig = static_cast<int>(static_cast<int8_t>(q1 & 0xff));
        pl[0] = static_cast<int8_t>(dig);
        q1 = (q1 - dig) >> 8;
        dig = static_cast<int>(static_cast<int8_t>(q1 & 0xff));
        pl[1] = static_cast<int8_t>(dig);
        pl[2] = static_cast<int8_t>((q1 - dig) >> 8);
        dig = static_cast<int>(static_cast<int8_t>(q2 & 0xff));
        pl[3] = static_cast<int8_t>(dig);
        pl[4] = static_cast<int8_t>((q2 - dig) >> 8);
        pl[5] = static_cast<int8_t>(valid);
        int8_t* const rec = digits + ((cg * n + r) * 192) + (a32 >> 4) * 32 + (a32 & 15);
#pragma unroll
        for (int k = 0; k < 6; ++k) rec[(k >> 1) * 64 + (k & 1) * 16] = pl[k];
    }
}

// per-attribute constants of the epilogue's fp32 filter: scales of the two fixed points and the error radii per
// non-NaN neighbor (half a unit, rounded up; zero for exactly representable columns)
__global__ void k_zcol(const int32_t* __restrict__ shift1, const int32_t* __restrict__ shift2,
                       const uint8_t* __restrict__ inex1, const uint8_t* __restrict__ inex2, int64_t m,
                       float4* __restrict__ zcol) {
    const int64_t j = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (j >= m) return;
    const double sc1 = ldexp(1.0, -shift1[j]), sc2 = ldexp(1.0, -shift2[j]);
    const double ra = inex1[j] ? 0.5 * sc1 * (1.0 + 0x1p-12) : 0.0, rb = inex2[j] ? 0.5 * sc2 * (1.0 + 0x1p-12) : 0.0;
    // (scales outside the fp32 range give 0 / inf: the filter then never accepts and fp64 decides)
    zcol[j] = make_float4(static_cast<float>(sc1), static_cast<float>(sc2), static_cast<float>(ra) * 1.000001f,
                          static_cast<float>(rb) * 1.000001f);
}


// observed z-scores into the plan's layout: z0t[j][r] = z0[node_of_row[r]][j]  (tiled transpose, one pass per null)
__global__ void __launch_bounds__(256) k_z_layout(const double* __restrict__ z0, const int32_t* __restrict__ node_of_row,
                                                  int64_t n, int64_t m, int64_t rows_pad, double* __restrict__ z0t) {
    __shared__ double tile[32][33];
    const int64_t r0 = static_cast<int64_t>(blockIdx.x) * 32, j0 = static_cast<int64_t>(blockIdx.y) * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int k = ty; k < 32; k += 8) {
        const int64_t r = r0 + k, j = j0 + tx;
        double v = __longlong_as_double(0x7FF8000000000000ll);
        if (r < n && j < m) v = z0[static_cast<int64_t>(node_of_row ? node_of_row[r] : r) * m + j];
        tile[k][tx] = v;
    }
    __syncthreads();
    for (int k = ty; k < 32; k += 8) {
        const int64_t j = j0 + k, r = r0 + tx;
        if (j < m && r < rows_pad) z0t[j * rows_pad + r] = tile[tx][k];
    }
}

// Exact re-evaluation of the undecided z-score comparisons.  One warp per entry: the lanes fetch 32 neighbors' values
// at a time (the dependent index -> permutation -> value loads are what a sequential walk spends its time on) and
// convert them, value and square, into a warp-private shared-memory line; then they are added IN ASCENDING NEIGHBOR
// ORDER -- the exact engine's accumulation order (score_one), so the result has the exact engine's bits -- at one
// broadcast load and two additions per neighbor.  A NaN neighbor contributes +0.0 to both sums (x + 0.0 == x except
// for the sign of a zero sum, which no comparison sees) and is left out of the count.
template <class T>
__global__ void __launch_bounds__(256) k_zfix(const int64_t* __restrict__ row_ptr, const int32_t* __restrict__ col_idx,
                                              const T* __restrict__ bt, const int32_t* __restrict__ perm, int64_t n,
                                              int64_t m, const double* __restrict__ z0,
                                              const uint64_t* __restrict__ flag_ij, const uint32_t* __restrict__ flag_p,
                                              unsigned int total, uint32_t* __restrict__ cneg,
                                              uint32_t* __restrict__ cpos, uint32_t* __restrict__ packed) {
    __shared__ double2 s_line[8][32];
    const int lane = threadIdx.x & 31;
    double2* const line = s_line[threadIdx.x >> 5];
    unsigned int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned int step = (gridDim.x * blockDim.x) >> 5;
    for (; k < total; k += step) {
        const uint64_t ij = flag_ij[k];
        const int64_t i = static_cast<int64_t>(ij >> 32), j = static_cast<int64_t>(ij & 0xffffffffu);
        const int32_t* pr = perm + static_cast<int64_t>(flag_p[k]) * n;
        const T* const col = bt + j * n;  // transposed copy: a flag's reads stay inside one column
        double sum = 0.0, sq = 0.0;
        int64_t cnt = 0;
        const int64_t e1 = row_ptr[i + 1];
        for (int64_t e0 = row_ptr[i]; e0 < e1; e0 += 32) {
            double2 mine = make_double2(0.0, 0.0);
            bool have = false;
            if (e0 + lane < e1) {
                const T v = col[pr[col_idx[e0 + lane]]];
                if (v == v) {
                    have = true;
                    mine = make_double2(static_cast<double>(v), sq_like_numpy<T>(v));
                }
            }
            cnt += __popc(__ballot_sync(0xffffffffu, have));
            __syncwarp();
            line[lane] = mine;
            __syncwarp();
            if (e1 - e0 >= 32) {
#pragma unroll
                for (int t = 0; t < 32; ++t) {
                    const double2 x = line[t];
                    sum += x.x;
                    sq += x.y;
                }
            } else {
                const int chunk = static_cast<int>(e1 - e0);
                for (int t = 0; t < chunk; ++t) {
                    const double2 x = line[t];
                    sum += x.x;
                    sq += x.y;
                }
            }
        }
        if (lane == 0) {
            const double z = zscore_from_sums(sum, sq, cnt);
            const double o = z0[i * m + j];
            if (packed) {
                const uint32_t inc = (z <= o ? 1u : 0u) + (z >= o ? 0x10000u : 0u);
                if (inc) atomicAdd(&packed[i * m + j], inc);
            } else {
                if (z <= o) atomicAdd(&cneg[i * m + j], 1u);
                if (z >= o) atomicAdd(&cpos[i * m + j], 1u);
            }
        }
    }
}

// the z plan: tiles of A as in the 'sum' plan, column groups of 32 attributes, z records, filter constants
static TcPlan* build_plan_z(sb_enrich* e, const TcPlan* main) {
    sb_ctx* ctx = e->ctx;
    cudaStream_t st = ctx->stream;
    TcPlan* pl = build_plan(e, /*zgroups=*/true);
    try {
        KernelTimer kt_prep(ctx, SB_K_PREP);
        PhaseTrace tr(ctx, "tc.plan.z_records");
        const int64_t n = e->n, m = e->m;
        TcOperand& op = pl->op;  // the z records stand in for the operand of the generic batch code
        op.built = true;
        op.D = 3;
        // squares: two digit planes; exactly representable columns (<= 14 magnitude bits) keep every bit
        std::vector<int32_t> h_kmax, h_lmin;
        std::vector<double> h_vmax;
        const int32_t flags = column_ranges(e, XF_SQUARE, h_kmax, h_lmin, h_vmax);
        if (flags & 1) {  // a square overflows to inf: exact SIMT engine
            pl->usable = false;
            return pl;
        }
        std::vector<int32_t> h_shift(m, 0);
        std::vector<uint8_t> h_inexact(m, 0);
        bool any2 = false;
        for (int64_t j = 0; j < m; ++j) {
            if (h_kmax[j] == INT_MIN) continue;
            const int bits = h_kmax[j] - h_lmin[j] + 1;
            if (bits <= 14) {
                h_shift[j] = -h_lmin[j];
            } else {
                h_shift[j] = inexact_shift(2, h_kmax[j], h_vmax[j], true);  // |q| <= 32639 (no differences in TCK_Z)
                h_inexact[j] = 1;
                any2 = true;
            }
        }
        pl->z_shift2.reserve(m);
        pl->z_inex2.reserve(m);
        SB_CUDA(cudaMemcpyAsync(pl->z_shift2.p, h_shift.data(), m * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        SB_CUDA(cudaMemcpyAsync(pl->z_inex2.p, h_inexact.data(), m, cudaMemcpyHostToDevice, st));
        // values: the 'sum' plan's fixed point (its shifts hold for three planes whatever D it chose: a column is
        // inexact only when three planes are not enough)
        pl->z_shift1 = main->op.shift.p;
        pl->z_inex1 = main->op.inexact.p;
        op.any_inexact = main->op.any_inexact || any2;
        pl->z_col.reserve(m);
        k_zcol<<<static_cast<unsigned>(sb_ceil_div(m, 256)), 256, 0, st>>>(pl->z_shift1, pl->z_shift2.p, pl->z_inex1,
                                                                          pl->z_inex2.p, m, pl->z_col.p);
        SB_LAUNCH_CHECK(ctx);
        op.digits.reserve(static_cast<size_t>(pl->n_cg) * n * 192);
        const unsigned qblocks = static_cast<unsigned>(
            std::min<int64_t>(sb_ceil_div(static_cast<int64_t>(pl->n_cg) * n * 32, 256), ctx->num_sms * 32));
        if (e->dtype == SB_F32)
            k_quantize_z<float><<<qblocks, 256, 0, st>>>(static_cast<const float*>(e->b), n, m, pl->n_cg, pl->z_shift1,
                                                         pl->z_shift2.p, op.digits.p);
        else
            k_quantize_z<double><<<qblocks, 256, 0, st>>>(static_cast<const double*>(e->b), n, m, pl->n_cg,
                                                          pl->z_shift1, pl->z_shift2.p, op.digits.p);
        SB_LAUNCH_CHECK(ctx);
        // undecided comparisons: one list (no buckets), a quarter of a batch's worst case at most
        if (op.any_inexact) {
            const int64_t padded_cells = static_cast<int64_t>(pl->n_rb) * TC_PROWS * 32 * pl->n_cg;
            int64_t cap = std::min<int64_t>(std::max<int64_t>(padded_cells / 4, 4ll << 20), 128ll << 20);
            if (getenv("SB_FLAG_CAP")) cap = atoll(getenv("SB_FLAG_CAP"));  // tests: force the overflow recovery
            pl->flag_cap = static_cast<unsigned int>(std::max<int64_t>(cap, static_cast<int64_t>(TC_PROWS) * 32 * pl->n_cg));
        } else {
            pl->flag_cap = 1;
        }
        SB_CUDA(cudaStreamSynchronize(st));  // host vectors above go out of scope
    } catch (...) {
        delete pl;
        throw;
    }
    return pl;
}

bool tc_perm_counts_z(sb_enrich* e, const int32_t* perm_dev, int64_t num_perm, uint32_t* cneg, uint32_t* cpos,
                      uint32_t* packed) {
    sb_ctx* ctx = e->ctx;
    cudaStream_t st = ctx->stream;
    PhaseTrace tr_all(ctx, "tc.perm_counts_z(total)");
    if (!e->tc) e->tc = build_plan(e);
    // +-inf / giant neighborhoods / fewer than 64 attributes: exact SIMT engine
    if (!e->tc->usable || e->m < 64 || e->m > 65535) return false;
    if (!e->tc->z) e->tc->z = build_plan_z(e, e->tc);
    TcPlan* pl = e->tc->z;
    if (!pl->usable) return false;
    const TcOperand& op = pl->op;
    SB_CHECK(!packed || num_perm < 65536, "packed counts hold fewer than 65536 permutations per call");
    const double* z0 = enrich_observed(e, SB_SCORE_ZSCORE);

    // observed z-scores in the plan's layout: [32 n_cg][rows_pad], internal row order (one pass per null)
    const int64_t rows_pad = static_cast<int64_t>(pl->n_rb) * TC_PROWS;
    DevBuf<double> z0t;
    z0t.reserve(static_cast<size_t>(rows_pad) * e->m);
    {
        dim3 grid(static_cast<unsigned>(sb_ceil_div(rows_pad, 32)), static_cast<unsigned>(sb_ceil_div(e->m, 32)));
        SB_CHECK(grid.y <= 65535, "z-score tensor path: too many attributes");
        k_z_layout<<<grid, 256, 0, st>>>(z0, pl->order, e->n, e->m, rows_pad, z0t.p);
        SB_LAUNCH_CHECK(ctx);
    }

    // batches as in tc_perm_counts: gathered tiles <= ~1/4 of the free memory seen at the first null (<= 40 GiB)
    const size_t tile_b = static_cast<size_t>(TC_KT) * 192;
    const size_t slot_bytes = static_cast<size_t>(pl->n_kt) * tile_b;
    if (ctx->bcat_budget == 0) {
        size_t free_b = 0, total_b = 0;
        SB_CUDA(cudaMemGetInfo(&free_b, &total_b));
        ctx->bcat_budget = std::max<size_t>(free_b / 4, 1);
    }
    const size_t budget = std::max(std::min<size_t>(std::max<size_t>(ctx->bcat_budget, slot_bytes * pl->n_cg), 40ull << 30),
                                   ctx->ws_bcat.n);
    int64_t pb = std::max<int64_t>(1, static_cast<int64_t>(budget / slot_bytes) / pl->n_cg);
    pb = std::min<int64_t>(std::min<int64_t>(pb, 65535), std::min<int64_t>(16384, num_perm));
    const int64_t n_batches = sb_ceil_div(num_perm, pb);
    pb = sb_ceil_div(num_perm, n_batches);  // equal batches
    ctx->ws_bcat.reserve(static_cast<size_t>(pb) * pl->n_cg * slot_bytes);
    ctx->ws_flag_ij.reserve(pl->flag_cap);
    ctx->ws_flag_p.reserve(pl->flag_cap);
    pl->flag_count.reserve(1);
    if (packed) {
        pl->cpk = packed;
    } else {
        ctx->ws_cpk.reserve(static_cast<size_t>(e->n) * e->m);
        SB_CUDA(cudaMemsetAsync(ctx->ws_cpk.p, 0, static_cast<size_t>(e->n) * e->m * sizeof(uint32_t), st));
        pl->cpk = ctx->ws_cpk.p;
    }
    pl->cpk_perms = 0;

    auto zgemm = [&](int mode, int q_first, int q_total, int rb0, int n_rb) {
        pl->z_z0t = z0t.p;
        run_batch_gemm(e, pl, op, mode | TCM_Z, q_first, q_total, q_total, rb0, n_rb);
    };
    const void* bt = op.any_inexact ? enrich_transposed(e) : nullptr;  // [m][n]
    auto zfix = [&](const int32_t* perm, unsigned int count) {
        if (!count) return;
        const unsigned blocks = static_cast<unsigned>(std::min<int64_t>(sb_ceil_div(count, 8), ctx->num_sms * 16));
        KernelTimer kt(ctx, SB_K_FIXUP);
        if (e->dtype == SB_F32)
            k_zfix<float><<<blocks, 256, 0, st>>>(e->row_ptr.p, e->col_idx.p, static_cast<const float*>(bt), perm, e->n,
                                                  e->m, z0, ctx->ws_flag_ij.p, ctx->ws_flag_p.p, count, cneg, cpos,
                                                  packed ? packed : pl->cpk);
        else
            k_zfix<double><<<blocks, 256, 0, st>>>(e->row_ptr.p, e->col_idx.p, static_cast<const double*>(bt), perm,
                                                   e->n, e->m, z0, ctx->ws_flag_ij.p, ctx->ws_flag_p.p, count, cneg,
                                                   cpos, packed ? packed : pl->cpk);
        SB_LAUNCH_CHECK(ctx);
    };

    int64_t flagged = 0, overflow_batches = 0, ktile_iters = 0;
    const int64_t tiles_per_pass = static_cast<int64_t>(pl->n_tiles) * pl->n_cg;
    for (int64_t b = 0; b < n_batches; ++b) {
        const int64_t p0 = b * pb, np = std::min(pb, num_perm - p0);
        const int32_t* perm = perm_dev + p0 * e->n;
        launch_gather(ctx, pl, op, perm, static_cast<int>(np * pl->n_cg), static_cast<int>(np), ctx->ws_bcat.p, st);
        if (!packed && pl->cpk_perms + np > 60000) flush_counts(e, pl, cneg, cpos);  // 16-bit fields about to overflow
        pl->cpk_perms += np;
        SB_CUDA(cudaMemsetAsync(pl->flag_count.p, 0, sizeof(unsigned int), st));
        zgemm(TCM_COUNT | TCM_FLAG, 0, static_cast<int>(np), 0, pl->n_rb);
        ktile_iters += tiles_per_pass * np;
        unsigned int h_count = 0;
        if (op.any_inexact) {
            SB_CUDA(cudaMemcpyAsync(&h_count, pl->flag_count.p, sizeof h_count, cudaMemcpyDeviceToHost, st));
            SB_CUDA(cudaStreamSynchronize(st));
        }
        if (h_count <= pl->flag_cap) {
            zfix(perm, h_count);
            flagged += h_count;
            continue;
        }
        // list overflow: the decided counts are in; re-emit the flags slot by slot in row-block ranges whose worst
        // case fits the list
        ++overflow_batches;
        const int rb_step = std::max<int>(1, static_cast<int>(pl->flag_cap / (static_cast<unsigned int>(TC_PROWS) * 32u * pl->n_cg)));
        for (int q = 0; q < np; ++q) {
            for (int rb0 = 0; rb0 < pl->n_rb; rb0 += rb_step) {
                SB_CUDA(cudaMemsetAsync(pl->flag_count.p, 0, sizeof(unsigned int), st));
                zgemm(TCM_FLAG, q, 1, rb0, std::min(rb_step, pl->n_rb - rb0));
                SB_CUDA(cudaMemcpyAsync(&h_count, pl->flag_count.p, sizeof h_count, cudaMemcpyDeviceToHost, st));
                SB_CUDA(cudaStreamSynchronize(st));
                SB_CHECK(h_count <= pl->flag_cap, "internal error: z-score flag list overflow in recovery");
                zfix(perm + static_cast<int64_t>(q) * e->n, h_count);  // flag_p holds slot 0 of this launch
                flagged += h_count;
            }
            ktile_iters += tiles_per_pass;
        }
    }
    if (!packed) flush_counts(e, pl, cneg, cpos);
    pl->cpk = nullptr;
    e->stats[0] = e->n * e->m * num_perm - flagged;
    e->stats[1] = flagged;
    e->stats[2] = pl->n_tiles_real;
    e->stats[3] = static_cast<int64_t>(pl->n_rb) * pl->n_kt;
    e->stats[4] = 3;
    e->stats[5] = ktile_iters;
    e->stats[6] = overflow_batches;
    return true;
}


}  // namespace sb

using namespace sb;

// ================================================================================================ self-test hook
extern "C" int sb_selftest_mma_i8(sb_ctx* ctx, int ncols, int ktiles, int variant, const int8_t* a_host,
                                  const int8_t* b_host, int32_t* d_host) {
    SB_API_BEGIN
    SB_CHECK(ctx && a_host && b_host && d_host, "sb_selftest_mma_i8: NULL argument");
    SB_CHECK(ncols == 64 || ncols == 128 || ncols == 192, "sb_selftest_mma_i8: ncols must be 64, 128 or 192");
    SB_CHECK(ktiles >= 1 && ktiles <= 4096, "sb_selftest_mma_i8: ktiles out of range");
    ctx->bind();
    const int D = ncols / 64;
    const int K = ktiles * TC_KT;
    const size_t tile_b = static_cast<size_t>(TC_KT) * ncols;
    // host-side tiling into the production layouts; A is a 0/1 matrix (any non-zero entry counts as 1) and is given
    // to BOTH CTAs of the pair (rows 0..127 and 128..255 of the unit), which must produce the same product
    const int kt_pad = static_cast<int>(sb_ceil_div(ktiles, TC_TPS) * TC_TPS);  // padding tiles of A are all zero
    std::vector<ulonglong2> at(static_cast<size_t>(kt_pad) * TC_PROWS, make_ulonglong2(0, 0));
    std::vector<int8_t> bt(static_cast<size_t>(ktiles) * tile_b);
    for (int kt = 0; kt < ktiles; ++kt)
        for (int r = 0; r < TC_ROWS; ++r) {
            uint64_t w = 0;
            for (int k = 0; k < TC_KT; ++k)
                if (a_host[static_cast<size_t>(r) * K + kt * TC_KT + k]) w |= 1ull << k;
            for (int rank = 0; rank < 2; ++rank)
                at[((static_cast<size_t>(kt / TC_TPS) * 2 + rank) * TC_TPS + kt % TC_TPS) * TC_ROWS + r] = split_mask(w);
        }
    // tile row k (the MMA's K position) holds operand row tc_kpos(k)
    const int hc = ncols / 32;  // 16-column chunks per half tile
    for (int kt = 0; kt < ktiles; ++kt)
        for (int k = 0; k < TC_KT; ++k)
            for (int c = 0; c < ncols; ++c) {
                const int nc = c >> 4;
                bt[static_cast<size_t>(kt) * tile_b + (nc / hc) * (tile_b / 2) + (k >> 3) * (hc * 128) + (nc % hc) * 128 +
                   (k & 7) * 16 + (c & 15)] = b_host[static_cast<size_t>(kt * TC_KT + tc_kpos(k)) * ncols + c];
            }
    std::vector<int32_t> ptr = {0, kt_pad}, kts(kt_pad);
    for (int i = 0; i < kt_pad; ++i) kts[i] = std::min(i, ktiles - 1);
    DevBuf<ulonglong2> d_a;
    DevBuf<int8_t> d_b;
    DevBuf<int32_t> d_ptr, d_kt, d_out;
    d_a.reserve(at.size());
    d_b.reserve(bt.size());
    d_ptr.reserve(2);
    d_kt.reserve(kt_pad);
    d_out.reserve(static_cast<size_t>(TC_PROWS) * ncols);
    cudaStream_t st = ctx->stream;
    SB_CUDA(cudaMemcpyAsync(d_a.p, at.data(), at.size() * sizeof(ulonglong2), cudaMemcpyHostToDevice, st));
    SB_CUDA(cudaMemcpyAsync(d_b.p, bt.data(), bt.size(), cudaMemcpyHostToDevice, st));
    SB_CUDA(cudaMemcpyAsync(d_ptr.p, ptr.data(), 2 * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    SB_CUDA(cudaMemcpyAsync(d_kt.p, kts.data(), kt_pad * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    SB_CUDA(cudaMemsetAsync(d_out.p, 0xff, static_cast<size_t>(TC_PROWS) * ncols * sizeof(int32_t), st));
    GemmParams gp{};
    gp.a_bits = d_a.p;
    gp.tile_ptr = d_ptr.p;
    gp.tile_kt = d_kt.p;
    gp.bcat = d_b.p;
    gp.n_kt = ktiles;
    gp.n_rb = 1;
    gp.n_cg = 1;
    gp.q_total = 1;
    gp.q_chunks = 1;
    gp.q_per = 1;
    gp.mode = TCM_RAW;
    gp.n = TC_PROWS;
    gp.m = 64;
    gp.mpad = 64;
    gp.pps = 1;
    gp.batch_perms = 1;
    gp.raw_out = d_out.p;
    gp.b_lbo = static_cast<uint32_t>(ncols) / 2 * 8;
    gp.b_sbo = 128;
    gp.q_wrap = INT_MAX;
    gp.b_kstep = 4 * (static_cast<uint32_t>(ncols) / 2) * 8;
    if (variant & 2) std::swap(gp.b_lbo, gp.b_sbo);  // deliberately wrong descriptor (negative control)
    launch_gemm_d(ctx, D, gp, 1);
    std::vector<int32_t> both(static_cast<size_t>(TC_PROWS) * ncols);
    SB_CUDA(cudaMemcpyAsync(both.data(), d_out.p, both.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaStreamSynchronize(st));
    const size_t half_elems = static_cast<size_t>(TC_ROWS) * ncols;
    for (size_t i = 0; i < half_elems; ++i) d_host[i] = both[i];
    if (!(variant & 2))
        SB_CHECK(memcmp(both.data(), both.data() + half_elems, half_elems * sizeof(int32_t)) == 0,
                 "sb_selftest_mma_i8: the two CTAs of the pair disagree");
    SB_API_END
}

// Streaming-rate probe: `grid` CTAs each run `slots` accumulations over the same `ktiles` L2-resident tiles
// through the production pipeline (no HBM traffic after the first touch).  Returns the device time in ms.
extern "C" int sb_selftest_mma_rate(sb_ctx* ctx, int ncols, int ktiles, int slots, int grid, int dbg,
                                    const uint32_t* desc_override /*NULL, or {b_lbo, b_sbo, b_kstep}*/,
                                    double* ms_out) {
    SB_API_BEGIN
    SB_CHECK(ctx && ms_out, "sb_selftest_mma_rate: NULL argument");
    SB_CHECK(ncols == 64 || ncols == 128 || ncols == 192, "sb_selftest_mma_rate: ncols must be 64, 128 or 192");
    SB_CHECK(ktiles >= 1 && ktiles <= 4096 && slots >= 1 && grid >= 1, "sb_selftest_mma_rate: bad sizes");
    ktiles = static_cast<int>(sb_ceil_div(ktiles, TC_TPS) * TC_TPS);
    ctx->bind();
    const int D = ncols / 64;
    const size_t tile_b = static_cast<size_t>(TC_KT) * ncols;
    DevBuf<ulonglong2> d_a;
    DevBuf<int8_t> d_b;
    DevBuf<int32_t> d_ptr, d_kt, d_out;
    d_a.reserve(static_cast<size_t>(ktiles) * TC_PROWS);
    d_b.reserve(static_cast<size_t>(ktiles) * tile_b);
    d_ptr.reserve(2);
    d_kt.reserve(ktiles);
    d_out.reserve(static_cast<size_t>(TC_PROWS) * ncols);
    cudaStream_t st = ctx->stream;
    std::vector<ulonglong2> ha(static_cast<size_t>(ktiles) * TC_PROWS, split_mask(0x5555555555555555ull));
    SB_CUDA(cudaMemsetAsync(d_b.p, 1, static_cast<size_t>(ktiles) * tile_b, st));
    if (const char* fill = getenv("SB_RATE_RANDOM")) {
        // operands with the statistics of a real null (random digits, masks filled to SB_RATE_RANDOM percent): the
        // tensor pipe's power draw, and with it the clock the chip holds under its cap, depends on the data
        const unsigned pct = static_cast<unsigned>(atoi(fill));
        uint64_t x = 0x9E3779B97F4A7C15ull;
        auto next = [&x]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
        for (auto& m : ha) {
            uint64_t w = 0;
            for (int b = 0; b < 64; ++b) w |= static_cast<uint64_t>(next() % 100 < pct) << b;
            m = split_mask(w);
        }
        std::vector<int8_t> hb(static_cast<size_t>(ktiles) * tile_b);
        for (auto& v : hb) v = static_cast<int8_t>(next() >> 24);
        SB_CUDA(cudaMemcpyAsync(d_b.p, hb.data(), hb.size(), cudaMemcpyHostToDevice, st));
    }
    SB_CUDA(cudaMemcpyAsync(d_a.p, ha.data(), ha.size() * sizeof(ulonglong2), cudaMemcpyHostToDevice, st));
    SB_CUDA(cudaStreamSynchronize(st));  // (host vectors)
    std::vector<int32_t> ptr = {0, ktiles}, kts(ktiles);
    for (int i = 0; i < ktiles; ++i) kts[i] = i;
    // all units share tiles [0, ktiles): one row block, the CTAs are spread over the q chunks
    SB_CUDA(cudaMemcpyAsync(d_ptr.p, ptr.data(), 2 * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    SB_CUDA(cudaMemcpyAsync(d_kt.p, kts.data(), ktiles * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    GemmParams gp{};
    gp.a_bits = d_a.p;
    gp.tile_ptr = d_ptr.p;
    gp.tile_kt = d_kt.p;
    gp.bcat = d_b.p;
    gp.n_kt = ktiles;
    gp.n_rb = 1;
    gp.n_cg = 1;
    gp.q_per = slots;
    gp.q_chunks = std::max(1, grid / 2);
    gp.q_total = slots * gp.q_chunks;
    gp.q_wrap = 1;
    gp.dbg = dbg;
    gp.mode = TCM_RAW;
    gp.n = TC_PROWS;
    gp.m = 64;
    gp.mpad = 64;
    gp.pps = 1;
    gp.batch_perms = 1;
    gp.raw_out = d_out.p;
    gp.b_lbo = static_cast<uint32_t>(ncols) / 2 * 8;
    gp.b_sbo = 128;
    gp.b_kstep = 4 * (static_cast<uint32_t>(ncols) / 2) * 8;
    grid = std::max(1, grid / 2);  // CTA pairs
    if (desc_override) {  // speed-only experiments with other descriptor fields (results are not checked)
        gp.b_lbo = desc_override[0];
        gp.b_sbo = desc_override[1];
        gp.b_kstep = desc_override[2];
    }
    cudaEvent_t e0, e1;
    SB_CUDA(cudaEventCreate(&e0));
    SB_CUDA(cudaEventCreate(&e1));
    DevBuf<long long> d_prof;
    d_prof.reserve(16);
    gp.prof = d_prof.p;
    launch_gemm_d(ctx, D, gp, grid);  // warm-up (also pulls the tiles into L2)
    SB_CUDA(cudaMemsetAsync(d_prof.p, 0, 16 * sizeof(long long), st));
    SB_CUDA(cudaEventRecord(e0, st));
    launch_gemm_d(ctx, D, gp, grid);
    SB_CUDA(cudaEventRecord(e1, st));
    SB_CUDA(cudaStreamSynchronize(st));
    if (getenv("SB_TRACE")) print_prof(ctx, d_prof.p, "rate probe");
    float ms = 0.f;
    SB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *ms_out = ms;
    SB_API_END
}
