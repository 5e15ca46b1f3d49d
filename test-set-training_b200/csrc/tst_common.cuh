// tst_common.cuh -- shared device-side definitions for libtst.so (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/tst.h"

#define TST_PURPOSE_BOOT 1u
#define TST_PURPOSE_FEAT 2u

#define KIND_INTERNAL 1u
#define KIND_PURE 2u
#define KIND_MIXED 3u
#define SMALL_MAX 32u  // nodes with <= 32 distinct rows are finished by one warp (k_grow_small)

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Random123; constants as CUDA/curand_philox4x32_x.h:88-91).  Counter-based:
// every draw is a pure function of (seed, tree id, node id, purpose), so trees are
// schedule-independent and identical on the CPU oracle.
__host__ __device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                        uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}

__host__ __device__ __forceinline__ uint32_t pick4(const uint4& v, uint32_t i) {
    return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w));
}

__device__ __forceinline__ uint32_t mulhi32(uint32_t a, uint32_t b) { return __umulhi(a, b); }

// The first m positions of a node's partial Fisher-Yates draw, by one warp: lane b evaluates Philox block b
// (4 positions per block) so the blocks run side by side, lane 0 walks the swaps in order.  s_perm[0..d) must
// hold the identity on entry.  Same draw as the serial loops elsewhere (oracle/hist_rf.c: draw_features).
__device__ __forceinline__ void warp_draw_primary(uint16_t* s_perm, int d, int m, uint32_t begin, uint32_t depth, uint32_t gid,
                                                  uint32_t k0, uint32_t k1, uint32_t lane) {
    for (int base = 0; base < m; base += 128) {
        const int lim = min(m, base + 128), nblk = (lim - base + 3) >> 2;
        uint4 rnd = make_uint4(0, 0, 0, 0);
        if ((int)lane < nblk) rnd = philox4x32_10(begin, depth, gid, TST_PURPOSE_FEAT | ((uint32_t)((base >> 2) + (int)lane) << 8), k0, k1);
        for (int i = base; i < lim; ++i) {
            const uint32_t v = __shfl_sync(0xffffffffu, pick4(rnd, (uint32_t)i & 3u), (i - base) >> 2);
            if (lane == 0) {
                const int j = i + (int)mulhi32(v, (uint32_t)(d - i));
                const uint16_t fj = s_perm[j];
                s_perm[j] = s_perm[i];
                s_perm[i] = fj;
            }
        }
    }
    __syncwarp();
}

// The same draw without the permutation array, for m <= DRAW_SPARSE_MAX: lane i < m owns position i.  Its swap partner is
// j_i = i + mulhi(v_i, d - i); the feature it ends up with is whatever sat at position j_i before step i, i.e. u_k of the LATEST
// earlier step k that wrote there (j_k == j_i), else j_i itself, where u_k (the value position k held before step k) is found by
// pointer jumping along "latest earlier step that wrote position k".  No shared-memory round trips and no serial walk by lane 0:
// ~115 warp instructions instead of ~230, and a dependency chain of one MATCH + log2(m) shuffles instead of m swaps.  Writes s_perm[0..m);
// returns j_i so that warp_draw_materialize can rebuild the whole permutation before a fallback draw continues it (rare).
#define DRAW_SPARSE_MAX 16
__device__ __forceinline__ uint32_t warp_draw_first(uint16_t* s_perm, int d, int m, uint32_t begin, uint32_t depth, uint32_t gid,
                                                    uint32_t k0, uint32_t k1, uint32_t lane) {
    if (m > DRAW_SPARSE_MAX) {
        for (int f = lane; f < d; f += 32) s_perm[f] = (uint16_t)f;
        __syncwarp();
        warp_draw_primary(s_perm, d, m, begin, depth, gid, k0, k1, lane);
        return 0u;
    }
    uint4 rnd = make_uint4(0, 0, 0, 0);
    if ((int)lane < ((m + 3) >> 2)) rnd = philox4x32_10(begin, depth, gid, TST_PURPOSE_FEAT | (lane << 8), k0, k1);
    const uint32_t src = lane >> 2;
    const uint32_t vx = __shfl_sync(0xffffffffu, rnd.x, src), vy = __shfl_sync(0xffffffffu, rnd.y, src);
    const uint32_t vz = __shfl_sync(0xffffffffu, rnd.z, src), vw = __shfl_sync(0xffffffffu, rnd.w, src);
    const uint32_t v = (lane & 2u) ? ((lane & 1u) ? vw : vz) : ((lane & 1u) ? vy : vx);
    const bool mine = (int)lane < m;
    const uint32_t j = mine ? lane + mulhi32(v, (uint32_t)d - lane) : 0xFFFFFFFFu;
    // latest earlier step that wrote my position (K) / my partner's position (Kj): ONE match over 32 keys -- lanes 0..15 present
    // their partner j, lanes 16..31 present the position (lane - 16) as a query; equal keys meet in the match mask
    uint32_t K = lane, Kj = 0xFFFFFFFFu;
    {
        const uint32_t q = lane - 16u;
        const uint32_t key = lane < 16u ? (mine ? j : 0x10000u + lane) : ((int)q < m ? q : 0x20000u + lane);
        const uint32_t mm = __match_any_sync(0xffffffffu, key) & 0xFFFFu;
        const uint32_t below = mm & ((1u << (lane & 15u)) - 1u);       // steps before mine (data lane) / before the queried position
        const uint32_t hb = below ? 31u - __clz(below) : 0xFFFFFFFFu;
        const uint32_t kq = __shfl_sync(0xffffffffu, hb, (lane + 16u) & 31u);  // data lane i reads query lane 16 + i
        if (lane < 16u) { Kj = hb; if (kq != 0xFFFFFFFFu) K = kq; }
    }
    uint32_t u = K;  // follow the chain to the step whose position nobody wrote before: it held its own index
    for (int s = 1; s < m; s <<= 1) u = __shfl_sync(0xffffffffu, u, u);
    const uint32_t uk = __shfl_sync(0xffffffffu, u, Kj & 31u);
    if (mine) s_perm[lane] = (uint16_t)(Kj != 0xFFFFFFFFu ? uk : j);
    __syncwarp();
    return j;
}

// s_perm[0..d) := the permutation after the m swaps of warp_draw_first (what warp_draw_primary would have left).
__device__ __forceinline__ void warp_draw_materialize(uint16_t* s_perm, int d, int m, uint32_t j, uint32_t lane) {
    if (m > DRAW_SPARSE_MAX) return;  // dealt on the array already
    for (int f = lane; f < d; f += 32) s_perm[f] = (uint16_t)f;
    __syncwarp();
    for (int i = 0; i < m; ++i) {
        const uint32_t ji = __shfl_sync(0xffffffffu, j, i);
        if (lane == 0) {
            const uint16_t fj = s_perm[ji];
            s_perm[ji] = s_perm[i];
            s_perm[i] = fj;
        }
    }
    __syncwarp();
}

// Warp argmax of (gain desc, key asc) with three REDUX instead of a 64-bit shuffle butterfly.  Valid gains are
// positive doubles (>= 2), whose bit patterns order like unsigned integers; key == 0xFFFFFFFF marks "no candidate".
__device__ __forceinline__ void warp_argmax_redux(double& gn, uint32_t& key) {
    const unsigned long long bits = key != 0xFFFFFFFFu ? (unsigned long long)__double_as_longlong(gn) : 0ull;
    const uint32_t hi = (uint32_t)(bits >> 32), mh = __reduce_max_sync(0xffffffffu, hi);
    const uint32_t lo = hi == mh ? (uint32_t)bits : 0u, ml = __reduce_max_sync(0xffffffffu, lo);
    const uint32_t k = (hi == mh && lo == ml) ? key : 0xFFFFFFFFu;
    key = __reduce_min_sync(0xffffffffu, k);
    gn = key != 0xFFFFFFFFu ? __longlong_as_double((long long)(((unsigned long long)mh << 32) | ml)) : -1.0;
}

// ---------------------------------------------------------------------------------------------
// Criterion 2 (squared error on a continuous target): the target is 16-bit fixed point and travels in the
// entry where the class does (y = multiplicity | yq << 8); a histogram cell is ONE 64-bit word, weight in the low 24 bits
// and the sum of w * yq in the upper 40 (a node weighs < 2^24, so neither field overflows) -- one 64-bit shared atomic per
// (row, feature), one 64-bit shuffle chain for the prefix scan, in every tier.
#define TST_REG 2u
#define REG_NBITS 24u
__device__ __forceinline__ unsigned long long reg_cell(uint32_t w, uint32_t yq) { return (unsigned long long)w | ((unsigned long long)(w * yq) << REG_NBITS); }
__device__ __forceinline__ uint32_t reg_N(unsigned long long P) { return (uint32_t)P & 0xFFFFFFu; }
__device__ __forceinline__ unsigned long long reg_S(unsigned long long P) { return P >> REG_NBITS; }

// ---------------------------------------------------------------------------------------------
// One in-bag row of one tree: x = row id, y = bootstrap multiplicity | class << 8.  Real entries have
// multiplicity >= 1, so y == 0 marks padding.  64-bit entries keep the class next to the row id
// (no label gather in the hot loops) and leave the full 32 bits to the row id.
typedef uint2 Entry;
// entry lists are read once per pass and written once per level: stream them (evict-first) so they do not
// push the binned table out of L2
__device__ __forceinline__ Entry ent_load(const Entry* p) { return __ldcs(p); }
__device__ __forceinline__ void ent_store(Entry* p, const Entry& e) { __stcs(p, e); }
__device__ __forceinline__ uint32_t ent_w(const Entry& e) { return e.y & 0xFFu; }
__device__ __forceinline__ uint32_t ent_cls(const Entry& e) { return e.y >> 8; }

struct Task {  // 32 B, one node waiting for a split search
    uint32_t tree;   // batch-local tree index
    uint32_t begin;  // segment [begin, end) of the tree's in-bag entry list
    uint32_t end;
    uint32_t depth;
    uint32_t slot;   // node slot: left child = slot + 1, right child = slot + 2 * (distinct rows going left)
    uint32_t pad0, pad1, pad2;
};

struct TreeDesc {  // one tree of the batch in flight
    uint32_t gid;      // RNG id = gid_base + part * n_trees + t
    uint32_t part;     // group * k + fold
    uint32_t gs;       // start of the group's rows in `sorted`
    uint32_t a;        // rows of the group sorted before the held-out fold
    uint32_t fsz;      // held-out fold size
    uint32_t n_train;  // bootstrap draws
    uint32_t fs;       // held-out rows are sorted[fs .. fs + fsz)
    uint32_t arena;    // node-arena slot (batch-local unless keep_forest)
    uint32_t n_inbag;  // distinct in-bag rows (written by k_boot_scan)
    uint32_t imp_row;  // row of the importance table: index of the tree among this call's trees
};

struct DevStats {
    unsigned long long row_visits_big, row_visits_mid, row_visits_small, nodes, leaves, searched, fallback, oof_steps, inbag_rows, row_visits_sub, row_visits_wide;
    unsigned long long prof[8];  // -DSUB_PROFILE builds: clock cycles of k_grow_sub (load, wide, phase 2 per CTA; narrow, small, idle per warp)
    uint32_t max_depth;
    uint32_t error;  // sticky: 1 = multiplicity overflow, 2 = queue overflow, 3 = bootstrap weight mismatch
};

struct RootSplit {
    uint32_t state;  // 0 = none, 1 = split found, 2 = leaf (stop rule), 3 = no valid split among the first m: generic path
    uint32_t key;    // feature << 8 | bin
    uint32_t N;
    uint32_t pad;
    uint32_t tot[TST_MAX_CLASSES];
    uint32_t bestL[TST_MAX_CLASSES];
};

struct GrowCtx {
    const uint8_t* bins_rm;  // [n][dstride]
    uint32_t dstride;
    int32_t d, C, m;
    int32_t max_depth;
    uint32_t mss, msl;
    uint32_t criterion;  // 0 = Gini; 1 = squared error on ordinal levels (labels are the levels of a quantised numeric target)
    uint32_t k0, k1;  // Philox key = seed
    Entry* entA;      // [trees][n_cap] ping
    Entry* entB;      // pong
    uint64_t n_cap;
    uint2* nodes;     // [arena slots][node_stride]
    uint64_t node_stride;
    uint32_t* leafc;  // [arena slots][leafc_stride]
    uint64_t leafc_stride;
    TreeDesc* trees;
    Task* qnext;      // big-node queue for the next level
    uint32_t* qnext_count;
    uint32_t qbig_cap;
    Task* qmid_next;  // warp-per-node queue for the next level (SMALL_MAX < rows <= mid_max)
    uint32_t* qmid_next_count;
    Task* qsub;       // CTA-per-subtree queue (k_grow_sub): SMALL_MAX < rows <= sub_max and weight < 2^16; filled by every level, drained once
    uint32_t* qsub_count;
    uint32_t sub_max; // 0 disables the tier
    uint32_t big_min; // a tree root with more distinct rows than this is a "big root" (dense root kernels + k_grow_big)
    Task* qwide_next; // four-warps-per-node queue for the next level (mid_max < rows <= wide_max, weight < 2^16): k_grow_wide
    uint32_t* qwide_next_count;
    uint32_t qwide_cap;
    uint32_t wide_max; // 0 disables the tier
    // size classes of the warp-per-node tier: nodes of SMALL_MAX < rows <= midc_max[0] and of midc_max[0] < rows <= midc_max[1] have their own
    // (dense) queues, served by instantiations of k_grow_mid4 with 2 and 4 row slots per lane (more resident warps); 0 = class off
    // CTA-per-node nodes that weigh < 2^16 (packed Gini) have their own queue when `bigl` is set: the LITE instantiation of k_grow_big
    Task* qbigl_next;
    uint32_t* qbigl_next_count;
    uint32_t bigl;
    Task* qmidc_next[2];
    uint32_t* qmidc_next_count[2];
    uint32_t midc_max[2];
    uint32_t mid_max; // == SMALL_MAX disables the mid tier
    uint32_t mid_packed; // k_grow_mid4: 16-bit packed class counts, nodes must weigh <= 65535
    Task* qsmall;
    uint32_t* qsmall_count;
    uint32_t qsmall_cap;
    void* qtiny;      // thread-per-subtree queue (16-byte TinyTask), rows <= tiny_max
    uint32_t* qtiny_count;
    uint32_t qtiny_cap;
    uint32_t tiny_max; // 0 disables the tier
    DevStats* stats;
    unsigned long long* imp;  // [call's trees][d] impurity decrease per feature in Q24 fixed point (nullptr: not collected)
    const RootSplit* rsplit;  // non-null: big roots were split by the dense root kernels (Task.pad0 == 1)
    uint32_t big_packed;  // k_grow_big: packed 16-bit histogram + scan for nodes that weigh < 2^16 (<= 4 classes, m <= 8)
    uint32_t stash_max_rows;  // nodes above this size gather instead: their stash would stream through HBM (the table itself is L2-resident)
    uint2* stash;             // non-null: [tree][n_cap] the m <= 8 sampled bins of each row, written by the histogram pass of k_grow_big for its partition pass
    int32_t big_tile; // k_grow_big stages whole rows in shared memory (dstride <= 128)
    int32_t HFS_max;  // words per feature reserved in k_grow_big's histogram region: max(FS, packed layout) -- 2-class tables have FS = 544 < 640
    int32_t MB;       // features per shared-histogram batch
    int32_t FS;       // histogram words per feature = 32 * (8 * C + 1)
};

__device__ __forceinline__ uint32_t hist_index(uint32_t bin, uint32_t cls, uint32_t C) {
    // 8 bins per lane, lane stride 8C+1 words (one pad word per 8 bins): the split scan reads without
    // bank conflicts.  (bin >> 3) * (8C + 1) + (bin & 7) * C + cls  ==  bin * C + (bin >> 3) + cls.
    return bin * C + (bin >> 3) + cls;
}

__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m;
    asm volatile("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

__device__ __forceinline__ double gini_gain(unsigned long long SL, uint32_t NL, unsigned long long SR, uint32_t NR) {
    // sklearn's Gini proxy (SK/tree/_criterion.pyx:147-163,661-687) minus a per-node constant:
    // two IEEE divisions and one add -- no multiply, so nothing can be contracted into an FMA.
    return __dadd_rn(__ddiv_rn(__ull2double_rn(SL), __uint2double_rn(NL)),
                     __ddiv_rn(__ull2double_rn(SR), __uint2double_rn(NR)));
}

// Squared-error proxy (SK/tree/_criterion.pyx MSE.proxy_impurity_improvement): S_L^2/N_L + S_R^2/N_R with
// S = sum over levels c of c * count_c; one multiply and one divide per side and one add, each rounded once
// (oracle/hist_rf.c: gain_mse).
__device__ __forceinline__ double mse_gain(unsigned long long SL, uint32_t NL, unsigned long long SR, uint32_t NR) {
    const double sl = __ull2double_rn(SL), sr = __ull2double_rn(SR);
    return __dadd_rn(__ddiv_rn(__dmul_rn(sl, sl), __uint2double_rn(NL)), __ddiv_rn(__dmul_rn(sr, sr), __uint2double_rn(NR)));
}
__device__ __forceinline__ float mse_gain_estimate(float SL, float NL, float SR, float NR) {
    return fmaf(SL, __fdividef(SL, NL), SR * __fdividef(SR, NR));  // no S^2 * N product: cannot overflow fp32
}

// fp32 ESTIMATE of the same quantity, used only to decide which candidates get the exact fp64 evaluation:
// (SL*NR + SR*NL) / (NL*NR) with one MUFU reciprocal.  All terms are positive, so the relative error is
// bounded by the sum of the roundings: 2 conversions + 3 products/fma (2^-24 each) + __fdividef (2 ulp)
// < 2^-21.  A candidate whose estimate is below TST_EST_MARGIN x (the largest estimate seen) cannot have
// the largest exact gain (margin 2^-18 = 8x the bound), so skipping it never changes the argmax or its
// tie-break; the winner itself is always evaluated exactly.  Results stay bit-identical to oracle/hist_rf.c.
#define TST_EST_MARGIN 0.99999619f  // 1 - 2^-18
__device__ __forceinline__ float gain_estimate(float SL, float NL, float SR, float NR) {
    return __fdividef(fmaf(SL, NR, SR * NL), NL * NR);
}

// ---------------------------------------------------------------------------------------------
// TMA (bulk async copy) + mbarrier helpers: 1-D global -> shared tiles of the in-bag entry lists are
// staged by the async proxy (SASS: UBLKCP) so the row-id stream never occupies LSU issue slots or
// registers and runs several tiles ahead of the gathers that depend on it.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// Mean decrease in impurity (SK/tree/_tree.pyx:1244-1280) accumulated while the tree grows: a split adds
// N*imp - N_L*imp_L - N_R*imp_R = gain - parent_term to its feature, with gain the proxy the split search maximised
// (S_L/N_L + S_R/N_R for Gini, S_L^2/N_L + S_R^2/N_R for squared error) and parent_term the same expression of the unsplit
// node (S/N resp. S^2/N).  Stored as floor(x * 2^24) in uint64: integer sums are independent of the order the nodes finish in.
__device__ __forceinline__ void imp_add(const GrowCtx& g, uint32_t imp_row, uint32_t f, double gain, double parent_term) {
    if (g.imp == nullptr) return;
    const double c = __dadd_rn(gain, -parent_term);
    if (c > 0.0) atomicAdd(&g.imp[(uint64_t)imp_row * (uint32_t)g.d + f], __double2ull_rd(__dmul_rn(c, 16777216.0)));
}
__device__ __forceinline__ double imp_parent_gini(unsigned long long S, uint32_t N) { return __ddiv_rn(__ull2double_rn(S), __uint2double_rn(N)); }
__device__ __forceinline__ double imp_parent_mse(unsigned long long S, uint32_t N) {
    const double s = __ull2double_rn(S);
    return __ddiv_rn(__dmul_rn(s, s), __uint2double_rn(N));
}

__device__ __forceinline__ void tiny_push(const GrowCtx& g, uint32_t tree, uint32_t begin, uint32_t rows, uint32_t depth, uint32_t slot, uint32_t buf) {
    // TinyTask = (begin, slot, depth | buf << 31, tree | rows << 16); buf = the entry buffer holding the rows (0 = entA, 1 = entB)
    uint32_t i = atomicAdd(g.qtiny_count, 1u);  // (the size histogram for the sort is taken by k_tiny_hist, not here)
    if (i < g.qtiny_cap) ((uint4*)g.qtiny)[i] = make_uint4(begin, slot, depth | (buf << 31), tree | (rows << 16));
    else g.stats->error = 2;
}

__device__ __forceinline__ void queue_push(Task* q, uint32_t* count, uint32_t cap, DevStats* st, const Task& t) {
    uint32_t i = atomicAdd(count, 1u);
    if (i < cap) q[i] = t; else st->error = 2;
}

// The queue of a node of `rows` > tiny_max distinct rows that weighs `weight` (an upper bound will do): warp per subtree, CTA per
// subtree (opt-in), a size class of the warp-per-node tier, four warps per node, CTA per node, or the warp-per-node tier.  The
// queue is SELECTED and pushed to once: the callers inline this, and eight copies of the 32-byte task store were 180 instructions.
__device__ __forceinline__ void route_push(const GrowCtx& g, uint32_t rows, uint32_t weight, const Task& t) {
    const bool light = weight <= 65535u;  // fits the packed 16-bit class counts
    Task* q = g.qnext;
    uint32_t* cnt = g.qnext_count;
    uint32_t cap = g.qbig_cap;
    if (rows <= SMALL_MAX) { q = g.qsmall; cnt = g.qsmall_count; cap = g.qsmall_cap; }
    else if (rows <= g.sub_max && light) { q = g.qsub; cnt = g.qsub_count; }
    else if (rows <= g.mid_max) {
        if (rows <= g.midc_max[0]) { q = g.qmidc_next[0]; cnt = g.qmidc_next_count[0]; }  // (<= 128 rows x 255 < 2^16: always packable)
        else if (rows <= g.midc_max[1]) { q = g.qmidc_next[1]; cnt = g.qmidc_next_count[1]; }
        else if (light || !g.mid_packed || g.criterion == TST_REG) { q = g.qmid_next; cnt = g.qmid_next_count; }
        // else: too heavy for 16-bit packed counts -> CTA per node
    } else if (rows <= g.wide_max && light) { q = g.qwide_next; cnt = g.qwide_next_count; cap = g.qwide_cap; }
    else if (light && g.bigl && t.pad0 == 0u) { q = g.qbigl_next; cnt = g.qbigl_next_count; }  // (a pre-split root keeps the generic kernel)
    queue_push(q, cnt, cap, g.stats, t);
}
