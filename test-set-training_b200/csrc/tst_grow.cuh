// tst_grow.cuh -- tree growth kernels (north_star items 2-5): per-node feature draw, class
// histogram, Gini split search, row partition, node records.
//
//   k_grow_big   one CTA per node with > 32 distinct rows: shared-memory histogram -> warp-shuffle
//                prefix scan -> fp64 gain -> block partition -> children.  HBM/L2-gather bound.
//   k_grow_small one warp per node with <= 32 rows: finishes the whole subtree in registers.
//
// Arithmetic restated from SK/tree/_splitter.pyx:262-504, _criterion.pyx:147-199,302-486,
// _partitioner.pyx:248-279, _tree.pyx:139-337 in the histogram form of
// SK/ensemble/_hist_gradient_boosting/histogram.pyx:336-480 / splitting.pyx:606-716;
// the normative spec is DESIGN.md section 2, the bit-exact checker is oracle/hist_rf.c.
#pragma once
#include "tst_common.cuh"

// ---------------------------------------------------------------------------------------------
// leaf / node record helpers
__device__ __forceinline__ void write_internal(uint2* nodes, uint32_t slot, uint32_t right_slot, uint32_t f, uint32_t b) {
    nodes[slot] = make_uint2(right_slot, (KIND_INTERNAL << 30) | (f << 8) | b);
}
__device__ __forceinline__ void write_pure_leaf(uint2* nodes, uint32_t slot, uint32_t N, uint32_t cls) {
    nodes[slot] = make_uint2(N, (KIND_PURE << 30) | cls);
}
__device__ __forceinline__ void write_mixed_leaf(uint2* nodes, uint32_t slot, uint32_t begin) {
    nodes[slot] = make_uint2(begin >> 1, (KIND_MIXED << 30));
}

__device__ __forceinline__ bool stop_rule(const GrowCtx& g, uint32_t N, uint32_t depth, bool pure) {
    // SK/tree/_tree.pyx:230-240 on bootstrap-weighted counts
    return pure || N < g.mss || N < 2u * g.msl || (g.max_depth > 0 && depth >= (uint32_t)g.max_depth);
}

// ---------------------------------------------------------------------------------------------
// One thread finishes a split: internal node record, leaf children written in place, other children
// routed by size to the CTA-per-node queue, the warp-per-node queue or the warp-per-subtree queue.
__device__ __forceinline__ void emit_split(const GrowCtx& g, const Task& t, uint2* nodes, uint32_t* leafc,
                                           const uint32_t* tot, const uint32_t* bestL, uint32_t N, uint32_t nLd,
                                           uint32_t bf, uint32_t bb, double gain /* the split's exact gain; < 0: recompute it from the counts */) {
    const int C = g.C;
    uint32_t NL = 0;
    for (int c = 0; c < C; ++c) NL += bestL[c];
    const uint32_t NR = N - NL;
    const uint32_t right_slot = t.slot + 2u * nLd;
    write_internal(nodes, t.slot, right_slot, bf, bb);
    unsigned long long new_leaves = 0;
    for (int side = 0; side < 2; ++side) {
        const uint32_t cb = side ? t.begin + nLd : t.begin, ce = side ? t.end : t.begin + nLd;
        const uint32_t cslot = side ? right_slot : t.slot + 1u;
        const uint32_t Nc = side ? NR : NL;
        int pc = -1;
        for (int c = 0; c < C; ++c) {
            uint32_t v = side ? tot[c] - bestL[c] : bestL[c];
            if (v == Nc) pc = c;
        }
        if (stop_rule(g, Nc, t.depth + 1u, pc >= 0)) {
            if (pc >= 0) write_pure_leaf(nodes, cslot, Nc, (uint32_t)pc);
            else {
                for (int c = 0; c < C; ++c) leafc[(uint64_t)(cb >> 1) * C + c] = side ? tot[c] - bestL[c] : bestL[c];
                write_mixed_leaf(nodes, cslot, cb);
            }
            new_leaves++;
        } else {
            Task ct;
            ct.tree = t.tree; ct.begin = cb; ct.end = ce; ct.depth = t.depth + 1u; ct.slot = cslot;
            ct.pad0 = ct.pad2 = 0;
            ct.pad1 = Nc;  // the node's weight: k_grow_big picks its packed path from it
            const uint32_t cr = ce - cb;
            if (cr <= g.tiny_max) tiny_push(g, t.tree, cb, cr, t.depth + 1u, cslot, (t.depth + 1u) & 1u);  // the children sit in the buffer of their depth
            else route_push(g, cr, Nc, ct);
        }
    }
    atomicAdd(&g.stats->nodes, 1ull + new_leaves);
    if (new_leaves) { atomicAdd(&g.stats->leaves, new_leaves); atomicMax(&g.stats->max_depth, t.depth + 1u); }
    atomicAdd(&g.stats->searched, 1ull);
    if (g.imp != nullptr) {  // the split's impurity decrease (the same doubles as the split search)
        unsigned long long SL = 0, SR = 0, S = 0;
        const bool mse = g.criterion == 1u;
        for (int c = 0; c < C; ++c) {
            const unsigned long long l = bestL[c], r = tot[c] - bestL[c], a = tot[c];
            SL += mse ? l * (unsigned)c : l * l; SR += mse ? r * (unsigned)c : r * r; S += mse ? a * (unsigned)c : a * a;
        }
        if (gain < 0.0) gain = mse ? mse_gain(SL, NL, SR, NR) : gini_gain(SL, NL, SR, NR);  // roots split by the dense root kernels
        imp_add(g, g.trees[t.tree].imp_row, bf, gain, mse ? imp_parent_mse(S, N) : imp_parent_gini(S, N));
    }
}

// ---------------------------------------------------------------------------------------------
// Criterion 2 (continuous squared error): leaves hold (N, S = sum of w * yq).  A leaf whose targets are all equal is a
// "pure" record (N, yq); any other leaf keeps N in the record and S in the leaf table (two words).
__device__ __forceinline__ void reg_write_leaf(uint2* nodes, uint32_t* leafc, uint32_t slot, uint32_t begin, uint32_t N, unsigned long long S,
                                               bool single_value, uint32_t yq) {
    if (single_value) nodes[slot] = make_uint2(N, (KIND_PURE << 30) | yq);
    else {
        leafc[(uint64_t)(begin >> 1) * 2] = (uint32_t)S;
        leafc[(uint64_t)(begin >> 1) * 2 + 1] = (uint32_t)(S >> 32);
        nodes[slot] = make_uint2(begin >> 1, (KIND_MIXED << 30) | N);
    }
}

// emit_split for criterion 2.  A child of one distinct row is a leaf (its target is trivially constant); whether a larger
// child is constant is found out when that child is processed (every tier starts a node with the min / max of its targets).
__device__ __forceinline__ void emit_split_reg(const GrowCtx& g, const Task& t, uint2* nodes, uint32_t* leafc, uint32_t N, unsigned long long S,
                                               uint32_t NL, unsigned long long SL, uint32_t nLd, uint32_t bf, uint32_t bb, double gain) {
    const uint32_t right_slot = t.slot + 2u * nLd;
    write_internal(nodes, t.slot, right_slot, bf, bb);
    unsigned long long new_leaves = 0;
    for (int side = 0; side < 2; ++side) {
        const uint32_t cb = side ? t.begin + nLd : t.begin, ce = side ? t.end : t.begin + nLd;
        const uint32_t cslot = side ? right_slot : t.slot + 1u;
        const uint32_t Nc = side ? N - NL : NL;
        const unsigned long long Sc = side ? S - SL : SL;
        const uint32_t cr = ce - cb;
        if (stop_rule(g, Nc, t.depth + 1u, cr == 1u)) {
            reg_write_leaf(nodes, leafc, cslot, cb, Nc, Sc, cr == 1u, cr == 1u ? (uint32_t)(Sc / Nc) : 0u);
            new_leaves++;
        } else {
            Task ct;
            ct.tree = t.tree; ct.begin = cb; ct.end = ce; ct.depth = t.depth + 1u; ct.slot = cslot;
            ct.pad0 = ct.pad2 = 0;
            ct.pad1 = Nc;
            if (cr <= g.tiny_max) tiny_push(g, t.tree, cb, cr, t.depth + 1u, cslot, (t.depth + 1u) & 1u);
            else route_push(g, cr, Nc, ct);
        }
    }
    atomicAdd(&g.stats->nodes, 1ull + new_leaves);
    if (new_leaves) { atomicAdd(&g.stats->leaves, new_leaves); atomicMax(&g.stats->max_depth, t.depth + 1u); }
    atomicAdd(&g.stats->searched, 1ull);
    imp_add(g, g.trees[t.tree].imp_row, bf, gain, imp_parent_mse(S, N));
}

// Split scan of one feature for criterion 2: cells are 64-bit (weight | sum << 24), lane l owns bins [8l, 8l+8) of the
// packed layout (80-byte lane segments).  Same structure as scan_feature_packed4: one 64-bit shuffle chain, fp32 estimate,
// fp64 gain S_L^2/N_L + S_R^2/N_R of the contenders, ties -> lowest bin.
__device__ __forceinline__ void scan_feature_reg(const uint32_t* __restrict__ hf, uint32_t msl, uint32_t lane, double& best_gain, uint32_t& best_bin,
                                                 unsigned long long& bestP, unsigned long long& tot_out) {
    const ulonglong2* hl = (const ulonglong2*)(hf + lane * 20u);
    unsigned long long p[8], run = 0;
    uint32_t occ = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const ulonglong2 hq = hl[q >> 1];
        const unsigned long long h = (q & 1) ? hq.y : hq.x;
        occ |= (h ? 1u : 0u) << q;
        run += h;
        p[q] = run;
    }
    unsigned long long incl = run;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const unsigned long long a = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= (uint32_t)off) incl += a;
    }
    const unsigned long long excl = incl - run, tot = __shfl_sync(0xffffffffu, incl, 31);
    tot_out = tot;
    const uint32_t N = reg_N(tot);
    const unsigned long long S = reg_S(tot);
    float ge[8], lmax = 0.0f;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        ge[q] = -1.0f;
        if ((occ >> q) & 1u) {
            const unsigned long long P = excl + p[q];
            const uint32_t NL = reg_N(P), NR = N - NL;
            if (NL >= msl && NR >= msl) {
                const unsigned long long SL = reg_S(P);
                ge[q] = mse_gain_estimate(__ull2float_rn(SL), __uint2float_rn(NL), __ull2float_rn(S - SL), __uint2float_rn(NR));
                lmax = fmaxf(lmax, ge[q]);
            }
        }
    }
    const float thr = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(lmax))) * TST_EST_MARGIN;
    double bg = -1.0;
    uint32_t bb = 0xFFFFFFFFu;
    unsigned long long bP = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        if (ge[q] >= thr) {
            const unsigned long long P = excl + p[q];
            const uint32_t NL = reg_N(P);
            const double gn = mse_gain(reg_S(P), NL, S - reg_S(P), N - NL);
            if (gn > bg) { bg = gn; bb = lane * 8u + q; bP = P; }
        }
    }
    double wg = bg;
    uint32_t wb = bb;
    warp_argmax_redux(wg, wb);
    if (wb != 0xFFFFFFFFu) {
        const uint32_t winner = __ffs(__ballot_sync(0xffffffffu, bb == wb && bg == wg)) - 1;
        bestP = __shfl_sync(0xffffffffu, bP, winner);
    }
    best_gain = wg;
    best_bin = wb;
}

// ---------------------------------------------------------------------------------------------
// Split scan of one feature histogram by one warp.  Lane l owns bins [8l, 8l+8).  Per class: local
// inclusive prefix, warp-shuffle exclusive scan of the lane totals, then L/R sums of squares are
// accumulated exactly in uint64.  Returns the warp-wide best (gain, bin) with ties -> lowest bin.
template <bool mse>  // mse: sL / sR hold sum_c c * count_c instead of sum_c count_c^2 (squared-error criterion on ordinal levels)
__device__ __forceinline__ void scan_feature_t(const uint32_t* __restrict__ hf, int C, uint32_t msl, uint32_t lane,
                                               double& best_gain, uint32_t& best_bin, uint32_t* tot_out, uint32_t* N_out) {
    unsigned long long sL[8], sR[8];
    uint32_t nL[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { sL[i] = 0; sR[i] = 0; nL[i] = 0; }
    uint32_t occ = 0, N = 0;
    const uint32_t* hl = hf + lane * (8u * C + 1u);
    for (int c = 0; c < C; ++c) {
        uint32_t p[8], run = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            uint32_t h = hl[i * C + c];
            occ |= (h ? 1u : 0u) << i;
            run += h;
            p[i] = run;
        }
        uint32_t incl = run;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= (uint32_t)off) incl += t;
        }
        uint32_t excl = incl - run;
        uint32_t tot = __shfl_sync(0xffffffffu, incl, 31);
        N += tot;
        if (tot_out && lane == 0) tot_out[c] = tot;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            uint32_t L = excl + p[i], R = tot - L;
            sL[i] += (unsigned long long)L * (mse ? (uint32_t)c : L);
            sR[i] += (unsigned long long)R * (mse ? (uint32_t)c : R);
            nL[i] += L;
        }
    }
    if (N_out && lane == 0) *N_out = N;
    // fp32 estimates first; only the candidates within TST_EST_MARGIN of the warp's best estimate get the fp64 gain
    float ge[8], lmax = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        ge[i] = -1.0f;
        if ((occ >> i) & 1u) {  // an empty bin repeats the previous partition and would lose the tie anyway
            uint32_t NL = nL[i], NR = N - NL;
            if (NL >= msl && NR >= msl) {
                ge[i] = mse ? mse_gain_estimate(__ull2float_rn(sL[i]), __uint2float_rn(NL), __ull2float_rn(sR[i]), __uint2float_rn(NR))
                            : gain_estimate(__ull2float_rn(sL[i]), __uint2float_rn(NL), __ull2float_rn(sR[i]), __uint2float_rn(NR));
                lmax = fmaxf(lmax, ge[i]);
            }
        }
    }
    const float thr = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(lmax))) * TST_EST_MARGIN;
    double bg = -1.0;
    uint32_t bb = 0xFFFFFFFFu;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        if (ge[i] >= thr) {
            double gn = mse ? mse_gain(sL[i], nL[i], sR[i], N - nL[i]) : gini_gain(sL[i], nL[i], sR[i], N - nL[i]);
            if (gn > bg) { bg = gn; bb = lane * 8u + i; }
        }
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        double og = __shfl_xor_sync(0xffffffffu, bg, off);
        uint32_t ob = __shfl_xor_sync(0xffffffffu, bb, off);
        if (og > bg || (og == bg && ob < bb)) { bg = og; bb = ob; }
    }
    best_gain = bg;
    best_bin = bb;
}

__device__ __forceinline__ void scan_feature(const uint32_t* __restrict__ hf, int C, uint32_t msl, uint32_t lane,
                                             double& best_gain, uint32_t& best_bin, uint32_t* tot_out, uint32_t* N_out, uint32_t criterion) {
    if (criterion == 1u) scan_feature_t<true>(hf, C, msl, lane, best_gain, best_bin, tot_out, N_out);
    else scan_feature_t<false>(hf, C, msl, lane, best_gain, best_bin, tot_out, N_out);
}

// ---------------------------------------------------------------------------------------------
// The same scan for a node that weighs < 2^16 with <= 4 classes, on a PACKED histogram: a cell is two words holding
// four 16-bit class counts, lane l owns bins [8l, 8l+8) in an 80-byte segment (conflict-free 128-bit reads), one 64-bit
// shuffle chain scans all classes, sums of squares fit 32 bits.  ~2.4x fewer instructions than scan_feature_t; the
// small nodes of the CTA tier (levels 8+ of a 1M-row tree) spend 40 % of their instructions in the scan.
#define BIGP_FS 640u  // words per feature of the packed histogram: 32 lanes x (8 bins x 2 words + 4 pad)
__device__ __forceinline__ void scan_feature_packed4(const uint32_t* __restrict__ hf, uint32_t msl, uint32_t lane, double& best_gain,
                                                     uint32_t& best_bin, uint32_t& bestP0, uint32_t& bestP1, uint32_t& T0, uint32_t& T1) {
    const uint4* hl = (const uint4*)(hf + lane * 20u);
    uint32_t p0[8], p1[8], run0 = 0, run1 = 0, occ = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const uint4 hq = hl[q >> 1];
        const uint32_t h0 = (q & 1) ? hq.z : hq.x, h1 = (q & 1) ? hq.w : hq.y;
        occ |= ((h0 | h1) ? 1u : 0u) << q;
        run0 += h0; run1 += h1;
        p0[q] = run0; p1[q] = run1;
    }
    const unsigned long long run = (unsigned long long)run0 | ((unsigned long long)run1 << 32);
    unsigned long long incl = run;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const unsigned long long a = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= (uint32_t)off) incl += a;
    }
    const unsigned long long excl = incl - run, tot = __shfl_sync(0xffffffffu, incl, 31);
    const uint32_t e0 = (uint32_t)excl, e1 = (uint32_t)(excl >> 32);
    T0 = (uint32_t)tot; T1 = (uint32_t)(tot >> 32);
    const uint32_t t0 = T0 & 0xFFFFu, t1 = T0 >> 16, t2 = T1 & 0xFFFFu, t3 = T1 >> 16, N = t0 + t1 + t2 + t3;
    float ge[8], lmax = 0.0f;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        ge[q] = -1.0f;
        if ((occ >> q) & 1u) {
            const uint32_t P0 = e0 + p0[q], P1 = e1 + p1[q];
            const uint32_t L0 = P0 & 0xFFFFu, L1 = P0 >> 16, L2 = P1 & 0xFFFFu, L3 = P1 >> 16;
            const uint32_t NL = L0 + L1 + L2 + L3, NR = N - NL;
            if (NL >= msl && NR >= msl) {
                const uint32_t R0 = t0 - L0, R1 = t1 - L1, R2 = t2 - L2, R3 = t3 - L3;
                const uint32_t SL = L0 * L0 + L1 * L1 + L2 * L2 + L3 * L3, SR = R0 * R0 + R1 * R1 + R2 * R2 + R3 * R3;
                ge[q] = gain_estimate(__uint2float_rn(SL), __uint2float_rn(NL), __uint2float_rn(SR), __uint2float_rn(NR));
                lmax = fmaxf(lmax, ge[q]);
            }
        }
    }
    const float thr = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(lmax))) * TST_EST_MARGIN;
    // exact fp64 gain of the contenders: usually ONE bin of one lane.  A loop over each lane's contender mask (lowest bin first)
    // with the bin's prefix picked by a select chain: one copy of the division code instead of eight (the kernels using this scan
    // were stalling on instruction fetch), typically one trip.
    uint32_t cm = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) cm |= (ge[q] >= thr ? 1u : 0u) << q;
    double bg = -1.0;
    uint32_t bb = 0xFFFFFFFFu, bP0 = 0, bP1 = 0;
#pragma unroll 1
    while (__any_sync(0xffffffffu, cm != 0u)) {
        if (cm) {
            const uint32_t q = (uint32_t)__ffs(cm) - 1u;
            cm &= cm - 1u;
            uint32_t a0 = p0[0], a1 = p1[0];
#pragma unroll
            for (int k = 1; k < 8; ++k) if (q == (uint32_t)k) { a0 = p0[k]; a1 = p1[k]; }
            const uint32_t P0 = e0 + a0, P1 = e1 + a1;
            const uint32_t L0 = P0 & 0xFFFFu, L1 = P0 >> 16, L2 = P1 & 0xFFFFu, L3 = P1 >> 16;
            const uint32_t NL = L0 + L1 + L2 + L3, NR = N - NL;
            const uint32_t R0 = t0 - L0, R1 = t1 - L1, R2 = t2 - L2, R3 = t3 - L3;
            const uint32_t SL = L0 * L0 + L1 * L1 + L2 * L2 + L3 * L3, SR = R0 * R0 + R1 * R1 + R2 * R2 + R3 * R3;
            const double gn = gini_gain(SL, NL, SR, NR);
            if (gn > bg) { bg = gn; bb = lane * 8u + q; bP0 = P0; bP1 = P1; }  // ascending bins: strict > keeps the lowest
        }
    }
    double wg = bg;
    uint32_t wb = bb;
    warp_argmax_redux(wg, wb);
    if (wb != 0xFFFFFFFFu) {
        const uint32_t winner = __ffs(__ballot_sync(0xffffffffu, bb == wb && bg == wg)) - 1;
        bestP0 = __shfl_sync(0xffffffffu, bP0, winner);
        bestP1 = __shfl_sync(0xffffffffu, bP1, winner);
    }
    best_gain = wg;
    best_bin = wb;
}

// ---------------------------------------------------------------------------------------------
// Warp-cooperative fetch of 32 binned rows (row-major table, `dstride` bytes each) into a shared tile.
// Lane l owns row l of the tile; 16-byte pieces of a row are loaded by neighbouring lanes in ONE
// instruction, so a 64-byte row costs one L1 wavefront and two L2 sectors instead of one sector per
// sampled feature.  tile pitch = dstride/4 + 1 words: the per-row byte reads that follow are conflict-free.
__device__ __forceinline__ void fetch_rows32(const uint8_t* __restrict__ bins, uint32_t dstride, uint32_t row, bool has,
                                             uint32_t* tile /* this warp's 32 x pitch words */, uint32_t lane) {
    // dstride is a power of two here (16 / 32 / 64 / 128): P pieces of 16 bytes per row
    const uint32_t P = dstride >> 4, lp = 31u - __clz(P), pitch = (dstride >> 2) + 1u;
    for (uint32_t i0 = 0; i0 < 32u * P; i0 += 128u) {
        uint4 v[4];
        uint32_t rs[4], pt[4];
        bool ok[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint32_t idx = i0 + q * 32u + lane;
            rs[q] = idx >> lp; pt[q] = idx & (P - 1u);
            // 64-byte rows (P = 4, pitch 17 words): one instruction covers rows {4q..4q+3} and {16+4q..16+4q+3} instead of
            // 8 consecutive rows, so the 32 (row, piece) word addresses of every store below fall into 32 different banks
            // (consecutive rows collide pairwise: 2 wavefronts per store)
            if (P == 4u) rs[q] = 4u * (uint32_t)q + ((lane >> 2) & 3u) + ((lane >> 4) << 4);
            uint32_t rid = __shfl_sync(0xffffffffu, row, rs[q] & 31u);
            ok[q] = __shfl_sync(0xffffffffu, has ? 1u : 0u, rs[q] & 31u) && idx < 32u * P;
            if (ok[q]) v[q] = __ldg((const uint4*)(bins + (uint64_t)rid * dstride) + pt[q]);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (ok[q]) {
                uint32_t* dst = tile + rs[q] * pitch + pt[q] * 4u;
                dst[0] = v[q].x; dst[1] = v[q].y; dst[2] = v[q].z; dst[3] = v[q].w;
            }
    }
    __syncwarp();
}

#ifndef BIG_MINB
#define BIG_MINB 3
#endif
// ---------------------------------------------------------------------------------------------
#ifndef BIGL_MINB
#define BIGL_MINB 4
#endif
// LITE: the packed-Gini nodes only (<= 4 classes, weight < 2^16, row tile, one batch of <= 8 features, no stash, no pre-split
// roots), from their own queue: every generic branch folds away, which is what lets the instantiation fit more CTAs per SM.
template <int THREADS, int CRIT = 0, bool LITE = false>  // CRIT: 0 Gini, 1 squared error on ordinal levels, 2 squared error on the 16-bit target (own instantiations: the Gini one keeps its registers)
__global__ void __launch_bounds__(THREADS, LITE ? BIGL_MINB : BIG_MINB) k_grow_big(GrowCtx g, const Task* __restrict__ tasks, uint32_t ntasks, const uint32_t* __restrict__ ntasks_dev) {
    constexpr int NW = THREADS / 32;
    constexpr bool MSE = CRIT == 1, REG = CRIT == 2;
    __shared__ unsigned long long s_fP[32], s_totP, s_bestP;  // criterion 2: left cell of each feature's best split, node total, best split's left cell
    __shared__ uint32_t s_ymin, s_ymax;
    extern __shared__ __align__(16) uint32_t s_hist[];  // [MB][FS] histogram, then (row tile) [THREADS][dstride/4 + 1]
    __shared__ uint32_t s_fL[8][4];  // packed nodes: left class counts of each feature's best split
    __shared__ uint16_t s_perm[TST_MAX_FEATURES];
    __shared__ uint32_t s_tot[TST_MAX_CLASSES];
    __shared__ uint32_t s_bestL[TST_MAX_CLASSES];
    __shared__ double s_fgain[32];
    __shared__ uint32_t s_fbin[32];
    __shared__ uint32_t s_feat[32];
    __shared__ uint32_t s_wl[2][4][NW];
    __shared__ double s_bgain;
    __shared__ uint32_t s_bkey, s_found, s_newj, s_N, s_stop, s_sj;  // s_sj: position of the best split's feature in the stash (or none)

    // thin levels are launched without a host round trip: the grid is an upper bound, the real count is on the device
    if (ntasks_dev) ntasks = min(ntasks, *ntasks_dev);
    if (blockIdx.x >= ntasks) return;
    const Task t = tasks[blockIdx.x];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    const int C = g.C, d = g.d, m = g.m;
    const TreeDesc td = g.trees[t.tree];
    const Entry* __restrict__ ent_in = ((t.depth & 1u) ? g.entB : g.entA) + (uint64_t)t.tree * g.n_cap;
    Entry* __restrict__ ent_out = ((t.depth & 1u) ? g.entA : g.entB) + (uint64_t)t.tree * g.n_cap;
    uint2* nodes = g.nodes + (uint64_t)td.arena * g.node_stride;
    uint32_t* leafc = g.leafc + (uint64_t)td.arena * g.leafc_stride;
    const uint8_t* __restrict__ bins = g.bins_rm;
    const uint32_t dstride = g.dstride;
    const bool use_tile = LITE || g.big_tile != 0;
    // a node that weighs < 2^16 (its parent wrote the weight into the task) uses the packed histogram and scan
    const bool packed = LITE || REG || (!MSE && g.big_packed != 0u && use_tile && C <= 4 && g.MB <= 8 && t.pad1 != 0u && t.pad1 <= 65535u);
    const uint32_t HFS = packed ? BIGP_FS : (uint32_t)g.FS;  // histogram words per feature
    const uint32_t pitch = (dstride >> 2) + 1u;
    uint32_t* s_tile = s_hist + (LITE ? 8u * BIGP_FS : (uint32_t)(g.MB * g.HFS_max)) + wid * 32u * pitch;  // this warp's rows (the histogram region holds either layout)

    for (int f = tid; f < d; f += THREADS) s_perm[f] = (uint16_t)f;
    if (tid == 0) { s_found = 0; s_bgain = -1.0; s_bkey = 0xFFFFFFFFu; s_stop = 0; s_sj = 0xFFFFFFFFu; s_ymin = 0xFFFFFFFFu; s_ymax = 0u; s_totP = 0ull; }
    __syncthreads();

    // roots already split by the dense column-major kernels (tst_root.cuh) only need the partition below
    const bool presplit = !LITE && t.pad0 == 1u && g.rsplit != nullptr && (g.rsplit[t.tree].state == 1u || g.rsplit[t.tree].state == 2u);
    if (presplit) {
        if (tid == 0) {
            const RootSplit& R = g.rsplit[t.tree];
            s_N = R.N; s_bkey = R.key; s_found = R.state == 1u; s_stop = R.state == 2u;
            for (int c = 0; c < C; ++c) { s_tot[c] = R.tot[c]; s_bestL[c] = R.bestL[c]; }
        }
        __syncthreads();
    }
    int pos = presplit ? d : 0;
    while (pos < d) {
        if (pos >= m && s_found) break;
        const bool primary = pos < m;
        const int nb = min(LITE ? 8 : g.MB, (primary ? m : d) - pos);
        // ---- per-node feature draw: partial Fisher-Yates keyed by (begin, depth, tree)
        if (tid == 0) {
            uint4 rnd = make_uint4(0, 0, 0, 0);
            for (int i = pos; i < pos + nb; ++i) {
                if (i == pos || (i & 3) == 0)
                    rnd = philox4x32_10(t.begin, t.depth, td.gid, TST_PURPOSE_FEAT | ((uint32_t)(i >> 2) << 8), g.k0, g.k1);
                int j = i + (int)mulhi32(pick4(rnd, i & 3), (uint32_t)(d - i));
                uint16_t fj = s_perm[j];
                s_perm[j] = s_perm[i];
                s_perm[i] = fj;
                s_feat[i - pos] = fj;
            }
            s_newj = 0xFFFFFFFFu;
            if (pos == m) atomicAdd(&g.stats->fallback, 1ull);
        }
        for (int i = tid; i < nb * (int)HFS; i += THREADS) s_hist[i] = 0;
        __syncthreads();

        // ---- class histogram: hist[j][bin][class] += multiplicity (SK/.../histogram.pyx:336-425 pattern)
        if (use_tile && nb <= 8) {
            uint2* stash_w = (!LITE && g.stash != nullptr && pos == 0 && t.end - t.begin <= g.stash_max_rows) ? g.stash + (uint64_t)t.tree * g.n_cap : nullptr;
            uint32_t featr[8];
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) featr[jj] = jj < nb ? s_feat[jj] : 0u;
            const uint8_t* trow = (const uint8_t*)(s_tile + lane * pitch);
            // the entry of the NEXT tile is requested before this tile's rows: the two dependent latencies
            // (entry list from HBM, then the row gather from L2) overlap across iterations
            uint32_t ymn = 0xFFFFFFFFu, ymx = 0u;
            Entry e_next = t.begin + tid < t.end ? ent_load(ent_in + t.begin + tid) : make_uint2(0u, 0u);
            for (uint32_t base = t.begin; base < t.end; base += THREADS) {
                const Entry e = e_next;
                {
                    const uint32_t nidx = base + THREADS + tid;
                    e_next = nidx < t.end ? ent_load(ent_in + nidx) : make_uint2(0u, 0u);
                }
                fetch_rows32(bins, dstride, e.x, e.y != 0u, s_tile, lane);
                if (e.y) {
                    const uint32_t cls = ent_cls(e), w = ent_w(e);
                    uint32_t sb0 = 0, sb1 = 0;  // the row's sampled bins, byte jj = feature jj of the draw
                    if (REG) {
                        const unsigned long long val = reg_cell(w, cls);  // `cls` is the 16-bit target here
                        ymn = min(ymn, cls); ymx = max(ymx, cls);
#pragma unroll
                        for (int jj = 0; jj < 8; ++jj)
                            if (jj < nb) {
                                const uint32_t bj = trow[featr[jj]];
                                if (jj < 4) sb0 |= bj << (8 * jj); else sb1 |= bj << (8 * (jj - 4));
                                atomicAdd((unsigned long long*)&s_hist[jj * BIGP_FS + (bj >> 3) * 20u + (bj & 7u) * 2u], val);
                            }
                    } else if (packed) {
                        const uint32_t coff = cls >> 1, val = w << (16u * (cls & 1u));
#pragma unroll
                        for (int jj = 0; jj < 8; ++jj)
                            if (jj < nb) {
                                const uint32_t bj = trow[featr[jj]];
                                if (jj < 4) sb0 |= bj << (8 * jj); else sb1 |= bj << (8 * (jj - 4));
                                atomicAdd(&s_hist[jj * BIGP_FS + (bj >> 3) * 20u + (bj & 7u) * 2u + coff], val);
                            }
                    } else {
#pragma unroll
                        for (int jj = 0; jj < 8; ++jj)
                            if (jj < nb) {
                                const uint32_t bj = trow[featr[jj]];
                                if (jj < 4) sb0 |= bj << (8 * jj); else sb1 |= bj << (8 * (jj - 4));
                                atomicAdd(&s_hist[jj * g.FS + hist_index(bj, cls, C)], w);
                            }
                    }
                    if (stash_w) __stcg(stash_w + (base + tid), make_uint2(sb0, sb1));
                }
                __syncwarp();
            }
            if (REG && pos == 0) {  // is the node's target constant?  (its parent could not know)
                ymn = __reduce_min_sync(0xffffffffu, ymn); ymx = __reduce_max_sync(0xffffffffu, ymx);
                if (lane == 0) { atomicMin(&s_ymin, ymn); atomicMax(&s_ymax, ymx); }
            }
        } else {
            uint32_t ymn = 0xFFFFFFFFu, ymx = 0u;
            constexpr int U = 4;
            for (uint32_t base = t.begin + tid; base < t.end; base += THREADS * U) {
                Entry e[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    uint32_t idx = base + u * THREADS;
                    e[u] = idx < t.end ? ent_load(ent_in + idx) : make_uint2(0u, 0u);
                }
                const int jr = (int)lane % nb;  // rotate features across lanes: spreads same-address atomics
                for (int jj = 0; jj < nb; ++jj) {
                    int j = jj + jr;
                    if (j >= nb) j -= nb;
                    uint32_t f = s_feat[j];
                    uint32_t b[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) b[u] = e[u].y ? (uint32_t)bins[(uint64_t)e[u].x * dstride + f] : 0u;
#pragma unroll
                    for (int u = 0; u < U; ++u)
                        if (e[u].y) {
                            if (REG) atomicAdd((unsigned long long*)&s_hist[j * BIGP_FS + (b[u] >> 3) * 20u + (b[u] & 7u) * 2u], reg_cell(ent_w(e[u]), ent_cls(e[u])));
                            else atomicAdd(&s_hist[j * g.FS + hist_index(b[u], ent_cls(e[u]), C)], ent_w(e[u]));
                        }
                }
                if (REG)
#pragma unroll
                    for (int u = 0; u < U; ++u) if (e[u].y) { ymn = min(ymn, ent_cls(e[u])); ymx = max(ymx, ent_cls(e[u])); }
            }
            if (REG && pos == 0) {
                ymn = __reduce_min_sync(0xffffffffu, ymn); ymx = __reduce_max_sync(0xffffffffu, ymx);
                if (lane == 0) { atomicMin(&s_ymin, ymn); atomicMax(&s_ymax, ymx); }
            }
        }
        __syncthreads();

        // ---- Gini split search: one warp per feature
        for (int j = wid; j < nb; j += NW) {
            double bg; uint32_t bb;
            bool want_tot = (pos == 0 && j == 0);
            if (REG) {
                unsigned long long P = 0, tot;
                scan_feature_reg(s_hist + j * BIGP_FS, g.msl, lane, bg, bb, P, tot);
                if (lane == 0) {
                    s_fP[j] = P;
                    if (want_tot) { s_totP = tot; s_N = reg_N(tot); }
                }
            } else if (packed) {
                uint32_t P0 = 0, P1 = 0, T0, T1;
                scan_feature_packed4(s_hist + j * BIGP_FS, g.msl, lane, bg, bb, P0, P1, T0, T1);
                if (lane == 0) {
                    s_fL[j][0] = P0 & 0xFFFFu; s_fL[j][1] = P0 >> 16; s_fL[j][2] = P1 & 0xFFFFu; s_fL[j][3] = P1 >> 16;
                    if (want_tot) {
                        s_tot[0] = T0 & 0xFFFFu; s_tot[1] = T0 >> 16; s_tot[2] = T1 & 0xFFFFu; s_tot[3] = T1 >> 16;
                        s_N = (T0 & 0xFFFFu) + (T0 >> 16) + (T1 & 0xFFFFu) + (T1 >> 16);
                    }
                }
            } else {
                scan_feature_t<MSE>(s_hist + j * g.FS, C, g.msl, lane, bg, bb, want_tot ? s_tot : nullptr, want_tot ? &s_N : nullptr);
            }
            if (lane == 0) { s_fgain[j] = bg; s_fbin[j] = bb; }
        }
        __syncthreads();
        if (tid == 0) {
            if (pos == 0) {
                bool pure = false;
                if (REG) pure = s_ymin == s_ymax;
                else for (int c = 0; c < C; ++c) pure |= (s_tot[c] == s_N);
                if (stop_rule(g, s_N, t.depth, pure)) s_stop = 1;
            }
            if (primary) {
                for (int j = 0; j < nb; ++j) {
                    if (s_fbin[j] == 0xFFFFFFFFu) continue;
                    uint32_t key = (s_feat[j] << 8) | s_fbin[j];  // ties: the feature drawn first (positions are visited in order), then lowest bin
                    if (s_fgain[j] > s_bgain) {
                        s_bgain = s_fgain[j]; s_bkey = key; s_newj = j; s_found = 1;
                        s_sj = (!LITE && pos == 0 && g.stash != nullptr && use_tile && nb <= 8 && t.end - t.begin <= g.stash_max_rows) ? (uint32_t)j : 0xFFFFFFFFu;
                    }
                }
            } else {
                for (int j = 0; j < nb; ++j) {
                    if (s_fbin[j] == 0xFFFFFFFFu) continue;
                    s_bgain = s_fgain[j]; s_bkey = (s_feat[j] << 8) | s_fbin[j]; s_newj = j; s_found = 1; s_sj = 0xFFFFFFFFu;
                    break;  // fallback: first feature in draw order with a valid split
                }
            }
        }
        if (tid == 0 && packed && s_newj != 0xFFFFFFFFu) {
            if (REG) s_bestP = s_fP[s_newj];
            else for (int c = 0; c < C; ++c) s_bestL[c] = s_fL[s_newj][c];
        }
        __syncthreads();
        if (s_stop) break;
        if (s_newj != 0xFFFFFFFFu && !packed) {
            // left class counts of the new best split, kept because the histogram may be overwritten
            const uint32_t bb = s_bkey & 0xFFu;
            const uint32_t* hl = s_hist + s_newj * g.FS + lane * (8u * C + 1u);
            for (int c = wid; c < C; c += NW) {
                uint32_t v = 0;
#pragma unroll
                for (int q = 0; q < 8; ++q) v += (lane * 8u + q <= bb) ? hl[q * C + c] : 0u;
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
                if (lane == 0) s_bestL[c] = v;
            }
        }
        pos += nb;
        __syncthreads();
    }

    const uint32_t r = t.end - t.begin;
    if (s_stop || !s_found) {
        // leaf: root that needs no split, or no feature separates the rows
        if (tid == 0) {
            if (REG) reg_write_leaf(nodes, leafc, t.slot, t.begin, s_N, reg_S(s_totP), s_ymin == s_ymax, s_ymin);
            else {
                int pc = -1;
                for (int c = 0; c < C; ++c) if (s_tot[c] == s_N) pc = c;
                if (pc >= 0) write_pure_leaf(nodes, t.slot, s_N, (uint32_t)pc);
                else {
                    for (int c = 0; c < C; ++c) leafc[(uint64_t)(t.begin >> 1) * C + c] = s_tot[c];
                    write_mixed_leaf(nodes, t.slot, t.begin);
                }
            }
            atomicAdd(&g.stats->nodes, 1ull);
            atomicAdd(&g.stats->leaves, 1ull);
            if (!s_stop) { atomicAdd(&g.stats->searched, 1ull); atomicAdd(&g.stats->row_visits_big, (unsigned long long)r); }
            atomicMax(&g.stats->max_depth, t.depth);
        }
        return;
    }

    // ---- row partition (SK/tree/_partitioner.pyx:248-279): left rows packed from the front in order,
    // right rows from the back; tree structure depends only on the sets, not on the order.
    const uint32_t bf = s_bkey >> 8, bb = s_bkey & 0xFFu;
    const uint32_t sj = s_sj;
    const uint2* stash_r = (!LITE && g.stash != nullptr && sj != 0xFFFFFFFFu) ? g.stash + (uint64_t)t.tree * g.n_cap : nullptr;
    uint32_t lbase = 0, rbase = 0;
    int par = 0;
    constexpr int PU = 4;  // chunks per barrier
    for (uint32_t c0 = t.begin; c0 < t.end; c0 += THREADS * PU, par ^= 1) {
        Entry e[PU];
        uint32_t balL[PU], balR[PU];
        bool goL[PU];
#pragma unroll
        for (int u = 0; u < PU; ++u) {
            uint32_t idx = c0 + u * THREADS + tid;
            e[u] = idx < t.end ? ent_load(ent_in + idx) : make_uint2(0u, 0u);
        }
        if (stash_r) {  // the split feature's bin of every row was left in the stash by the histogram pass
#pragma unroll
            for (int u = 0; u < PU; ++u) {
                const uint32_t idx = c0 + u * THREADS + tid;
                uint2 sv = make_uint2(0u, 0u);
                if (e[u].y) sv = __ldcs(stash_r + idx);
                goL[u] = e[u].y && (((sj < 4u ? sv.x : sv.y) >> (8u * (sj & 3u))) & 0xFFu) <= bb;
            }
        } else {
#pragma unroll
            for (int u = 0; u < PU; ++u) goL[u] = e[u].y && bins[(uint64_t)e[u].x * dstride + bf] <= bb;
        }
#pragma unroll
        for (int u = 0; u < PU; ++u) {
            balL[u] = __ballot_sync(0xffffffffu, goL[u]);
            balR[u] = __ballot_sync(0xffffffffu, e[u].y && !goL[u]);
            if (lane == 0) s_wl[par][u][wid] = __popc(balL[u]) | (__popc(balR[u]) << 16);  // (left, right) counts packed
        }
        __syncthreads();
        {   // packed exclusive scan over the PU x NW (chunk, warp) counts, in element order: lane = u * NW + w
            static_assert(PU * NW <= 32, "one warp scans all chunk counts");
            uint32_t cnt = lane < PU * NW ? s_wl[par][lane / NW][lane % NW] : 0u, incl = cnt;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                uint32_t a = __shfl_up_sync(0xffffffffu, incl, off);
                if (lane >= (uint32_t)off) incl += a;
            }
            const uint32_t tot = __shfl_sync(0xffffffffu, incl, 31);
#pragma unroll
            for (int u = 0; u < PU; ++u) {
                const uint32_t pre = __shfl_sync(0xffffffffu, incl - cnt, u * NW + wid);
                const uint32_t pl = pre & 0xFFFFu, pr = pre >> 16;
                if (goL[u]) ent_store(ent_out + t.begin + lbase + pl + __popc(balL[u] & lanemask_lt()), e[u]);
                else if (e[u].y) ent_store(ent_out + t.end - 1u - (rbase + pr + __popc(balR[u] & lanemask_lt())), e[u]);
            }
            lbase += tot & 0xFFFFu; rbase += tot >> 16;
        }
    }
    const uint32_t nLd = lbase;
    if (nLd == 0 || nLd == r) {  // cannot happen for a valid split; never loop on it
        if (tid == 0) g.stats->error = 4;
        return;
    }

    // ---- node record + children (leaf children are finished here, the rest are queued)
    if (tid == 0) {
        if (REG) emit_split_reg(g, t, nodes, leafc, s_N, reg_S(s_totP), reg_N(s_bestP), reg_S(s_bestP), nLd, bf, bb, s_bgain);
        else emit_split(g, t, nodes, leafc, s_tot, s_bestL, s_N, nLd, bf, bb, presplit ? -1.0 : s_bgain);
        atomicAdd(&g.stats->row_visits_big, (unsigned long long)r);
    }
}

// ---------------------------------------------------------------------------------------------
// One warp finishes the subtree under a node with <= 32 distinct rows.  Lane i holds row i of the
// node; a candidate threshold is each lane's own bin; class counts left of it are accumulated in a
// per-lane shared column over a broadcast loop; the partition is a ballot + shuffle permutation.
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_grow_small(GrowCtx g, const Task* __restrict__ tasks, uint32_t ntasks) {
    extern __shared__ uint32_t s_mem[];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const int C = g.C, d = g.d, m = g.m;
    // per-warp shared layout (words): L[C][32] | tot[C] | src[32] | stack[32][3] | perm[(d+1)/2]
    const uint32_t per_warp = (uint32_t)C * 32u + (uint32_t)C + 32u + 96u + (uint32_t)((d + 1) / 2);
    uint32_t* s_L = s_mem + warp * per_warp;
    uint32_t* s_tot = s_L + C * 32;
    uint32_t* s_src = s_tot + C;
    uint32_t* s_stack = s_src + 32;
    uint16_t* s_perm = (uint16_t*)(s_stack + 96);
    const uint8_t* __restrict__ bins = g.bins_rm;
    const uint32_t dstride = g.dstride;

    for (uint32_t ti = blockIdx.x * WARPS + warp; ti < ntasks; ti += gridDim.x * WARPS) {
        const Task t = tasks[ti];
        const TreeDesc td = g.trees[t.tree];
        const Entry* __restrict__ ent = ((t.depth & 1u) ? g.entB : g.entA) + (uint64_t)t.tree * g.n_cap;
        uint2* nodes = g.nodes + (uint64_t)td.arena * g.node_stride;
        uint32_t* leafc = g.leafc + (uint64_t)td.arena * g.leafc_stride;
        const uint32_t r = t.end - t.begin;
        Entry e = lane < r ? ent_load(ent + t.begin + lane) : make_uint2(0u, 0u);
        uint32_t cls = ent_cls(e);
        uint32_t lo = 0, hi = r, depth = t.depth, slot = t.slot, sp = 0;
        unsigned long long visits = 0, n_nodes = 0, n_leaves = 0, n_search = 0, n_fb = 0;
        uint32_t maxd = 0;
        while (true) {
            const bool active = lane >= lo && lane < hi;
            const uint32_t w = active ? ent_w(e) : 0u;
            const uint32_t abs_begin = t.begin + lo;
            // node totals (SK/tree/_criterion.pyx:302-370)
            for (int c = lane; c < C; c += 32) s_tot[c] = 0;
            __syncwarp();
            if (active) atomicAdd(&s_tot[cls], w);
            __syncwarp();
            uint32_t N = w;
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) N += __shfl_xor_sync(0xffffffffu, N, off);
            const uint32_t cls_lo = __shfl_sync(0xffffffffu, cls, lo);
            const bool pure = s_tot[cls_lo] == N;
            n_nodes++;
            maxd = max(maxd, depth);
            bool found = false;
            double best_gain = -1.0;
            uint32_t best_key = 0xFFFFFFFFu;
            if (!stop_rule(g, N, depth, pure)) {
                visits += hi - lo;
                n_search++;
                for (int f = lane; f < d; f += 32) s_perm[f] = (uint16_t)f;
                __syncwarp();
                uint4 rnd = make_uint4(0, 0, 0, 0);
                for (int i = 0; i < d; ++i) {
                    if (i >= m && found) break;
                    uint32_t f = 0;
                    if (lane == 0) {
                        if ((i & 3) == 0)
                            rnd = philox4x32_10(abs_begin, depth, td.gid, TST_PURPOSE_FEAT | ((uint32_t)(i >> 2) << 8), g.k0, g.k1);
                        int j = i + (int)mulhi32(pick4(rnd, i & 3), (uint32_t)(d - i));
                        uint16_t fj = s_perm[j];
                        s_perm[j] = s_perm[i];
                        s_perm[i] = fj;
                        f = fj;
                        if (i == m) n_fb++;
                    }
                    f = __shfl_sync(0xffffffffu, f, 0);
                    const uint32_t b = active ? bins[(uint64_t)e.x * dstride + f] : 0u;
                    const uint32_t packed = b | (cls << 8) | (w << 16);
                    for (int c = 0; c < C; ++c) s_L[c * 32 + lane] = 0;
                    for (uint32_t kk = lo; kk < hi; ++kk) {
                        uint32_t pk = __shfl_sync(0xffffffffu, packed, kk);
                        if (active && (pk & 0xFFu) <= b) s_L[((pk >> 8) & 0xFFu) * 32 + lane] += pk >> 16;
                    }
                    uint32_t NL = 0;
                    unsigned long long SL = 0, SR = 0;
                    for (int c = 0; c < C; ++c) {
                        uint32_t L = s_L[c * 32 + lane], R = s_tot[c] - L;
                        NL += L;
                        SL += (unsigned long long)L * (g.criterion == 1u ? (uint32_t)c : L);
                        SR += (unsigned long long)R * (g.criterion == 1u ? (uint32_t)c : R);
                    }
                    const uint32_t NR = N - NL;
                    double gn = -1.0;
                    uint32_t gb = 0xFFFFFFFFu;
                    if (active && NL >= g.msl && NR >= g.msl) { gn = g.criterion == 1u ? mse_gain(SL, NL, SR, NR) : gini_gain(SL, NL, SR, NR); gb = b; }
#pragma unroll
                    for (int off = 16; off >= 1; off >>= 1) {
                        double og = __shfl_xor_sync(0xffffffffu, gn, off);
                        uint32_t ob = __shfl_xor_sync(0xffffffffu, gb, off);
                        if (og > gn || (og == gn && ob < gb)) { gn = og; gb = ob; }
                    }
                    if (gb != 0xFFFFFFFFu) {
                        uint32_t key = (f << 8) | gb;
                        if (i < m) {
                            if (gn > best_gain) { best_gain = gn; best_key = key; found = true; }  // ties: the feature drawn first
                        } else {
                            best_gain = gn; best_key = key; found = true;
                        }
                    }
                }
            }
            if (!found) {
                // leaf
                n_leaves++;
                if (pure) {
                    if (lane == 0) write_pure_leaf(nodes, slot, N, cls_lo);
                } else {
                    for (int c = lane; c < C; c += 32) leafc[(uint64_t)(abs_begin >> 1) * C + c] = s_tot[c];
                    if (lane == 0) write_mixed_leaf(nodes, slot, abs_begin);
                }
                if (sp == 0) break;
                --sp;
                uint32_t lh = s_stack[sp * 3];
                lo = lh & 0xFFu; hi = lh >> 8; depth = s_stack[sp * 3 + 1]; slot = s_stack[sp * 3 + 2];
                __syncwarp();
                continue;
            }
            // partition inside the warp (stable)
            const uint32_t bf = best_key >> 8, bb = best_key & 0xFFu;
            const bool goL = active && bins[(uint64_t)e.x * dstride + bf] <= bb;
            const uint32_t balL = __ballot_sync(0xffffffffu, goL);
            const uint32_t balR = __ballot_sync(0xffffffffu, active && !goL);
            const uint32_t nLd = __popc(balL);
            if (nLd == 0 || nLd == hi - lo) {  // cannot happen for a valid split; never loop on it
                if (lane == 0) g.stats->error = 4;
                break;
            }
            uint32_t dest = lane;
            if (goL) dest = lo + __popc(balL & lanemask_lt());
            else if (active) dest = lo + nLd + __popc(balR & lanemask_lt());
            s_src[dest] = lane;
            __syncwarp();
            const uint32_t src = s_src[lane];
            __syncwarp();
            e.x = __shfl_sync(0xffffffffu, e.x, src);
            e.y = __shfl_sync(0xffffffffu, e.y, src);
            cls = ent_cls(e);
            const uint32_t right_slot = slot + 2u * nLd;
            if (lane == 0) {
                write_internal(nodes, slot, right_slot, bf, bb);
                if (g.imp != nullptr) {
                    unsigned long long S = 0;
                    for (int c = 0; c < C; ++c) S += (unsigned long long)s_tot[c] * (g.criterion == 1u ? (uint32_t)c : s_tot[c]);
                    imp_add(g, td.imp_row, bf, best_gain, g.criterion == 1u ? imp_parent_mse(S, N) : imp_parent_gini(S, N));
                }
                s_stack[sp * 3] = (lo + nLd) | (hi << 8);
                s_stack[sp * 3 + 1] = depth + 1u;
                s_stack[sp * 3 + 2] = right_slot;
            }
            __syncwarp();
            sp++;
            hi = lo + nLd;
            depth += 1u;
            slot += 1u;
        }
        if (lane == 0) {
            atomicAdd(&g.stats->row_visits_small, visits);
            atomicAdd(&g.stats->nodes, n_nodes);
            atomicAdd(&g.stats->leaves, n_leaves);
            atomicAdd(&g.stats->searched, n_search);
            if (n_fb) atomicAdd(&g.stats->fallback, n_fb);
            atomicMax(&g.stats->max_depth, maxd);
        }
        __syncwarp();
    }
}
