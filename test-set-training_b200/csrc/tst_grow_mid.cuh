// tst_grow_mid.cuh -- warp-per-node kernel for nodes with 33 .. 32*R distinct rows.
//
// The CTA-per-node kernel pays ~16 us of barriers and serial bookkeeping per node, which dominated
// the levels where a tree has tens of thousands of nodes of a few dozen rows.  Here one warp owns
// one node: its rows stay in registers (R per lane, element i = lane + 32 k), one 256 x C shared
// histogram per warp is rebuilt per sampled feature, the split scan is the same warp-shuffle prefix
// scan as the CTA kernel (scan_feature), and the partition is a ballot prefix written straight to
// the tree's other entry buffer.  No __syncthreads anywhere.  Children are routed like emit_split.
#pragma once
#include "tst_grow.cuh"

__device__ __forceinline__ void warp_argmax_db(double& gn, uint32_t& key) {
    // (gain desc, key asc); every lane ends with the winner
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        double og = __shfl_xor_sync(0xffffffffu, gn, off);
        uint32_t ok = __shfl_xor_sync(0xffffffffu, key, off);
        if (og > gn || (og == gn && ok < key)) { gn = og; key = ok; }
    }
}

template <int WARPS, int R>
__global__ void __launch_bounds__(WARPS * 32) k_grow_mid(GrowCtx g, const Task* __restrict__ tasks, uint32_t ntasks, const uint32_t* __restrict__ ntasks_dev) {
    extern __shared__ uint32_t s_mem[];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const int C = g.C, d = g.d, m = g.m;
    // per-warp shared layout (words): hist[FS] | tot[C] | bestL[C] | perm[(d+1)/2]
    const uint32_t per_warp = (uint32_t)g.FS + 2u * (uint32_t)C + (uint32_t)((d + 1) / 2);
    uint32_t* s_hist = s_mem + warp * per_warp;
    uint32_t* s_tot = s_hist + g.FS;
    uint32_t* s_bestL = s_tot + C;
    uint16_t* s_perm = (uint16_t*)(s_bestL + C);
    const uint8_t* __restrict__ bins = g.bins_rm;
    const uint32_t dstride = g.dstride;
    const uint32_t ti = blockIdx.x * WARPS + warp;
    if (ntasks_dev) ntasks = min(ntasks, *ntasks_dev);
    if (ti >= ntasks) return;

    const Task t = tasks[ti];
    const TreeDesc td = g.trees[t.tree];
    const Entry* __restrict__ ent_in = ((t.depth & 1u) ? g.entB : g.entA) + (uint64_t)t.tree * g.n_cap;
    Entry* __restrict__ ent_out = ((t.depth & 1u) ? g.entA : g.entB) + (uint64_t)t.tree * g.n_cap;
    uint2* nodes = g.nodes + (uint64_t)td.arena * g.node_stride;
    uint32_t* leafc = g.leafc + (uint64_t)td.arena * g.leafc_stride;
    const uint32_t r = t.end - t.begin;

    Entry e[R];
#pragma unroll
    for (int k = 0; k < R; ++k) {
        uint32_t i = lane + 32u * k;
        e[k] = i < r ? ent_load(ent_in + t.begin + i) : make_uint2(0u, 0u);  // multiplicity 0 = padding
    }
    for (int f = lane; f < d; f += 32) s_perm[f] = (uint16_t)f;
    __syncwarp();

    bool found = false, stop = false;
    double best_gain = -1.0;
    uint32_t best_key = 0xFFFFFFFFu, N = 0;
    uint4 rnd = make_uint4(0, 0, 0, 0);
    for (int i = 0; i < d; ++i) {
        if (i >= m && found) break;
        // per-node feature draw (lane 0 walks the Fisher-Yates permutation)
        uint32_t f = 0;
        if (lane == 0) {
            if ((i & 3) == 0)
                rnd = philox4x32_10(t.begin, t.depth, td.gid, TST_PURPOSE_FEAT | ((uint32_t)(i >> 2) << 8), g.k0, g.k1);
            int j = i + (int)mulhi32(pick4(rnd, i & 3), (uint32_t)(d - i));
            uint16_t fj = s_perm[j];
            s_perm[j] = s_perm[i];
            s_perm[i] = fj;
            f = fj;
            if (i == m) atomicAdd(&g.stats->fallback, 1ull);
        }
        f = __shfl_sync(0xffffffffu, f, 0);
        // class histogram of this feature
        for (int w = lane; w < g.FS; w += 32) s_hist[w] = 0;
        uint32_t b[R];
#pragma unroll
        for (int k = 0; k < R; ++k) b[k] = e[k].y ? (uint32_t)bins[(uint64_t)e[k].x * dstride + f] : 0u;
        __syncwarp();
#pragma unroll
        for (int k = 0; k < R; ++k)
            if (e[k].y) atomicAdd(&s_hist[hist_index(b[k], ent_cls(e[k]), C)], ent_w(e[k]));
        __syncwarp();
        double bg;
        uint32_t bb;
        scan_feature(s_hist, C, g.msl, lane, bg, bb, i == 0 ? s_tot : nullptr, i == 0 ? &N : nullptr, g.criterion);
        if (i == 0) {
            N = __shfl_sync(0xffffffffu, N, 0);
            __syncwarp();
            bool pure = false;
            for (int c = 0; c < C; ++c) pure |= (s_tot[c] == N);
            if (stop_rule(g, N, t.depth, pure)) { stop = true; break; }
        }
        bool newbest = false;
        if (bb != 0xFFFFFFFFu) {
            uint32_t key = (f << 8) | bb;
            if (i < m) {
                if (bg > best_gain) newbest = true;  // ties: the feature drawn first
            } else {
                newbest = true;
            }
            if (newbest) { best_gain = bg; best_key = key; found = true; }
        }
        if (newbest) {
            // left class counts of the new best split (the histogram is overwritten by the next feature)
            const uint32_t* hl = s_hist + lane * (8u * C + 1u);
            for (int c = 0; c < C; ++c) {
                uint32_t v = 0;
#pragma unroll
                for (int q = 0; q < 8; ++q) v += (lane * 8u + q <= bb) ? hl[q * C + c] : 0u;
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
                if (lane == 0) s_bestL[c] = v;
            }
        }
        __syncwarp();
    }

    if (stop || !found) {
        if (lane == 0) {
            int pc = -1;
            for (int c = 0; c < C; ++c) if (s_tot[c] == N) pc = c;
            if (pc >= 0) write_pure_leaf(nodes, t.slot, N, (uint32_t)pc);
            else {
                for (int c = 0; c < C; ++c) leafc[(uint64_t)(t.begin >> 1) * C + c] = s_tot[c];
                write_mixed_leaf(nodes, t.slot, t.begin);
            }
            atomicAdd(&g.stats->nodes, 1ull);
            atomicAdd(&g.stats->leaves, 1ull);
            if (!stop) { atomicAdd(&g.stats->searched, 1ull); atomicAdd(&g.stats->row_visits_mid, (unsigned long long)r); }
            atomicMax(&g.stats->max_depth, t.depth);
        }
        return;
    }

    // ---- stable partition into the other entry buffer
    const uint32_t bf = best_key >> 8, bbin = best_key & 0xFFu;
    uint32_t goL = 0;  // bit k: element k goes left
#pragma unroll
    for (int k = 0; k < R; ++k)
        if (e[k].y && bins[(uint64_t)e[k].x * dstride + bf] <= bbin) goL |= 1u << k;
    uint32_t nLd = 0;
#pragma unroll
    for (int k = 0; k < R; ++k) nLd += __popc(__ballot_sync(0xffffffffu, (goL >> k) & 1u));
    if (nLd == 0 || nLd == r) {
        if (lane == 0) g.stats->error = 4;
        return;
    }
    uint32_t lbase = 0, rbase = 0;
#pragma unroll
    for (int k = 0; k < R; ++k) {
        const bool l = (goL >> k) & 1u, v = e[k].y != 0u;
        const uint32_t balL = __ballot_sync(0xffffffffu, l), balR = __ballot_sync(0xffffffffu, v && !l);
        if (l) ent_store(ent_out + t.begin + lbase + __popc(balL & lanemask_lt()), e[k]);
        else if (v) ent_store(ent_out + t.begin + nLd + rbase + __popc(balR & lanemask_lt()), e[k]);
        lbase += __popc(balL);
        rbase += __popc(balR);
    }
    if (lane == 0) {
        emit_split(g, t, nodes, leafc, s_tot, s_bestL, N, nLd, bf, bbin, best_gain);
        atomicAdd(&g.stats->row_visits_mid, (unsigned long long)r);
    }
}

// ---------------------------------------------------------------------------------------------
// <= 4 classes: the same warp-per-node kernel with packed arithmetic.  A node of <= 256 rows weighs
// <= 65280 < 2^16, so a histogram cell is two 32-bit words holding four 16-bit class counts (one
// shared atomic per row, 2 KB per feature), the prefix scan runs on the packed words (one 64-bit
// shuffle chain for all classes), and only the OCCUPIED bins are turned into candidates, compacted
// into a shared list and spread evenly over the lanes, so the fp64 gain is evaluated ~rows/32 times
// per lane instead of 8.  The next feature's bins are prefetched while the current one is scanned.
#define MID4_HIST 640u   // 32 lanes x (8 bins x 2 words + 4 pad): 80-byte lane segments, conflict-free 128-bit access
#define MID4_CAND 1024u  // 256 candidates: (P0, P1) pairs [0, 512) | bin [512, 768) | gain estimate [768, 1024)

#ifndef MID_MINB
#define MID_MINB 5
#endif
// The candidate list of a node holds at most min(256, rows) entries: a kernel instantiated for <= 32 * R rows reserves 4 * NC words
// per warp with NC = 32 * R candidates (NC = 256 for R = 8: the MID4_CAND layout).
__host__ __device__ constexpr uint32_t mid4_ncand(int R) { return 32u * (uint32_t)R; }
__host__ __device__ constexpr uint32_t mid4_per_warp(int R, int d) { return (MID4_HIST + 4u * mid4_ncand(R) + 8u + (uint32_t)((d + 1) / 2) + 3u) & ~3u; }

template <int WARPS, int R, bool REG = false, int MINB = MID_MINB>  // REG: criterion 2 -- a cell is (weight | sum of w * yq << 24) instead of four 16-bit class counts
__global__ void __launch_bounds__(WARPS * 32, MINB) k_grow_mid4(GrowCtx g, const Task* __restrict__ tasks, uint32_t ntasks, const uint32_t* __restrict__ ntasks_dev) {
    static_assert(REG || 32 * R * 255 < 65536, "packed 16-bit class counts need <= 256 rows per node");
    extern __shared__ __align__(16) uint32_t s_mem[];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const int C = g.C, d = g.d, m = g.m;
    constexpr uint32_t NC = mid4_ncand(R);  // candidates: (P0, P1) pairs [0, 2 NC) | bin [2 NC, 3 NC) | gain estimate [3 NC, 4 NC)
    const uint32_t per_warp = mid4_per_warp(R, d);  // 16-byte aligned warp slices
    uint32_t* s_hist = s_mem + warp * per_warp;
    uint32_t* s_cand = s_hist + MID4_HIST;
    uint32_t* s_tot = s_cand + 4u * NC;  // [4] totals, [4] left counts of the best split (for emit_split)
    uint32_t* s_bestL = s_tot + 4;
    uint16_t* s_perm = (uint16_t*)(s_bestL + 4);
    const uint8_t* __restrict__ bins = g.bins_rm;
    const uint32_t dstride = g.dstride;
    const uint32_t ti = blockIdx.x * WARPS + warp;
    if (ntasks_dev) ntasks = min(ntasks, *ntasks_dev);
    if (ti >= ntasks) return;

    const Task t = tasks[ti];
    const TreeDesc td = g.trees[t.tree];
    const Entry* __restrict__ ent_in = ((t.depth & 1u) ? g.entB : g.entA) + (uint64_t)t.tree * g.n_cap;
    Entry* __restrict__ ent_out = ((t.depth & 1u) ? g.entA : g.entB) + (uint64_t)t.tree * g.n_cap;
    uint2* nodes = g.nodes + (uint64_t)td.arena * g.node_stride;
    uint32_t* leafc = g.leafc + (uint64_t)td.arena * g.leafc_stride;
    const uint32_t r = t.end - t.begin;

    Entry e[R];
#pragma unroll
    for (int k = 0; k < R; ++k) {
        uint32_t i = lane + 32u * k;
        e[k] = i < r ? ent_load(ent_in + t.begin + i) : make_uint2(0u, 0u);
    }
    uint4 rnd = make_uint4(0, 0, 0, 0);
    const uint32_t draw_j = warp_draw_first(s_perm, d, m, t.begin, t.depth, td.gid, g.k0, g.k1, lane);  // the first m positions of the node's draw

    bool found = false, stop = false;
    double best_gain = -1.0;
    uint32_t best_key = 0xFFFFFFFFu, bestP0 = 0, bestP1 = 0;
    uint32_t N = 0, T0 = 0, T1 = 0;
    uint32_t ymn = 0xFFFFFFFFu, ymx = 0u;  // criterion 2: is the node's target constant?
    if (REG) {
#pragma unroll
        for (int k = 0; k < R; ++k) if (e[k].y) { ymn = min(ymn, ent_cls(e[k])); ymx = max(ymx, ent_cls(e[k])); }
        ymn = __reduce_min_sync(0xffffffffu, ymn); ymx = __reduce_max_sync(0xffffffffu, ymx);
    }
    uint32_t bnext[R];
    {
        const uint32_t f0 = s_perm[0];
#pragma unroll
        for (int k = 0; k < R; ++k) bnext[k] = e[k].y ? (uint32_t)bins[(uint64_t)e[k].x * dstride + f0] : 0u;
    }
    for (int i = 0; i < d; ++i) {
        if (i >= m && found) break;
        if (i >= m) {  // fallback: one more position of the draw, no prefetch possible
            if (i == m) warp_draw_materialize(s_perm, d, m, draw_j, lane);
            if (lane == 0) {
                if (i == m || (i & 3) == 0)
                    rnd = philox4x32_10(t.begin, t.depth, td.gid, TST_PURPOSE_FEAT | ((uint32_t)(i >> 2) << 8), g.k0, g.k1);
                int j = i + (int)mulhi32(pick4(rnd, i & 3), (uint32_t)(d - i));
                uint16_t fj = s_perm[j];
                s_perm[j] = s_perm[i];
                s_perm[i] = fj;
                if (i == m) atomicAdd(&g.stats->fallback, 1ull);
            }
            __syncwarp();
        }
        const uint32_t f = s_perm[i];
        uint32_t b[R];
        if (i < m) {
#pragma unroll
            for (int k = 0; k < R; ++k) b[k] = bnext[k];
            if (i + 1 < m) {
                const uint32_t fn = s_perm[i + 1];
#pragma unroll
                for (int k = 0; k < R; ++k) bnext[k] = e[k].y ? (uint32_t)bins[(uint64_t)e[k].x * dstride + fn] : 0u;
            }
        } else {
#pragma unroll
            for (int k = 0; k < R; ++k) b[k] = e[k].y ? (uint32_t)bins[(uint64_t)e[k].x * dstride + f] : 0u;
        }
        // packed class histogram of this feature
        {   // the lane's own 8 bins are zeroed here and read back below with 128-bit accesses
            uint4* hz = (uint4*)(s_hist + lane * 20u);
            hz[0] = hz[1] = hz[2] = hz[3] = make_uint4(0u, 0u, 0u, 0u);
        }
        __syncwarp();
#pragma unroll
        for (int k = 0; k < R; ++k) {
            if (e[k].y) {
                if (REG) atomicAdd((unsigned long long*)&s_hist[(b[k] >> 3) * 20u + (b[k] & 7u) * 2u], reg_cell(ent_w(e[k]), ent_cls(e[k])));
                else atomicAdd(&s_hist[(b[k] >> 3) * 20u + (b[k] & 7u) * 2u + (ent_cls(e[k]) >> 1)], ent_w(e[k]) << (16u * (ent_cls(e[k]) & 1u)));
            }
        }
        __syncwarp();
        // packed prefix scan over bins
        const uint4* hl = (const uint4*)(s_hist + lane * 20u);
        uint32_t p0[8], p1[8], run0 = 0, run1 = 0, occ = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const uint4 hq = hl[q >> 1];  // two bins per 128-bit load (the compiler keeps one load per pair)
            const uint32_t h0 = (q & 1) ? hq.z : hq.x, h1 = (q & 1) ? hq.w : hq.y;
            occ |= ((h0 | h1) ? 1u : 0u) << q;
            if (REG) {  // one 64-bit field pair: the sum straddles the word boundary, so the carry must propagate
                const unsigned long long acc = ((unsigned long long)run0 | ((unsigned long long)run1 << 32)) + ((unsigned long long)h0 | ((unsigned long long)h1 << 32));
                run0 = (uint32_t)acc; run1 = (uint32_t)(acc >> 32);
            } else { run0 += h0; run1 += h1; }
            p0[q] = run0; p1[q] = run1;
        }
        unsigned long long run = (unsigned long long)run0 | ((unsigned long long)run1 << 32), incl = run;
        uint32_t cnt = __popc(occ), cincl = cnt;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            unsigned long long a = __shfl_up_sync(0xffffffffu, incl, off);
            uint32_t ca = __shfl_up_sync(0xffffffffu, cincl, off);
            if (lane >= (uint32_t)off) { incl += a; cincl += ca; }
        }
        const unsigned long long excl = incl - run, tot = __shfl_sync(0xffffffffu, incl, 31);
        const uint32_t e0 = (uint32_t)excl, e1 = (uint32_t)(excl >> 32);
        const uint32_t n_cand = __shfl_sync(0xffffffffu, cincl, 31);
        if (i == 0) {
            T0 = (uint32_t)tot; T1 = (uint32_t)(tot >> 32);
            const uint32_t t0 = T0 & 0xFFFFu, t1 = T0 >> 16, t2 = T1 & 0xFFFFu, t3 = T1 >> 16;
            N = REG ? reg_N(tot) : t0 + t1 + t2 + t3;
            const bool pure = REG ? ymn == ymx : (t0 == N || t1 == N || t2 == N || t3 == N);
            if (stop_rule(g, N, t.depth, pure)) { stop = true; break; }
        }
        // candidate list = occupied bins in ascending order: left counts packed | bin | fp32 gain estimate (SoA)
        {
            uint32_t j = cincl - cnt;
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if ((occ >> q) & 1u) {
                    if (REG) {
                        const unsigned long long P = excl + ((unsigned long long)p0[q] | ((unsigned long long)p1[q] << 32));
                        ((uint2*)s_cand)[j] = make_uint2((uint32_t)P, (uint32_t)(P >> 32));
                    } else
                    ((uint2*)s_cand)[j] = make_uint2(e0 + p0[q], e1 + p1[q]);  // one 64-bit store for both packed words
                    s_cand[2u * NC + j] = lane * 8u + q;
                    ++j;
                }
        }
        __syncwarp();
        const uint32_t t0 = T0 & 0xFFFFu, t1 = T0 >> 16, t2 = T1 & 0xFFFFu, t3 = T1 >> 16;
        const unsigned long long Stot = reg_S((unsigned long long)T0 | ((unsigned long long)T1 << 32));  // criterion 2
        // pass 1: fp32 estimates (a node here weighs < 2^16, so the sums of squares fit 32 bits)
        float lmax = 0.0f;
        for (uint32_t j = lane; j < n_cand; j += 32) {
            const uint2 PP = ((const uint2*)s_cand)[j];
            const uint32_t P0 = PP.x, P1 = PP.y;
            const uint32_t L0 = P0 & 0xFFFFu, L1 = P0 >> 16, L2 = P1 & 0xFFFFu, L3 = P1 >> 16;
            const unsigned long long P = (unsigned long long)P0 | ((unsigned long long)P1 << 32);
            const uint32_t NL = REG ? reg_N(P) : L0 + L1 + L2 + L3, NR = N - NL;
            float ge = -1.0f;
            if (REG) {
                if (NL >= g.msl && NR >= g.msl) {
                    ge = mse_gain_estimate(__ull2float_rn(reg_S(P)), __uint2float_rn(NL), __ull2float_rn(Stot - reg_S(P)), __uint2float_rn(NR));
                    lmax = fmaxf(lmax, ge);
                }
            } else
            if (NL >= g.msl && NR >= g.msl) {
                const uint32_t R0 = t0 - L0, R1 = t1 - L1, R2 = t2 - L2, R3 = t3 - L3;
                const uint32_t SL = L0 * L0 + L1 * L1 + L2 * L2 + L3 * L3, SR = R0 * R0 + R1 * R1 + R2 * R2 + R3 * R3;
                ge = gain_estimate(__uint2float_rn(SL), __uint2float_rn(NL), __uint2float_rn(SR), __uint2float_rn(NR));
                lmax = fmaxf(lmax, ge);
            }
            s_cand[3u * NC + j] = __float_as_uint(ge);
        }
        float thr = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(lmax)));  // positive floats order like integers
        if (i < m && found) thr = fmaxf(thr, __double2float_rd(best_gain));  // a feature that cannot reach the node's best is skipped
        thr *= TST_EST_MARGIN;
        // pass 2: exact fp64 gain of the contenders only
        double bg = -1.0;
        uint32_t bb = 0xFFFFFFFFu, bP0 = 0, bP1 = 0;
        for (uint32_t j = lane; j < n_cand; j += 32) {
            if (__uint_as_float(s_cand[3u * NC + j]) >= thr) {
                const uint2 PP = ((const uint2*)s_cand)[j];
                const uint32_t P0 = PP.x, P1 = PP.y, bin = s_cand[2u * NC + j];
                const uint32_t L0 = P0 & 0xFFFFu, L1 = P0 >> 16, L2 = P1 & 0xFFFFu, L3 = P1 >> 16;
                const unsigned long long P = (unsigned long long)P0 | ((unsigned long long)P1 << 32);
                const uint32_t NL = REG ? reg_N(P) : L0 + L1 + L2 + L3, NR = N - NL;
                const uint32_t R0 = t0 - L0, R1 = t1 - L1, R2 = t2 - L2, R3 = t3 - L3;
                const uint32_t SL = L0 * L0 + L1 * L1 + L2 * L2 + L3 * L3, SR = R0 * R0 + R1 * R1 + R2 * R2 + R3 * R3;
                const double gn = REG ? mse_gain(reg_S(P), NL, Stot - reg_S(P), NR) : gini_gain(SL, NL, SR, NR);
                if (gn > bg) { bg = gn; bb = bin; bP0 = P0; bP1 = P1; }  // ascending bins: strict > keeps the lowest
            }
        }
        double wg = bg;
        uint32_t wb = bb;
        warp_argmax_redux(wg, wb);
        if (wb != 0xFFFFFFFFu) {
            const uint32_t key = (f << 8) | wb;
            bool newbest;
            if (i < m) newbest = wg > best_gain;  // ties: the feature drawn first
            else newbest = true;
            if (newbest) {
                const uint32_t winner = __ffs(__ballot_sync(0xffffffffu, bb == wb && bg == wg)) - 1;
                bestP0 = __shfl_sync(0xffffffffu, bP0, winner);
                bestP1 = __shfl_sync(0xffffffffu, bP1, winner);
                best_gain = wg; best_key = key; found = true;
            }
        }
        __syncwarp();
    }

    if (lane == 0) {
        s_tot[0] = T0 & 0xFFFFu; s_tot[1] = T0 >> 16; s_tot[2] = T1 & 0xFFFFu; s_tot[3] = T1 >> 16;
        s_bestL[0] = bestP0 & 0xFFFFu; s_bestL[1] = bestP0 >> 16; s_bestL[2] = bestP1 & 0xFFFFu; s_bestL[3] = bestP1 >> 16;
    }
    __syncwarp();
    const unsigned long long Ptot = (unsigned long long)T0 | ((unsigned long long)T1 << 32), Pbest = (unsigned long long)bestP0 | ((unsigned long long)bestP1 << 32);
    if (stop || !found) {
        if (lane == 0) {
            int pc = -1;
            if (REG) reg_write_leaf(nodes, leafc, t.slot, t.begin, N, reg_S(Ptot), ymn == ymx, ymn);
            else {
            for (int c = 0; c < C; ++c) if (s_tot[c] == N) pc = c;
            if (pc >= 0) write_pure_leaf(nodes, t.slot, N, (uint32_t)pc);
            else {
                for (int c = 0; c < C; ++c) leafc[(uint64_t)(t.begin >> 1) * C + c] = s_tot[c];
                write_mixed_leaf(nodes, t.slot, t.begin);
            }
            }
            atomicAdd(&g.stats->nodes, 1ull);
            atomicAdd(&g.stats->leaves, 1ull);
            if (!stop) { atomicAdd(&g.stats->searched, 1ull); atomicAdd(&g.stats->row_visits_mid, (unsigned long long)r); }
            atomicMax(&g.stats->max_depth, t.depth);
        }
        return;
    }

    // ---- stable partition into the other entry buffer
    const uint32_t bf = best_key >> 8, bbin = best_key & 0xFFu;
    uint32_t goL = 0;
#pragma unroll
    for (int k = 0; k < R; ++k)
        if (e[k].y && bins[(uint64_t)e[k].x * dstride + bf] <= bbin) goL |= 1u << k;
    uint32_t nLd = 0;
#pragma unroll
    for (int k = 0; k < R; ++k) nLd += __popc(__ballot_sync(0xffffffffu, (goL >> k) & 1u));
    if (nLd == 0 || nLd == r) {
        if (lane == 0) g.stats->error = 4;
        return;
    }
    uint32_t lbase = 0, rbase = 0;
#pragma unroll
    for (int k = 0; k < R; ++k) {
        const bool l = (goL >> k) & 1u, v = e[k].y != 0u;
        const uint32_t balL = __ballot_sync(0xffffffffu, l), balR = __ballot_sync(0xffffffffu, v && !l);
        if (l) ent_store(ent_out + t.begin + lbase + __popc(balL & lanemask_lt()), e[k]);
        else if (v) ent_store(ent_out + t.begin + nLd + rbase + __popc(balR & lanemask_lt()), e[k]);
        lbase += __popc(balL);
        rbase += __popc(balR);
    }
    if (lane == 0) {
        if (REG) emit_split_reg(g, t, nodes, leafc, N, reg_S(Ptot), reg_N(Pbest), reg_S(Pbest), nLd, bf, bbin, best_gain);
        else emit_split(g, t, nodes, leafc, s_tot, s_bestL, N, nLd, bf, bbin, best_gain);
        atomicAdd(&g.stats->row_visits_mid, (unsigned long long)r);
    }
}

