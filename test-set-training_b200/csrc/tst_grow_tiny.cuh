// tst_grow_tiny.cuh -- thread-per-subtree kernel for nodes with <= TMAX (8) distinct rows, <= 4 classes.
//
// Most nodes of a fully grown tree hold a handful of rows.  A warp per such node wastes most lanes on
// bookkeeping (Philox, argmax shuffles, stack traffic); here ONE THREAD owns the node and finishes its
// whole subtree, so a warp advances 32 subtrees at once.  The rows never move: they sit in fixed
// register slots and a node is a bit mask over the slots; begin / slot numbers of the children follow
// from pop counts, so node identity (begin, depth) and therefore every random draw is the same as in
// the other tiers and in oracle/hist_rf.c.  Class counts are packed 4 x 13 bits as in k_grow_small4.
// The feature draw is a sparse Fisher-Yates: only the positions touched so far are recorded.
#pragma once
#include "tst_grow_small.cuh"

#define TINY_FY_MAX 64  // positions of the draw that can be recorded (host enables the tier only for d <= 64)

// Counting sort of the tiny tasks by row count (largest first): lanes of a warp then own subtrees of the same size.
// k_tiny_hist counts the tasks per size (one atomic per warp and size), k_tiny_cursors turns the counts into bucket
// offsets, k_tiny_sort scatters.
__global__ void __launch_bounds__(256) k_tiny_hist(const uint4* __restrict__ in, uint32_t n, uint32_t* __restrict__ hist) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t r = i < n ? (in[i].w >> 16) : 0xFFFFu;
    const uint32_t peers = __match_any_sync(0xffffffffu, r);
    if (i < n && (threadIdx.x & 31u) == __ffs(peers) - 1u) atomicAdd(&hist[r], (uint32_t)__popc(peers));
}

__global__ void k_tiny_cursors(const uint32_t* __restrict__ hist, int tmax, uint32_t* __restrict__ cursor) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        uint32_t acc = 0;
        for (int r = tmax; r >= 0; --r) { cursor[r] = acc; acc += hist[r]; }
    }
}

__global__ void __launch_bounds__(256) k_tiny_sort(const uint4* __restrict__ in, uint32_t n, uint32_t* __restrict__ cursor, uint4* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < n;
    uint4 t = make_uint4(0u, 0u, 0u, 0u);
    if (live) t = in[i];
    // one atomic per (warp, size class) instead of one per task: the 7 cursors are the only contended addresses
    const uint32_t r = live ? (t.w >> 16) : 0xFFFFu;
    const uint32_t peers = __match_any_sync(0xffffffffu, r);
    const uint32_t leader = __ffs(peers) - 1u, lane = threadIdx.x & 31u;
    uint32_t base = 0;
    if (live && lane == leader) base = atomicAdd(&cursor[r], (uint32_t)__popc(peers));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (live) out[base + __popc(peers & ((1u << lane) - 1u))] = t;
}

#ifndef TINY_MINB
#define TINY_MINB 5  // 96 registers (168 unbounded, 3 CTAs per SM): 22.1 -> 19.2 ms per 640 trees; 4 (115 registers) 19.7, 6 (80, spills) 19.4
#endif
template <int TMAX, bool REG = false>  // REG: criterion 2, words are weight (13 bits) | w * yq (39 bits)
__global__ void __launch_bounds__(128, TINY_MINB) k_grow_tiny4(GrowCtx g, const uint4* __restrict__ tasks, uint32_t ntasks) {
    const uint32_t ti = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = ti < ntasks;
    const int C = g.C, d = g.d, m = g.m;
    const uint8_t* __restrict__ bins = g.bins_rm;
    const uint32_t dstride = g.dstride;
    unsigned long long visits = 0;
    uint32_t n_nodes = 0, n_leaves = 0, n_search = 0, n_fb = 0, maxd = 0;
    uint4 tk = make_uint4(0u, 0u, 0u, 0u);
    if (live) tk = tasks[ti];
    if (live && (tk.w >> 16) != 0u) {  // 0 rows: an unused slot of a warp's reservation in k_grow_small4
        const uint32_t tree = tk.w & 0xFFFFu, r = tk.w >> 16;
        const uint32_t gid = g.trees[tree].gid, arena = g.trees[tree].arena, imp_row = g.trees[tree].imp_row;
        const Entry* __restrict__ ent = ((tk.z >> 31) ? g.entB : g.entA) + (uint64_t)tree * g.n_cap + tk.x;  // bit 31: which buffer the producer parked the rows in
        uint2* nodes = g.nodes + (uint64_t)arena * g.node_stride;
        uint32_t* leafc = g.leafc + (uint64_t)arena * g.leafc_stride;
        uint32_t row[TMAX], yv[TMAX];
        unsigned long long cw[TMAX];
#pragma unroll
        for (int k = 0; k < TMAX; ++k) {
            row[k] = 0; cw[k] = 0; yv[k] = 0;
            if (k < (int)r) {
                const Entry e = ent_load(ent + k);
                row[k] = e.x;
                yv[k] = ent_cls(e);
                cw[k] = REG ? ((unsigned long long)ent_w(e) | ((unsigned long long)(ent_w(e) * ent_cls(e)) << 13)) : ((unsigned long long)ent_w(e) << (13u * ent_cls(e)));
            }
        }
        // criterion 2: a node (bit mask over the slots) is constant when its targets' min equals their max
        auto reg_const = [&](uint32_t msk, uint32_t& y_any) {
            uint32_t mn = 0xFFFFFFFFu, mx = 0u;
#pragma unroll
            for (int k = 0; k < TMAX; ++k) if ((msk >> k) & 1u) { mn = min(mn, yv[k]); mx = max(mx, yv[k]); }
            y_any = mn;
            return mn == mx;
        };
        // Every loop iteration is one split search: leaf children are written at split time and never come back
        // through the loop, so the lanes of a warp do the same kind of work (tasks are sorted by row count).
        uint32_t st_mask[TMAX], st_begin[TMAX], st_slot[TMAX], st_depth[TMAX];
        unsigned long long st_tot[TMAX];
        int sp = 0;
        uint32_t mask = (1u << r) - 1u, begin = tk.x, slot = tk.y, depth = tk.z & 0x7FFFFFFFu;
        unsigned long long tot = 0;
#pragma unroll
        for (int k = 0; k < TMAX; ++k) tot += cw[k];
        bool have = true;
        {   // the task's root was checked by its producer, except tree roots (and, for criterion 2, constant targets): check once
            const uint32_t t0 = F13(tot, 0), t1 = F13(tot, 1), t2 = F13(tot, 2), t3 = F13(tot, 3), N = REG ? t0 : t0 + t1 + t2 + t3;
            uint32_t yc = 0;
            const bool pure = REG ? reg_const(mask, yc) : (t0 == N || t1 == N || t2 == N || t3 == N);
            if (stop_rule(g, N, depth, pure)) {
                n_nodes++; n_leaves++; maxd = max(maxd, depth);
                if (REG) reg_write_leaf(nodes, leafc, slot, begin, N, REG13_S(tot), pure, yc);
                else if (pure) write_pure_leaf(nodes, slot, N, t0 == N ? 0u : (t1 == N ? 1u : (t2 == N ? 2u : 3u)));
                else {
                    for (int c = 0; c < C; ++c) leafc[(uint64_t)(begin >> 1) * C + c] = F13(tot, c);
                    write_mixed_leaf(nodes, slot, begin);
                }
                have = false;
            }
        }
        while (have) {
            const uint32_t t0 = F13(tot, 0), t1 = F13(tot, 1), t2 = F13(tot, 2), t3 = F13(tot, 3);
            const uint32_t N = REG ? t0 : t0 + t1 + t2 + t3;
            const unsigned long long Stot = REG13_S(tot);
            n_nodes++;
            maxd = max(maxd, depth);
            visits += __popc(mask);
            n_search++;
            bool found = false;
            double best_gain = -1.0;
            uint32_t best_key = 0xFFFFFFFFu;
            {
                uint8_t jpos[TINY_FY_MAX], jval[TINY_FY_MAX];
                uint4 rnd = make_uint4(0, 0, 0, 0);
                for (int i = 0; i < d; ++i) {
                    if (i >= m && found) break;
                    if ((i & 3) == 0) rnd = philox4x32_10(begin, depth, gid, TST_PURPOSE_FEAT | ((uint32_t)(i >> 2) << 8), g.k0, g.k1);
                    if (i == m) n_fb++;
                    // sparse Fisher-Yates step i: perm[i] <-> perm[j]; untouched positions hold their own index
                    const uint32_t j = (uint32_t)i + mulhi32(pick4(rnd, i & 3), (uint32_t)(d - i));
                    uint32_t vj = j, vi = (uint32_t)i;
                    for (int q = 0; q < i; ++q) {
                        const uint32_t p = jpos[q];
                        if (p == j) vj = jval[q];
                        if (p == (uint32_t)i) vi = jval[q];
                    }
                    jpos[i] = (uint8_t)j; jval[i] = (uint8_t)vi;
                    const uint32_t f = vj;
                    uint32_t b[TMAX];
#pragma unroll
                    for (int k = 0; k < TMAX; ++k) b[k] = ((mask >> k) & 1u) ? (uint32_t)bins[(uint64_t)row[k] * dstride + f] : 0xFFFFu;
                    double fg = -1.0;
                    uint32_t fkey = 0xFFFFFFFFu;
#pragma unroll
                    for (int c = 0; c < TMAX; ++c) {
                        unsigned long long acc = 0;
#pragma unroll
                        for (int k = 0; k < TMAX; ++k) if (b[k] <= b[c]) acc += cw[k] & (((mask >> k) & 1u) ? ~0ull : 0ull);
                        const uint32_t L0 = F13(acc, 0), L1 = F13(acc, 1), L2 = F13(acc, 2), L3 = F13(acc, 3);
                        const uint32_t NL = REG ? L0 : L0 + L1 + L2 + L3, NR = N - NL;
                        if (((mask >> c) & 1u) && NL >= g.msl && NR >= g.msl) {
                            const uint32_t R0 = t0 - L0, R1 = t1 - L1, R2 = t2 - L2, R3 = t3 - L3;
                            unsigned long long SL = (unsigned long long)L0 * L0 + (unsigned long long)L1 * L1 + (unsigned long long)L2 * L2 + (unsigned long long)L3 * L3;
                            unsigned long long SR = (unsigned long long)R0 * R0 + (unsigned long long)R1 * R1 + (unsigned long long)R2 * R2 + (unsigned long long)R3 * R3;
                            const double gn = REG ? mse_gain(REG13_S(acc), NL, Stot - REG13_S(acc), NR) : gini_gain(SL, NL, SR, NR);
                            const uint32_t key = (f << 8) | b[c];
                            if (gn > fg || (gn == fg && key < fkey)) { fg = gn; fkey = key; }
                        }
                    }
                    if (fkey != 0xFFFFFFFFu) {
                        if (i < m) {
                            if (fg > best_gain) { best_gain = fg; best_key = fkey; found = true; }  // ties: the feature drawn first
                        } else {
                            best_gain = fg; best_key = fkey; found = true;  // fallback: first feature in draw order that separates
                        }
                    }
                }
            }
            if (!found) {  // no feature separates the rows: a mixed leaf
                n_leaves++;
                if (REG) reg_write_leaf(nodes, leafc, slot, begin, N, Stot, false, 0u);
                else {
                    for (int c = 0; c < C; ++c) leafc[(uint64_t)(begin >> 1) * C + c] = F13(tot, c);
                    write_mixed_leaf(nodes, slot, begin);
                }
            } else {
                const uint32_t bf = best_key >> 8, bb = best_key & 0xFFu;
                uint32_t mL = 0;
                unsigned long long totL = 0;
#pragma unroll
                for (int k = 0; k < TMAX; ++k)
                    if (((mask >> k) & 1u) && (uint32_t)bins[(uint64_t)row[k] * dstride + bf] <= bb) { mL |= 1u << k; totL += cw[k]; }
                const uint32_t nLd = __popc(mL);
                if (nLd == 0 || mL == mask) { g.stats->error = 4; break; }
                const uint32_t right_slot = slot + 2u * nLd;
                write_internal(nodes, slot, right_slot, bf, bb);
                imp_add(g, imp_row, bf, best_gain, REG ? imp_parent_mse(Stot, N) : imp_parent_gini((unsigned long long)(t0 * t0 + t1 * t1 + t2 * t2 + t3 * t3), N));
                // children: leaves are finished here, the others are stacked (right first, so the left one is next)
#pragma unroll
                for (int side = 1; side >= 0; --side) {
                    const unsigned long long ct = side ? tot - totL : totL;
                    const uint32_t cm = side ? (mask & ~mL) : mL, cb = side ? begin + nLd : begin, cs = side ? right_slot : slot + 1u;
                    const uint32_t c0 = F13(ct, 0), c1 = F13(ct, 1), c2 = F13(ct, 2), c3 = F13(ct, 3), Nc = REG ? c0 : c0 + c1 + c2 + c3;
                    uint32_t yc = 0;
                    const bool cpure = REG ? reg_const(cm, yc) : (c0 == Nc || c1 == Nc || c2 == Nc || c3 == Nc);
                    if (stop_rule(g, Nc, depth + 1u, cpure)) {
                        n_nodes++; n_leaves++; maxd = max(maxd, depth + 1u);
                        if (REG) reg_write_leaf(nodes, leafc, cs, cb, Nc, REG13_S(ct), cpure, yc);
                        else if (cpure) write_pure_leaf(nodes, cs, Nc, c0 == Nc ? 0u : (c1 == Nc ? 1u : (c2 == Nc ? 2u : 3u)));
                        else {
                            for (int c = 0; c < C; ++c) leafc[(uint64_t)(cb >> 1) * C + c] = F13(ct, c);
                            write_mixed_leaf(nodes, cs, cb);
                        }
                    } else {
                        st_mask[sp] = cm; st_begin[sp] = cb; st_slot[sp] = cs; st_depth[sp] = depth + 1u; st_tot[sp] = ct;
                        ++sp;
                    }
                }
            }
            if (sp == 0) break;
            --sp;
            mask = st_mask[sp]; begin = st_begin[sp]; slot = st_slot[sp]; depth = st_depth[sp]; tot = st_tot[sp];
        }
    }
    // per-warp statistics
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        visits += __shfl_xor_sync(0xffffffffu, visits, off);
        n_nodes += __shfl_xor_sync(0xffffffffu, n_nodes, off);
        n_leaves += __shfl_xor_sync(0xffffffffu, n_leaves, off);
        n_search += __shfl_xor_sync(0xffffffffu, n_search, off);
        n_fb += __shfl_xor_sync(0xffffffffu, n_fb, off);
        maxd = max(maxd, __shfl_xor_sync(0xffffffffu, maxd, off));
    }
    if ((threadIdx.x & 31u) == 0 && n_nodes) {
        atomicAdd(&g.stats->row_visits_small, visits);
        atomicAdd(&g.stats->nodes, (unsigned long long)n_nodes);
        atomicAdd(&g.stats->leaves, (unsigned long long)n_leaves);
        atomicAdd(&g.stats->searched, (unsigned long long)n_search);
        if (n_fb) atomicAdd(&g.stats->fallback, (unsigned long long)n_fb);
        atomicMax(&g.stats->max_depth, maxd);
    }
}
