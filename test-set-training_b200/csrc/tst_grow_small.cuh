// tst_grow_small.cuh -- warp-per-subtree kernel for nodes with <= 32 distinct rows, <= 4 classes.
//
// Same arithmetic as k_grow_small (tst_grow.cuh) but register-only:
//   * class counts are packed 4 x 13 bits in one 64-bit word (a node of <= 32 rows x multiplicity
//     <= 255 weighs <= 8160 < 2^13), the row's bin rides in bits 52..59, so "weights left of my
//     threshold" is one compare + one 64-bit add per broadcast row;
//   * the warp is cut into groups of G = 4 / 8 / 16 / 32 lanes (G >= rows in the node) and every
//     group evaluates a different sampled feature, so a 4-row node evaluates 8 features at once.
// Bit-exact with oracle/hist_rf.c: candidates are the occupied bins, gain in fp64, ties -> lowest
// (feature, bin); fallback draws take the first feature in draw order that separates the rows.
#pragma once
#include <type_traits>
#include "tst_grow.cuh"

__device__ __forceinline__ unsigned long long warp_sum64(unsigned long long v) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

__device__ __forceinline__ void warp_argmax(double& gn, uint32_t& key) {
    // (gain desc, key asc); every lane ends with the winner
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        double og = __shfl_xor_sync(0xffffffffu, gn, off);
        uint32_t ok = __shfl_xor_sync(0xffffffffu, key, off);
        if (og > gn || (og == gn && ok < key)) { gn = og; key = ok; }
    }
}

#define F13(x, c) ((uint32_t)((x) >> (13 * (c))) & 0x1FFFu)
// criterion 2: the 64-bit word of a row is weight (13 bits) | w * yq (39 bits) [| bin << 52 while comparing]
#define REG13_S(x) (((x) >> 13) & 0x7FFFFFFFFFull)

#ifndef SMALL_MINB
#define SMALL_MINB 7  // 72 registers: measured 117.4 ms per 640 trees vs 121.8 (8 CTAs, 64 registers, spills), 120 (6), 122.6 (5)
#endif
#define TINY_RESERVE 8u  // tiny-queue slots a warp of k_grow_small4 reserves per atomic
__host__ __device__ constexpr uint32_t small4_per_warp(int d) { return (256u + 32u + 96u + (uint32_t)((d + 1) / 2) + 3u) & ~3u; }

// REG: criterion 2 (continuous squared error), same register-only scheme on (weight, sum) words.
// MULTI: several positions of the draw per lane group and pass (the throughput instantiation: its extra live values cost the
// one-position path 10 % in latency-bound jobs, so the host picks it for deep batches and for criterion 2 only).
template <int WARPS, bool REG = false, bool MULTI = false>
__global__ void __launch_bounds__(WARPS * 32, SMALL_MINB) k_grow_small4(GrowCtx g, const Task* __restrict__ tasks, uint32_t ntasks) {
    extern __shared__ __align__(16) uint32_t s_mem[];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const int C = g.C, d = g.d, m = g.m;
    // per-warp shared layout (words): P[4][32] u64 | src[32] | stack[32][3] | perm[(d+1)/2], rounded to 16 bytes
    const uint32_t per_warp = small4_per_warp(d);
    unsigned long long* s_P = (unsigned long long*)(s_mem + warp * per_warp);
    uint32_t* s_src = s_mem + warp * per_warp + 256;
    uint32_t* s_stack = s_src + 32;
    uint16_t* s_perm = (uint16_t*)(s_stack + 96);
    const uint8_t* __restrict__ bins = g.bins_rm;
    const uint32_t dstride = g.dstride;

    // Work counters and tiny-task slots are kept per warp across its tasks: one atomic per 8 pushed tasks and six at
    // the end instead of one per push / six per task (the queue counter is a single contended address).
    unsigned long long visits = 0, n_nodes = 0, n_leaves = 0, n_search = 0, n_fb = 0;
    uint32_t maxd = 0;
    uint32_t res_i = 0, res_end = 0;  // lane 0: reserved tiny-queue slots [res_i, res_end)
    for (uint32_t ti = blockIdx.x * WARPS + warp; ti < ntasks; ti += gridDim.x * WARPS) {
        const Task t = tasks[ti];
        const TreeDesc td = g.trees[t.tree];
        const Entry* __restrict__ ent = ((t.depth & 1u) ? g.entB : g.entA) + (uint64_t)t.tree * g.n_cap;
        uint2* nodes = g.nodes + (uint64_t)td.arena * g.node_stride;
        uint32_t* leafc = g.leafc + (uint64_t)td.arena * g.leafc_stride;
        const uint32_t r = t.end - t.begin;
        Entry e = lane < r ? ent_load(ent + t.begin + lane) : make_uint2(0u, 0u);
        uint32_t lo = 0, hi = r, depth = t.depth, slot = t.slot, sp = 0;
        while (true) {
            const bool active = lane >= lo && lane < hi;
            const uint32_t rr = hi - lo;
            const uint32_t abs_begin = t.begin + lo;
            const unsigned long long cw = !active ? 0ull : REG ? ((unsigned long long)ent_w(e) | ((unsigned long long)(ent_w(e) * ent_cls(e)) << 13))
                                                               : ((unsigned long long)ent_w(e) << (13u * ent_cls(e)));
            const unsigned long long tot = warp_sum64(cw);
            const uint32_t t0 = F13(tot, 0), t1 = F13(tot, 1), t2 = F13(tot, 2), t3 = F13(tot, 3);
            const uint32_t N = REG ? t0 : t0 + t1 + t2 + t3;
            const unsigned long long Stot = REG13_S(tot);
            uint32_t ymn = 0, ymx = 0;
            if (REG) {
                ymn = __reduce_min_sync(0xffffffffu, active ? ent_cls(e) : 0xFFFFFFFFu);
                ymx = __reduce_max_sync(0xffffffffu, active ? ent_cls(e) : 0u);
            }
            const bool pure = REG ? ymn == ymx : (t0 == N || t1 == N || t2 == N || t3 == N);
            const bool stop = stop_rule(g, N, depth, pure);
            if (!stop && rr <= g.tiny_max) {
                // hand the node to the thread-per-subtree tier: park its rows in the entry buffer of its depth
                Entry* eb = ((depth & 1u) ? g.entB : g.entA) + (uint64_t)t.tree * g.n_cap;
                if (active) ent_store(eb + t.begin + lane, e);
                if (lane == 0) {
                    if (res_i == res_end) { res_i = atomicAdd(g.qtiny_count, TINY_RESERVE); res_end = res_i + TINY_RESERVE; }
                    if (res_i < g.qtiny_cap) ((uint4*)g.qtiny)[res_i] = make_uint4(abs_begin, slot, depth | ((depth & 1u) << 31), t.tree | (rr << 16));
                    else g.stats->error = 2;
                    ++res_i;
                }
                if (sp == 0) break;
                --sp;
                uint32_t lh = s_stack[sp * 3];
                lo = lh & 0xFFu; hi = lh >> 8; depth = s_stack[sp * 3 + 1]; slot = s_stack[sp * 3 + 2];
                __syncwarp();
                continue;
            }
            n_nodes++;
            maxd = max(maxd, depth);
            bool found = false;
            double best_gain = -1.0;
            uint32_t best_key = 0xFFFFFFFFu;
            if (!stop) {
                visits += rr;
                n_search++;
                // group layout: G lanes per feature
                const uint32_t gshift = rr <= 4 ? 2u : (rr <= 8 ? 3u : (rr <= 16 ? 4u : 5u));
                const uint32_t G = 1u << gshift, nf = 32u >> gshift;
                const uint32_t gi = lane >> gshift, sub = lane & (G - 1u), gbase = lane & ~(G - 1u);
                const bool gvalid = sub < rr;
                const uint32_t srcl = lo + (gvalid ? sub : 0u);
                Entry e_g;
                e_g.x = __shfl_sync(0xffffffffu, e.x, srcl);
                e_g.y = __shfl_sync(0xffffffffu, e.y, srcl);
                const unsigned long long cw_g = !gvalid ? 0ull : REG ? ((unsigned long long)ent_w(e_g) | ((unsigned long long)(ent_w(e_g) * ent_cls(e_g)) << 13))
                                                                     : ((unsigned long long)ent_w(e_g) << (13u * ent_cls(e_g)));
                const uint8_t* __restrict__ rowp = bins + (uint64_t)e_g.x * dstride;
                const uint32_t draw_j = warp_draw_first(s_perm, d, m, abs_begin, depth, td.gid, g.k0, g.k1, lane);
                // bin <= b  <=>  high word of P <= (b << 20) | 0xFFFFF (bits 32..51 hold class counts)
                uint4 rnd = make_uint4(0, 0, 0, 0);
                for (int pos0 = 0; pos0 < d;) {
                    if (pos0 >= m && found) break;
                    const int limit = pos0 < m ? m : d;
                    const int nfc = min((int)nf, limit - pos0);
                    if (pos0 >= m) {  // fallback: extend the draw lazily (rare)
                        if (pos0 == m) warp_draw_materialize(s_perm, d, m, draw_j, lane);
                        if (lane == 0) {
                            for (int i = pos0; i < pos0 + nfc; ++i) {
                                if (i == pos0 || (i & 3) == 0)
                                    rnd = philox4x32_10(abs_begin, depth, td.gid, TST_PURPOSE_FEAT | ((uint32_t)(i >> 2) << 8), g.k0, g.k1);
                                int j = i + (int)mulhi32(pick4(rnd, i & 3), (uint32_t)(d - i));
                                uint16_t fj = s_perm[j];
                                s_perm[j] = s_perm[i];
                                s_perm[i] = fj;
                            }
                            if (pos0 == m) n_fb++;
                        }
                        __syncwarp();
                    }
                    if constexpr (MULTI) {
                        // T = 2 or 4 positions per group and pass while the sampled features outnumber the groups (nodes of > 8 rows
                        // with m = 8): group gi takes positions pos0 + q nf + gi, q < T.  One broadcast loop, one estimate REDUX and one
                        // arg-max serve them all, which divides the fixed cost of a pass (the 17..32-row nodes ran 8 passes, now 2).
                        auto multi = [&](auto TT) {
                            constexpr int T = decltype(TT)::value;
                            const int nfcT = min(T * (int)nf, m - pos0);
                            uint32_t posv[T], bv[T];
                            bool lv[T];
#pragma unroll
                            for (int q = 0; q < T; ++q) {
                                const bool v = (int)(q * nf + gi) < nfcT;
                                posv[q] = (uint32_t)pos0 + (v ? q * nf + gi : 0u);
                                lv[q] = gvalid && v;
                                bv[q] = lv[q] ? (uint32_t)rowp[s_perm[posv[q]]] : 0u;
                            }
                            __syncwarp();
#pragma unroll
                            for (int q = 0; q < T; ++q) s_P[32u * q + lane] = cw_g | ((unsigned long long)bv[q] << 52);
                            __syncwarp();
                            unsigned long long acc[T];
#pragma unroll
                            for (int q = 0; q < T; ++q) acc[q] = 0ull;
                            for (uint32_t kk = 0; kk < rr; kk += 4) {
#pragma unroll
                                for (int q = 0; q < T; ++q) {
                                    const ulonglong2* gp = (const ulonglong2*)(s_P + 32u * q + gbase);
                                    const ulonglong2 a = gp[kk >> 1], c = gp[(kk >> 1) + 1];
                                    const uint32_t cmp = (bv[q] << 20) | 0xFFFFFu;
                                    if ((uint32_t)(a.x >> 32) <= cmp) acc[q] += a.x;
                                    if ((uint32_t)(a.y >> 32) <= cmp) acc[q] += a.y;
                                    if ((uint32_t)(c.x >> 32) <= cmp) acc[q] += c.x;
                                    if ((uint32_t)(c.y >> 32) <= cmp) acc[q] += c.y;
                                }
                            }
                            float ge[T], gmax = 0.0f;
#pragma unroll
                            for (int q = 0; q < T; ++q) {
                                if constexpr (REG) {
                                    const uint32_t NL = F13(acc[q], 0), NR = N - NL;
                                    const unsigned long long SLr = REG13_S(acc[q]), SRr = Stot - SLr;
                                    lv[q] = lv[q] && NL >= g.msl && NR >= g.msl;
                                    ge[q] = lv[q] ? mse_gain_estimate(__ull2float_rn(SLr), __uint2float_rn(NL), __ull2float_rn(SRr), __uint2float_rn(NR)) : 0.0f;
                                } else {
                                    const uint32_t L0 = F13(acc[q], 0), L1 = F13(acc[q], 1), L2 = F13(acc[q], 2), L3 = F13(acc[q], 3);
                                    const uint32_t NL = L0 + L1 + L2 + L3, NR = N - NL;
                                    const uint32_t R0 = t0 - L0, R1 = t1 - L1, R2 = t2 - L2, R3 = t3 - L3;
                                    const uint32_t SL = L0 * L0 + L1 * L1 + L2 * L2 + L3 * L3, SR = R0 * R0 + R1 * R1 + R2 * R2 + R3 * R3;
                                    lv[q] = lv[q] && NL >= g.msl && NR >= g.msl;
                                    ge[q] = lv[q] ? gain_estimate(__uint2float_rn(SL), __uint2float_rn(NL), __uint2float_rn(SR), __uint2float_rn(NR)) : 0.0f;
                                }
                                gmax = fmaxf(gmax, ge[q]);
                            }
                            float thr = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(gmax)));
                            if (found) thr = fmaxf(thr, __double2float_rd(best_gain));
                            thr *= TST_EST_MARGIN;
                            double gn = -1.0;
                            uint32_t key = 0xFFFFFFFFu;
#pragma unroll
                            for (int q = 0; q < T; ++q) {
                                if (lv[q] && ge[q] >= thr) {
                                    double gq;
                                    if constexpr (REG) {
                                        const uint32_t NL = F13(acc[q], 0), NR = N - NL;
                                        const unsigned long long SLr = REG13_S(acc[q]);
                                        gq = mse_gain(SLr, NL, Stot - SLr, NR);
                                    } else {
                                        const uint32_t L0 = F13(acc[q], 0), L1 = F13(acc[q], 1), L2 = F13(acc[q], 2), L3 = F13(acc[q], 3);
                                        const uint32_t NL = L0 + L1 + L2 + L3, NR = N - NL;
                                        const uint32_t R0 = t0 - L0, R1 = t1 - L1, R2 = t2 - L2, R3 = t3 - L3;
                                        const uint32_t SL = L0 * L0 + L1 * L1 + L2 * L2 + L3 * L3, SR = R0 * R0 + R1 * R1 + R2 * R2 + R3 * R3;
                                        gq = gini_gain(SL, NL, SR, NR);
                                    }
                                    if (gq > gn) { gn = gq; key = (posv[q] << 8) | bv[q]; }  // positions ascend with q: a tie stays with the feature drawn first
                                }
                            }
                            warp_argmax_redux(gn, key);
                            if (key != 0xFFFFFFFFu && gn > best_gain) {  // an earlier pass keeps a tie
                                best_gain = gn; best_key = ((uint32_t)s_perm[key >> 8] << 8) | (key & 0xFFu); found = true;
                            }
                            pos0 += nfcT;
                        };
                        if (pos0 < m && 2 * (int)nf < m - pos0) { multi(std::integral_constant<int, 4>{}); continue; }
                        if (pos0 < m && (int)nf < m - pos0) { multi(std::integral_constant<int, 2>{}); continue; }
                    }
                    const bool pvalid = (int)gi < nfc;
                    const uint32_t pos = (uint32_t)pos0 + (pvalid ? gi : 0u);
                    const uint32_t f = s_perm[pos];
                    const bool lv = gvalid && pvalid;
                    const uint32_t b = lv ? (uint32_t)rowp[f] : 0u;
                    const unsigned long long P = cw_g | ((unsigned long long)b << 52);
                    // weights left of my threshold: every lane of the group reads the group's rows from shared memory
                    // (two rows per 128-bit broadcast load); lanes past the node's rows hold P = 0 and add nothing
                    __syncwarp();
                    s_P[lane] = P;
                    __syncwarp();
                    const uint32_t cmp = (b << 20) | 0xFFFFFu;
                    unsigned long long acc = 0;
                    const ulonglong2* gp = (const ulonglong2*)(s_P + gbase);
                    for (uint32_t kk = 0; kk < rr; kk += 4) {  // G >= 4 is a multiple of 4: never reads past the group
                        const ulonglong2 a = gp[kk >> 1], c = gp[(kk >> 1) + 1];
                        if ((uint32_t)(a.x >> 32) <= cmp) acc += a.x;
                        if ((uint32_t)(a.y >> 32) <= cmp) acc += a.y;
                        if ((uint32_t)(c.x >> 32) <= cmp) acc += c.x;
                        if ((uint32_t)(c.y >> 32) <= cmp) acc += c.y;
                    }
                    const uint32_t L0 = F13(acc, 0), L1 = F13(acc, 1), L2 = F13(acc, 2), L3 = F13(acc, 3);
                    const uint32_t NL = REG ? L0 : L0 + L1 + L2 + L3, NR = N - NL;
                    const uint32_t R0 = t0 - L0, R1 = t1 - L1, R2 = t2 - L2, R3 = t3 - L3;
                    const bool ok = lv && NL >= g.msl && NR >= g.msl;
                    // a node here weighs < 2^13: the sums of squares fit 32 bits
                    const uint32_t SL = L0 * L0 + L1 * L1 + L2 * L2 + L3 * L3, SR = R0 * R0 + R1 * R1 + R2 * R2 + R3 * R3;
                    const unsigned long long SLr = REG13_S(acc), SRr = Stot - SLr;  // criterion 2
                    double gn = -1.0;
                    uint32_t key = 0xFFFFFFFFu;
                    if (pos0 < m) {
                        // fp32 estimate first; the exact fp64 gain only where it can still win (tst_common.cuh: gain_estimate)
                        const float ge = !ok ? 0.0f : REG ? mse_gain_estimate(__ull2float_rn(SLr), __uint2float_rn(NL), __ull2float_rn(SRr), __uint2float_rn(NR))
                                                          : gain_estimate(__uint2float_rn(SL), __uint2float_rn(NL), __uint2float_rn(SR), __uint2float_rn(NR));
                        float thr = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(ge)));
                        if (found) thr = fmaxf(thr, __double2float_rd(best_gain));
                        thr *= TST_EST_MARGIN;
                        if (ok && ge >= thr) {
                            gn = REG ? mse_gain(SLr, NL, SRr, NR) : gini_gain(SL, NL, SR, NR);
                            key = (pos << 8) | b;  // ties among the positions evaluated side by side: the feature drawn first, then lowest bin
                        }
                        warp_argmax_redux(gn, key);
                        if (key != 0xFFFFFFFFu && gn > best_gain) {  // ... and an earlier pass keeps a tie
                            best_gain = gn; best_key = ((uint32_t)s_perm[key >> 8] << 8) | (key & 0xFFu); found = true;
                        }
                    } else {
                        // fallback: the first position in draw order that has any valid split
                        if (ok) {
                            gn = REG ? mse_gain(SLr, NL, SRr, NR) : gini_gain(SL, NL, SR, NR);
                            key = (f << 8) | b;
                        }
                        uint32_t mp = ok ? pos : 0xFFFFFFFFu;
#pragma unroll
                        for (int off = 16; off >= 1; off >>= 1) mp = min(mp, __shfl_xor_sync(0xffffffffu, mp, off));
                        if (mp != 0xFFFFFFFFu) {
                            if (!(ok && pos == mp)) { gn = -1.0; key = 0xFFFFFFFFu; }
                            warp_argmax(gn, key);
                            best_gain = gn; best_key = key; found = true;
                        }
                    }
                    pos0 += nfc;
                }
            }
            if (!found) {
                n_leaves++;
                if (REG) {
                    if (lane == 0) reg_write_leaf(nodes, leafc, slot, abs_begin, N, Stot, pure, ymn);
                } else if (pure) {
                    const uint32_t pc = t0 == N ? 0u : (t1 == N ? 1u : (t2 == N ? 2u : 3u));
                    if (lane == 0) write_pure_leaf(nodes, slot, N, pc);
                } else {
                    if ((int)lane < C) leafc[(uint64_t)(abs_begin >> 1) * C + lane] = F13(tot, lane);
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
            if (nLd == 0 || nLd == rr) {  // cannot happen for a valid split; never loop on it
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
            const uint32_t right_slot = slot + 2u * nLd;
            if (lane == 0) {
                write_internal(nodes, slot, right_slot, bf, bb);
                imp_add(g, td.imp_row, bf, best_gain, REG ? imp_parent_mse(Stot, N) : imp_parent_gini((unsigned long long)(t0 * t0 + t1 * t1 + t2 * t2 + t3 * t3), N));
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
        __syncwarp();
    }
    if (lane == 0) {
        for (; res_i < res_end; ++res_i)  // unused reserved slots become empty tasks (0 rows): sorted last, skipped by k_grow_tiny4
            if (res_i < g.qtiny_cap) ((uint4*)g.qtiny)[res_i] = make_uint4(0u, 0u, 0u, 0u);
        if (n_nodes) {
            atomicAdd(&g.stats->row_visits_small, visits);
            atomicAdd(&g.stats->nodes, n_nodes);
            atomicAdd(&g.stats->leaves, n_leaves);
            atomicAdd(&g.stats->searched, n_search);
            if (n_fb) atomicAdd(&g.stats->fallback, n_fb);
            atomicMax(&g.stats->max_depth, maxd);
        }
    }
}
