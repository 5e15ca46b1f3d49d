// tst_grow_wide.cuh -- four warps per node for nodes of 257 .. WIDE_MAXR (1024) distinct rows that weigh < 2^16 (<= 4 classes, Gini).
//
// These were the deep levels of the CTA-per-node kernel: k_grow_big spends eight barriers, a 256-thread CTA and ~20 k
// warp-instructions on a node of a few hundred rows.  Here the node's entry list is cached in shared memory once (8 bytes per
// row), the rows are fetched ONCE through a row tile (fetch_rows32, 32 rows per warp and step) and the bins at the first 8
// positions of the draw kept beside the entries (8 bytes per row), and every warp owns its sampled features END TO END: its
// own packed 16-bit histogram (warp-privatised: no cross-warp atomics), filled from shared memory, the prefix scan and the
// fp32-estimate / fp64-gain evaluation (scan_feature_packed4).  No barrier inside the search; one to combine the warps' best
// splits, two in the partition (per-warp blocks, ballots parked in shared memory).  Positions that were not staged (beyond the
// 8th, fallback draws, rows wider than 64 bytes) gather their bytes from the table, WIDE_U per lane in flight.
// 64 registers, 27 KB of shared memory and 128 threads per node: 8 CTAs per SM.
// Same draws, candidates, gains and tie-breaks as every other tier; checked against oracle/hist_rf.c.
#pragma once
#include "tst_grow_mid.cuh"

#ifndef WIDE_MAXR
#define WIDE_MAXR 1024u
#endif
#ifndef WIDE_MINB
#define WIDE_MINB 8
#endif
#ifndef WIDE_U
#define WIDE_U 4
#endif

#define WIDE_MAXCH 32u  // ballot words a warp may have to park: its block of the partition is <= WIDE_MAXR / WARPS rounded up to 32 rows
__host__ __device__ __forceinline__ size_t wide_smem_bytes(int warps, int d) {
    // histograms [warps][640] (first the warps' row tiles) | entries [WIDE_MAXR] x 8 bytes | sampled bytes [2][WIDE_MAXR] words |
    // permutation (d 16-bit values) | parked ballots [warps][WIDE_MAXCH]
    return 4u * ((size_t)warps * MID4_HIST + 4u * WIDE_MAXR + (size_t)((d + 1) / 2) + (size_t)warps * WIDE_MAXCH) + 16u;
}

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32, WIDE_MINB) k_grow_wide(GrowCtx g, const Task* __restrict__ tasks, uint32_t ntasks, const uint32_t* __restrict__ ntasks_dev) {
    constexpr int THREADS = WARPS * 32;
    constexpr uint32_t MAXCH = WIDE_MAXCH;
    static_assert(WIDE_MAXR / WARPS / 32u + 1u <= WIDE_MAXCH, "parked ballots");
    extern __shared__ __align__(16) uint32_t s_mem[];
    __shared__ double s_wgain[WARPS];
    __shared__ uint32_t s_wpos[WARPS], s_wbin[WARPS], s_wP0[WARPS], s_wP1[WARPS], s_cnt[WARPS];
    __shared__ uint32_t s_tot[4], s_bestL[4], s_T0, s_T1;

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
    const uint32_t r = t.end - t.begin;

    uint32_t* s_hist = s_mem + wid * MID4_HIST;
    Entry* s_ent = (Entry*)(s_mem + WARPS * MID4_HIST);
    uint32_t* s_samp = (uint32_t*)(s_ent + WIDE_MAXR);  // [0][i]: the bins of row i at positions 0..3 of the draw, [1][i]: positions 4..7
    uint16_t* s_perm = (uint16_t*)(s_samp + 2u * WIDE_MAXR);
    uint32_t* s_bal = (uint32_t*)(s_perm + ((d + 1) / 2) * 2) + wid * MAXCH;

    for (uint32_t i = tid; i < r; i += THREADS) s_ent[i] = ent_load(ent_in + t.begin + i);
    uint32_t draw_j = 0;
    if (wid == 0) draw_j = warp_draw_first(s_perm, d, m, t.begin, t.depth, td.gid, g.k0, g.k1, lane);
    __syncthreads();

    // ---- the rows are fetched ONCE: 32 rows per warp and step through a row tile in the warp's histogram space (16-byte
    // pieces: one L1 wavefront per row instead of one per sampled byte), and the bins at the first 8 positions of the draw are
    // kept, 8 bytes per row.  Every histogram pass and the partition then read shared memory only: no global load is left in
    // them, which is what lets the kernel run at 64 registers and 8 CTAs per SM.
    const bool staged = dstride <= 64u && dstride >= 16u && (dstride & (dstride - 1u)) == 0u;
    const uint32_t mstage = staged ? (uint32_t)min(m, 8) : 0u;  // (a fallback draw re-deals the positions >= m: never staged)
    if (staged) {
        const uint32_t pitch = (dstride >> 2) + 1u;
        uint32_t featr[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) featr[j] = s_perm[min(j, m - 1)];  // (positions >= m are not dealt: their bytes are never used)
        const uint8_t* trow = (const uint8_t*)(s_hist + lane * pitch);
        for (uint32_t c0 = wid * 32u; c0 < r; c0 += THREADS) {
            const uint32_t i = c0 + lane;
            const bool has = i < r;
            const uint32_t row = has ? s_ent[i].x : 0u;
            fetch_rows32(bins, dstride, row, has, s_hist, lane);
            uint32_t s0 = 0, s1 = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) { s0 |= (uint32_t)trow[featr[j]] << (8 * j); s1 |= (uint32_t)trow[featr[j + 4]] << (8 * j); }
            if (has) { s_samp[i] = s0; s_samp[WIDE_MAXR + i] = s1; }
            __syncwarp();
        }
        __syncthreads();
    }

    // One feature of the node (position `pos` of the draw), by one warp, into that warp's histogram.
    // (Two features per pass over the entries -- two histograms per warp -- measured slower: 76 -> 85 ms.)
    auto feature = [&](uint32_t pos, double& bg, uint32_t& bb, uint32_t& P0, uint32_t& P1, uint32_t& X0, uint32_t& X1) {
        {
            uint4* hz = (uint4*)(s_hist + lane * 20u);
            hz[0] = hz[1] = hz[2] = hz[3] = make_uint4(0u, 0u, 0u, 0u);
        }
        __syncwarp();
        if (pos < mstage) {
            const uint32_t* sp = s_samp + (pos >> 2) * WIDE_MAXR;
            const uint32_t sh = 8u * (pos & 3u);
            for (uint32_t i = lane; i < r; i += 32u) {
                const Entry q = s_ent[i];
                const uint32_t b = (sp[i] >> sh) & 0xFFu, cls = ent_cls(q);
                atomicAdd(&s_hist[(b >> 3) * 20u + (b & 7u) * 2u + (cls >> 1)], ent_w(q) << (16u * (cls & 1u)));
            }
        } else {
        // straight from the table (WIDE_U byte gathers per lane in flight): positions of the draw that were not staged --
        // beyond the 8th, fallback draws, rows wider than 64 bytes
        const uint32_t f = s_perm[pos];
        for (uint32_t i0 = 0; i0 < r; i0 += 32u * WIDE_U) {
            Entry q[WIDE_U];
            uint32_t b[WIDE_U];
#pragma unroll
            for (int u = 0; u < WIDE_U; ++u) {
                const uint32_t i = i0 + 32u * u + lane;
                q[u] = i < r ? s_ent[i] : make_uint2(0u, 0u);
            }
#pragma unroll
            for (int u = 0; u < WIDE_U; ++u) b[u] = q[u].y ? (uint32_t)bins[(uint64_t)q[u].x * dstride + f] : 0u;
#pragma unroll
            for (int u = 0; u < WIDE_U; ++u)
                if (q[u].y) {
                    const uint32_t cls = ent_cls(q[u]);
                    atomicAdd(&s_hist[(b[u] >> 3) * 20u + (b[u] & 7u) * 2u + (cls >> 1)], ent_w(q[u]) << (16u * (cls & 1u)));
                }
        }
        }
        __syncwarp();
        scan_feature_packed4(s_hist, g.msl, lane, bg, bb, P0, P1, X0, X1);
        __syncwarp();
    };

    // ---- rounds of positions of the draw, warp w takes positions lo + w, lo + w + WARPS, ...: first the m sampled features (best
    // gain; ties -> the feature drawn first, then its lowest bin); if none of them splits the node (rare), the draw continues
    // WARPS positions at a time and the first position that separates the rows wins.  One loop, so `feature` is inlined once.
    double best_gain = -1.0;
    uint32_t best_pos = 0xFFFFFFFFu, best_bin = 0, bestP0 = 0, bestP1 = 0, T0 = 0, T1 = 0;
    uint32_t t0 = 0, t1 = 0, t2 = 0, t3 = 0, N = 0;
    bool found = false, stop = false;
    for (uint32_t lo = 0, hi = (uint32_t)m;;) {
        double wg = -1.0;
        uint32_t wp = 0xFFFFFFFFu, wbin = 0, wP0 = 0, wP1 = 0;
        for (uint32_t j = lo + wid; j < hi; j += WARPS) {
            double bg; uint32_t bb, P0 = 0, P1 = 0, X0, X1;
            feature(j, bg, bb, P0, P1, X0, X1);
            T0 = X0; T1 = X1;
            if (bb != 0xFFFFFFFFu && (lo ? wp == 0xFFFFFFFFu : bg > wg)) { wg = bg; wp = j; wbin = bb; wP0 = P0; wP1 = P1; }  // earlier position keeps a tie
        }
        if (lane == 0) {
            s_wgain[wid] = wg; s_wpos[wid] = wp; s_wbin[wid] = wbin; s_wP0[wid] = wP0; s_wP1[wid] = wP1;
            if (wid == 0 && lo == 0) { s_T0 = T0; s_T1 = T1; }  // m >= 1: warp 0 always scanned position 0
        }
        __syncthreads();
        if (lo == 0) {
            T0 = s_T0; T1 = s_T1;
            t0 = T0 & 0xFFFFu; t1 = T0 >> 16; t2 = T1 & 0xFFFFu; t3 = T1 >> 16; N = t0 + t1 + t2 + t3;
            const bool pure = t0 == N || t1 == N || t2 == N || t3 == N;
            stop = stop_rule(g, N, t.depth, pure);  // only a tree root can arrive unchecked
            if (stop) break;
        }
#pragma unroll
        for (int w = 0; w < WARPS; ++w) {
            const uint32_t p = s_wpos[w];
            if (p != 0xFFFFFFFFu && (lo ? p < best_pos : (s_wgain[w] > best_gain || (s_wgain[w] == best_gain && p < best_pos)))) {
                best_gain = s_wgain[w]; best_pos = p; best_bin = s_wbin[w]; bestP0 = s_wP0[w]; bestP1 = s_wP1[w]; found = true;
            }
        }
        if (found || hi >= (uint32_t)d) break;
        lo = hi; hi = min((uint32_t)d, hi + WARPS);
        if (wid == 0 && lo == (uint32_t)m) warp_draw_materialize(s_perm, d, m, draw_j, lane);
        if (tid == 0) {  // deal the next positions (Philox block i >> 2 holds the draws of positions 4 (i >> 2) ..)
            uint4 rnd = make_uint4(0, 0, 0, 0);
            for (uint32_t i = lo; i < hi; ++i) {
                if (i == lo || (i & 3u) == 0u) rnd = philox4x32_10(t.begin, t.depth, td.gid, TST_PURPOSE_FEAT | ((i >> 2) << 8), g.k0, g.k1);
                const uint32_t jj = i + mulhi32(pick4(rnd, i & 3u), (uint32_t)d - i);
                const uint16_t fj = s_perm[jj];
                s_perm[jj] = s_perm[i];
                s_perm[i] = fj;
            }
            if (lo == (uint32_t)m) atomicAdd(&g.stats->fallback, 1ull);
        }
        __syncthreads();
    }
    if (tid == 0) {
        s_tot[0] = t0; s_tot[1] = t1; s_tot[2] = t2; s_tot[3] = t3;
        s_bestL[0] = bestP0 & 0xFFFFu; s_bestL[1] = bestP0 >> 16; s_bestL[2] = bestP1 & 0xFFFFu; s_bestL[3] = bestP1 >> 16;
    }
    if (stop || !found) {
        if (tid == 0) {
            int pc = -1;
            for (int c = 0; c < C; ++c) if (s_tot[c] == N) pc = c;
            if (pc >= 0) write_pure_leaf(nodes, t.slot, N, (uint32_t)pc);
            else {
                for (int c = 0; c < C; ++c) leafc[(uint64_t)(t.begin >> 1) * C + c] = s_tot[c];
                write_mixed_leaf(nodes, t.slot, t.begin);
            }
            atomicAdd(&g.stats->nodes, 1ull);
            atomicAdd(&g.stats->leaves, 1ull);
            if (!stop) { atomicAdd(&g.stats->searched, 1ull); atomicAdd(&g.stats->row_visits_wide, (unsigned long long)r); }
            atomicMax(&g.stats->max_depth, t.depth);
        }
        return;
    }

    // ---- stable partition: every warp owns a contiguous block; the split feature is gathered once, the ballots wait in shared memory
    const uint32_t bf = s_perm[best_pos], bbin = best_bin;
    const uint32_t rpw = ((r + THREADS - 1u) / THREADS) * 32u;
    const uint32_t blo = min(r, wid * rpw), bhi = min(r, blo + rpw);
    {
        uint32_t c = 0, k = 0;
        for (uint32_t i0 = blo; i0 < bhi; i0 += 32u, ++k) {
            const uint32_t i = i0 + lane;
            bool l;
            if (best_pos < mstage) l = i < bhi && ((s_samp[(best_pos >> 2) * WIDE_MAXR + min(i, r - 1u)] >> (8u * (best_pos & 3u))) & 0xFFu) <= bbin;
            else l = i < bhi && bins[(uint64_t)s_ent[min(i, r - 1u)].x * dstride + bf] <= bbin;
            const uint32_t bal = __ballot_sync(0xffffffffu, l);
            if (lane == 0) s_bal[k] = bal;
            c += __popc(bal);
        }
        if (lane == 0) s_cnt[wid] = c;
    }
    __syncthreads();
    uint32_t lbase = 0, nL = 0;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) { const uint32_t c = s_cnt[w]; if (w < (int)wid) lbase += c; nL += c; }
    if (nL == 0 || nL == r) {  // cannot happen for a valid split; never loop on it
        if (tid == 0) g.stats->error = 4;
        return;
    }
    {
        uint32_t rbase = blo - lbase, k = 0;
        for (uint32_t i0 = blo; i0 < bhi; i0 += 32u, ++k) {
            const uint32_t i = i0 + lane;
            const bool v = i < bhi;
            const uint32_t balL = s_bal[k], balR = ~balL & __ballot_sync(0xffffffffu, v);
            const bool l = (balL >> lane) & 1u;
            if (v) {
                const Entry e = s_ent[i];
                if (l) ent_store(ent_out + t.begin + lbase + __popc(balL & lanemask_lt()), e);
                else ent_store(ent_out + t.begin + nL + rbase + __popc(balR & lanemask_lt()), e);
            }
            lbase += __popc(balL);
            rbase += __popc(balR);
        }
    }
    __syncthreads();  // s_tot / s_bestL written by thread 0 above are read by it again: keep the CTA together until here
    if (tid == 0) {
        emit_split(g, t, nodes, leafc, s_tot, s_bestL, N, nL, bf, bbin, best_gain);
        atomicAdd(&g.stats->row_visits_wide, (unsigned long long)r);
    }
}
