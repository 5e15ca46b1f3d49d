// tst_grow_sub.cuh -- K7: one CTA finishes the WHOLE subtree under a node of <= RCAP distinct rows (SURVEY 2 inventory K7).
//
// The level-synchronous tiers re-gather a row's sampled bytes from the L2-resident table at every level and pay a
// launch / queue round trip per level.  Here the subtree's rows are fetched ONCE (warp-cooperative 16-byte pieces, one
// L1 wavefront per 64-byte row) into shared memory together with (row id, multiplicity | class); after that the subtree
// is grown without touching global memory except for the node records it writes:
//
//   phase 0  load: entries -> s_rowid / s_wc, binned rows -> s_rows (pitch = dstride/4 + 1 words: conflict-free byte reads),
//            class totals of the subtree root.
//   phase 1  nodes of > SUB_WIDE rows, CTA-cooperative: warp_draw_primary, then WARPS / m warps build the histogram of
//            each sampled feature in that feature's own packed 16-bit histogram (warp-privatised per feature), one warp
//            per feature scans it (scan_feature_packed4), the CTA partitions the node's index list in shared memory.
//   phase 2  nodes of 33 .. SUB_WIDE rows, one warp each from a shared LIFO stack (larger child pushed, smaller child
//            continued: the stack never holds more than RCAP / 33 nodes): rows' indices in registers, packed histogram
//            -> in-place prefix -> every row looks up the left counts at its own bin (candidates = occupied bins, no
//            compaction pass) -> fp32 estimate -> fp64 gain of the contenders -> REDUX arg-max; ballot partition of the
//            16-bit index list.  A child of <= 32 rows is finished on the spot by the same warp with the register-only
//            algorithm of k_grow_small4 (rows read from the shared row cache instead of the table).
//   phase 3  subtrees of <= tiny_max rows are parked in the tree's other entry buffer and handed to k_grow_tiny4 with
//            ONE queue reservation per CTA.
//
// Node identity (begin, depth) and slot numbering are the ones of the other tiers, so every Philox draw and every node
// record is the same as in oracle/hist_rf.c.  Requires <= 4 classes, Gini, rows of <= 64 bytes and a
// subtree root that weighs < 2^16 (16-bit packed class counts); everything else stays on the level-synchronous tiers.
#pragma once
#include "tst_grow_small.cuh"

#define SUB_HW 640u     // words of one packed histogram: 32 lanes x (8 bins x 2 words + 4 pad), as MID4_HIST / BIGP_FS
#define SUB_WIDE 256u   // nodes above this many rows are searched by the whole CTA, the others by one warp
#define SUB_RS 8        // row slots per lane of the warp-per-node phase (32 * 8 = SUB_WIDE rows)
#define SUB_STACK 112u  // pending nodes: <= RCAP / 33 disjoint segments of > 32 rows + <= 2 * RCAP / 257 children of phase 1
#define SUB_MAXW 32u    // wide-node stack: disjoint segments of > SUB_WIDE rows (<= RCAP / 257), generous

struct SubNode { uint32_t lo, hi, depth, slot, T0, T1; };  // [lo, hi) relative to the task's begin; T0 = c0 | c1 << 16, T1 = c2 | c3 << 16
struct SubStats { uint32_t visits, nodes, leaves, searched, fallback, maxd; };

__device__ __forceinline__ uint32_t sub_weight(uint32_t T0, uint32_t T1) { return (T0 & 0xFFFFu) + (T0 >> 16) + (T1 & 0xFFFFu) + (T1 >> 16); }
__device__ __forceinline__ int sub_pure_class(uint32_t T0, uint32_t T1, uint32_t N) {
    return (T0 & 0xFFFFu) == N ? 0 : ((T0 >> 16) == N ? 1 : ((T1 & 0xFFFFu) == N ? 2 : ((T1 >> 16) == N ? 3 : -1)));
}
__device__ __forceinline__ uint4 sub_pack(const SubNode& n, uint32_t slot0, uint32_t depth0) {
    return make_uint4(n.lo | (n.hi << 16), (n.slot - slot0) | ((n.depth - depth0) << 16), n.T0, n.T1);
}
__device__ __forceinline__ SubNode sub_unpack(const uint4& v, uint32_t slot0, uint32_t depth0) {
    SubNode n;
    n.lo = v.x & 0xFFFFu; n.hi = v.x >> 16; n.slot = slot0 + (v.y & 0xFFFFu); n.depth = depth0 + (v.y >> 16); n.T0 = v.z; n.T1 = v.w;
    return n;
}

// One thread finishes a split inside the CTA: internal record, leaf children written in place; live children are
// returned (bit 0 of `live` = left, bit 1 = right).  Same bookkeeping as emit_split (tst_grow.cuh).
__device__ __forceinline__ void sub_emit(const GrowCtx& g, uint2* nodes, uint32_t* leafc, uint32_t tbegin, const SubNode& nd, uint32_t key,
                                         uint32_t P0, uint32_t P1, uint32_t nLd, SubNode& cl, SubNode& cr, uint32_t& live, SubStats& st, bool writer,
                                         uint32_t imp_row, double gain) {
    // `writer`: this thread stores the records (a whole warp may call with uniform arguments: every lane gets the children)
    const int C = g.C;
    const uint32_t right_slot = nd.slot + 2u * nLd;
    if (writer) {
        write_internal(nodes, nd.slot, right_slot, key >> 8, key & 0xFFu);
        const uint32_t a0 = nd.T0 & 0xFFFFu, a1 = nd.T0 >> 16, a2 = nd.T1 & 0xFFFFu, a3 = nd.T1 >> 16;
        imp_add(g, imp_row, key >> 8, gain, imp_parent_gini((unsigned long long)a0 * a0 + (unsigned long long)a1 * a1 + (unsigned long long)a2 * a2 + (unsigned long long)a3 * a3, a0 + a1 + a2 + a3));
    }
    cl.lo = nd.lo; cl.hi = nd.lo + nLd; cl.depth = nd.depth + 1u; cl.slot = nd.slot + 1u; cl.T0 = P0; cl.T1 = P1;
    cr.lo = nd.lo + nLd; cr.hi = nd.hi; cr.depth = nd.depth + 1u; cr.slot = right_slot; cr.T0 = nd.T0 - P0; cr.T1 = nd.T1 - P1;  // fields never borrow: T_c >= P_c
    live = 0;
    uint32_t new_leaves = 0;
#pragma unroll
    for (int side = 0; side < 2; ++side) {
        const SubNode& c = side ? cr : cl;
        const uint32_t Nc = sub_weight(c.T0, c.T1);
        const int pc = sub_pure_class(c.T0, c.T1, Nc);
        if (stop_rule(g, Nc, c.depth, pc >= 0)) {
            if (writer) {
                if (pc >= 0) write_pure_leaf(nodes, c.slot, Nc, (uint32_t)pc);
                else {
                    const uint32_t cb = tbegin + c.lo;
                    const uint32_t cnt[4] = {c.T0 & 0xFFFFu, c.T0 >> 16, c.T1 & 0xFFFFu, c.T1 >> 16};
                    for (int k = 0; k < C; ++k) leafc[(uint64_t)(cb >> 1) * C + k] = cnt[k];
                    write_mixed_leaf(nodes, c.slot, cb);
                }
            }
            new_leaves++;
        } else {
            live |= 1u << side;
        }
    }
    st.nodes += 1u + new_leaves;
    st.leaves += new_leaves;
    st.searched += 1u;
    st.visits += nd.hi - nd.lo;
    if (new_leaves) st.maxd = max(st.maxd, nd.depth + 1u);
}

__device__ __forceinline__ void sub_leaf(const GrowCtx& g, uint2* nodes, uint32_t* leafc, uint32_t tbegin, const SubNode& nd, bool searched, SubStats& st, bool writer) {
    // a node that ends as a leaf after all: the subtree root under a stop rule, or no feature separates its rows
    const int C = g.C;
    const uint32_t N = sub_weight(nd.T0, nd.T1);
    const int pc = sub_pure_class(nd.T0, nd.T1, N);
    if (writer) {
        if (pc >= 0) write_pure_leaf(nodes, nd.slot, N, (uint32_t)pc);
        else {
            const uint32_t cb = tbegin + nd.lo;
            const uint32_t cnt[4] = {nd.T0 & 0xFFFFu, nd.T0 >> 16, nd.T1 & 0xFFFFu, nd.T1 >> 16};
            for (int k = 0; k < C; ++k) leafc[(uint64_t)(cb >> 1) * C + k] = cnt[k];
            write_mixed_leaf(nodes, nd.slot, cb);
        }
    }
    st.nodes += 1u; st.leaves += 1u;
    if (searched) { st.searched += 1u; st.visits += nd.hi - nd.lo; }
    st.maxd = max(st.maxd, nd.depth);
}

// ---------------------------------------------------------------------------------------------
// One warp: split search + partition of a node of 33 .. 32 * SUB_RS rows.  Returns false when no feature separates
// the rows.  On success idx_out[lo, hi) holds the partitioned index list and (key, P0, P1, nLd) describe the split.
__device__ __forceinline__ bool sub_narrow_node(const GrowCtx& g, uint32_t lane, const uint8_t* __restrict__ rows8, uint32_t pitch4,
                                                const uint16_t* __restrict__ s_wc, uint32_t* s_hist, uint16_t* s_perm,
                                                const uint16_t* idx_in, uint16_t* idx_out, const SubNode& nd, uint32_t abs_begin, uint32_t gid,
                                                uint32_t& out_key, uint32_t& out_P0, uint32_t& out_P1, uint32_t& out_nLd, uint32_t& n_fb, double& out_gain) {
    const int d = g.d, m = g.m;
    const uint32_t r = nd.hi - nd.lo, ns = (r + 31u) >> 5;
    uint32_t rw[SUB_RS];  // local row index | (multiplicity | class << 8) << 16; 0 = empty slot (multiplicities are >= 1)
#pragma unroll
    for (int k = 0; k < SUB_RS; ++k) {
        rw[k] = 0u;
        const uint32_t i = lane + 32u * k;
        if (i < r) {
            const uint32_t li = idx_in[nd.lo + i];
            rw[k] = li | ((uint32_t)s_wc[li] << 16);
        }
    }
    for (int f = lane; f < d; f += 32) s_perm[f] = (uint16_t)f;
    __syncwarp();
    warp_draw_primary(s_perm, d, m, abs_begin, nd.depth, gid, g.k0, g.k1, lane);
    const uint32_t t0 = nd.T0 & 0xFFFFu, t1 = nd.T0 >> 16, t2 = nd.T1 & 0xFFFFu, t3 = nd.T1 >> 16, N = t0 + t1 + t2 + t3;

    bool found = false;
    double best_gain = -1.0;
    uint32_t best_key = 0xFFFFFFFFu, bestP0 = 0, bestP1 = 0;
    uint4 rnd = make_uint4(0, 0, 0, 0);
    for (int i = 0; i < d; ++i) {
        if (i >= m && found) break;
        if (i >= m) {  // fallback: one more position of the draw
            if (lane == 0) {
                if (i == m || (i & 3) == 0)
                    rnd = philox4x32_10(abs_begin, nd.depth, gid, TST_PURPOSE_FEAT | ((uint32_t)(i >> 2) << 8), g.k0, g.k1);
                const int j = i + (int)mulhi32(pick4(rnd, i & 3), (uint32_t)(d - i));
                const uint16_t fj = s_perm[j];
                s_perm[j] = s_perm[i];
                s_perm[i] = fj;
                if (i == m) n_fb++;
            }
            __syncwarp();
        }
        const uint32_t f = s_perm[i];
        {   // the lane's own 8 bins
            uint4* hz = (uint4*)(s_hist + lane * 20u);
            hz[0] = hz[1] = hz[2] = hz[3] = make_uint4(0u, 0u, 0u, 0u);
        }
        __syncwarp();
        uint32_t bpk[2] = {0u, 0u};  // the rows' bins of this feature, byte k = slot k
#pragma unroll
        for (int k = 0; k < SUB_RS; ++k) {
            if ((uint32_t)k >= ns) break;
            if (rw[k] >> 16) {
                const uint32_t b = rows8[(rw[k] & 0xFFFFu) * pitch4 + f], wc = rw[k] >> 16;
                bpk[k >> 2] |= b << (8 * (k & 3));
                atomicAdd(&s_hist[(b >> 3) * 20u + (b & 7u) * 2u + (wc >> 9)], (wc & 0xFFu) << (16u * ((wc >> 8) & 1u)));
            }
        }
        __syncwarp();
        // in-place inclusive prefix over bins: cell(bin) := class counts of the rows with bin' <= bin
        {
            uint4* hl = (uint4*)(s_hist + lane * 20u);
            uint4 c[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) c[q] = hl[q];
            c[0].z += c[0].x; c[0].w += c[0].y;
#pragma unroll
            for (int q = 1; q < 4; ++q) { c[q].x += c[q - 1].z; c[q].y += c[q - 1].w; c[q].z += c[q].x; c[q].w += c[q].y; }
            const unsigned long long run = (unsigned long long)c[3].z | ((unsigned long long)c[3].w << 32);
            unsigned long long incl = run;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const unsigned long long a = __shfl_up_sync(0xffffffffu, incl, off);
                if (lane >= (uint32_t)off) incl += a;
            }
            const unsigned long long excl = incl - run;
            const uint32_t e0 = (uint32_t)excl, e1 = (uint32_t)(excl >> 32);
#pragma unroll
            for (int q = 0; q < 4; ++q) { c[q].x += e0; c[q].y += e1; c[q].z += e0; c[q].w += e1; hl[q] = c[q]; }
        }
        __syncwarp();
        // pass 1: fp32 estimates at the rows' own bins (the node weighs < 2^16: sums of squares fit 32 bits)
        float ge[SUB_RS], lmax = 0.0f;
#pragma unroll
        for (int k = 0; k < SUB_RS; ++k) {
            ge[k] = -1.0f;
            if ((uint32_t)k >= ns) continue;
            if (rw[k] >> 16) {
                const uint32_t b = (bpk[k >> 2] >> (8 * (k & 3))) & 0xFFu;
                const uint2 PP = *(const uint2*)(s_hist + (b >> 3) * 20u + (b & 7u) * 2u);
                const uint32_t L0 = PP.x & 0xFFFFu, L1 = PP.x >> 16, L2 = PP.y & 0xFFFFu, L3 = PP.y >> 16;
                const uint32_t NL = L0 + L1 + L2 + L3, NR = N - NL;
                if (NL >= g.msl && NR >= g.msl) {
                    const uint32_t R0 = t0 - L0, R1 = t1 - L1, R2 = t2 - L2, R3 = t3 - L3;
                    const uint32_t SL = L0 * L0 + L1 * L1 + L2 * L2 + L3 * L3, SR = R0 * R0 + R1 * R1 + R2 * R2 + R3 * R3;
                    ge[k] = gain_estimate(__uint2float_rn(SL), __uint2float_rn(NL), __uint2float_rn(SR), __uint2float_rn(NR));
                    lmax = fmaxf(lmax, ge[k]);
                }
            }
        }
        float thr = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(lmax)));  // positive floats order like integers
        if (i < m && found) thr = fmaxf(thr, __double2float_rd(best_gain));  // a feature that cannot reach the node's best is skipped
        thr *= TST_EST_MARGIN;
        // pass 2: exact fp64 gain of the contenders only
        double bg = -1.0;
        uint32_t bb = 0xFFFFFFFFu, bP0 = 0, bP1 = 0;
#pragma unroll
        for (int k = 0; k < SUB_RS; ++k) {
            if ((uint32_t)k >= ns) continue;
            if (ge[k] >= thr) {
                const uint32_t b = (bpk[k >> 2] >> (8 * (k & 3))) & 0xFFu;
                const uint2 PP = *(const uint2*)(s_hist + (b >> 3) * 20u + (b & 7u) * 2u);
                const uint32_t L0 = PP.x & 0xFFFFu, L1 = PP.x >> 16, L2 = PP.y & 0xFFFFu, L3 = PP.y >> 16;
                const uint32_t NL = L0 + L1 + L2 + L3, NR = N - NL;
                const uint32_t R0 = t0 - L0, R1 = t1 - L1, R2 = t2 - L2, R3 = t3 - L3;
                const uint32_t SL = L0 * L0 + L1 * L1 + L2 * L2 + L3 * L3, SR = R0 * R0 + R1 * R1 + R2 * R2 + R3 * R3;
                const double gn = gini_gain(SL, NL, SR, NR);
                if (gn > bg || (gn == bg && b < bb)) { bg = gn; bb = b; bP0 = PP.x; bP1 = PP.y; }
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
    if (!found) return false;

    // ---- stable partition of the index list into the other buffer
    const uint32_t bf = best_key >> 8, bbin = best_key & 0xFFu;
    uint32_t goL = 0;
#pragma unroll
    for (int k = 0; k < SUB_RS; ++k)
        if ((rw[k] >> 16) && rows8[(rw[k] & 0xFFFFu) * pitch4 + bf] <= bbin) goL |= 1u << k;
    uint32_t nLd = 0;
#pragma unroll
    for (int k = 0; k < SUB_RS; ++k) {
        if ((uint32_t)k >= ns) break;
        nLd += __popc(__ballot_sync(0xffffffffu, (goL >> k) & 1u));
    }
    uint32_t lbase = 0, rbase = 0;
#pragma unroll
    for (int k = 0; k < SUB_RS; ++k) {
        if ((uint32_t)k >= ns) break;
        const bool l = (goL >> k) & 1u, v = (rw[k] >> 16) != 0u;
        const uint32_t balL = __ballot_sync(0xffffffffu, l), balR = __ballot_sync(0xffffffffu, v && !l);
        if (l) idx_out[nd.lo + lbase + __popc(balL & lanemask_lt())] = (uint16_t)rw[k];
        else if (v) idx_out[nd.lo + nLd + rbase + __popc(balR & lanemask_lt())] = (uint16_t)rw[k];
        lbase += __popc(balL);
        rbase += __popc(balR);
    }
    __syncwarp();
    out_key = best_key; out_P0 = bestP0; out_P1 = bestP1; out_nLd = nLd; out_gain = best_gain;
    return true;
}

// ---------------------------------------------------------------------------------------------
// One warp finishes the subtree under a node of <= 32 rows (the algorithm of k_grow_small4 on the shared row cache):
// lane i holds row i, class counts packed 4 x 13 bits, sub-warp groups evaluate several sampled features at once.
// Subtrees of <= tiny_max rows are parked in `ent_park` and appended to the CTA's tiny-task list.
__device__ __forceinline__ void sub_small_subtree(const GrowCtx& g, uint32_t lane, const uint8_t* __restrict__ rows8, uint32_t pitch4,
                                                  const uint32_t* __restrict__ s_rowid, const uint16_t* __restrict__ s_wc, uint32_t* scratch,
                                                  uint16_t* s_perm, const uint16_t* idx_in, const SubNode& root, uint32_t tbegin, uint32_t gid,
                                                  Entry* __restrict__ ent_park, uint2* nodes, uint32_t* leafc, uint2* s_tiny, uint32_t* s_ntiny,
                                                  uint32_t slot0, uint32_t depth0, SubStats& st, uint32_t imp_row) {
    const int C = g.C, d = g.d, m = g.m;
    // scratch (the warp's histogram region): P[32] u64 | src[32] | stack[32][3]
    unsigned long long* s_P = (unsigned long long*)scratch;
    uint32_t* s_src = scratch + 64;
    uint32_t* s_stack = s_src + 32;
    const uint32_t r = root.hi - root.lo, base_abs = tbegin + root.lo;
    uint32_t li = 0;
    Entry e = make_uint2(0u, 0u);
    if (lane < r) {
        li = idx_in[root.lo + lane];
        e = make_uint2(s_rowid[li], (uint32_t)s_wc[li]);
    }
    uint32_t lo = 0, hi = r, depth = root.depth, slot = root.slot, sp = 0;
    while (true) {
        const bool active = lane >= lo && lane < hi;
        const uint32_t rr = hi - lo;
        const uint32_t abs_begin = base_abs + lo;
        const unsigned long long cw = active ? ((unsigned long long)ent_w(e) << (13u * ent_cls(e))) : 0ull;
        const unsigned long long tot = warp_sum64(cw);
        const uint32_t t0 = F13(tot, 0), t1 = F13(tot, 1), t2 = F13(tot, 2), t3 = F13(tot, 3);
        const uint32_t N = t0 + t1 + t2 + t3;
        const bool pure = t0 == N || t1 == N || t2 == N || t3 == N;
        const bool stop = stop_rule(g, N, depth, pure);
        if (!stop && rr <= g.tiny_max) {
            // hand the node to the thread-per-subtree tier: park its rows, remember the task
            if (active) ent_store(ent_park + abs_begin + (lane - lo), e);
            if (lane == 0) {
                const uint32_t i = atomicAdd(s_ntiny, 1u);
                s_tiny[i] = make_uint2((abs_begin - tbegin) | (rr << 16), (slot - slot0) | ((depth - depth0) << 16));
            }
            if (sp == 0) break;
            --sp;
            const uint32_t lh = s_stack[sp * 3];
            lo = lh & 0xFFu; hi = lh >> 8; depth = s_stack[sp * 3 + 1]; slot = s_stack[sp * 3 + 2];
            __syncwarp();
            continue;
        }
        st.nodes++;
        st.maxd = max(st.maxd, depth);
        bool found = false;
        double best_gain = -1.0;
        uint32_t best_key = 0xFFFFFFFFu;
        if (!stop) {
            st.visits += rr;
            st.searched++;
            const uint32_t gshift = rr <= 4 ? 2u : (rr <= 8 ? 3u : (rr <= 16 ? 4u : 5u));
            const uint32_t G = 1u << gshift, nf = 32u >> gshift;
            const uint32_t gi = lane >> gshift, sub = lane & (G - 1u), gbase = lane & ~(G - 1u);
            const bool gvalid = sub < rr;
            const uint32_t srcl = lo + (gvalid ? sub : 0u);
            const uint32_t li_g = __shfl_sync(0xffffffffu, li, srcl);
            const uint32_t ey_g = __shfl_sync(0xffffffffu, e.y, srcl);
            const unsigned long long cw_g = gvalid ? ((unsigned long long)(ey_g & 0xFFu) << (13u * (ey_g >> 8))) : 0ull;
            const uint8_t* __restrict__ rowp = rows8 + li_g * pitch4;
            for (int f = lane; f < d; f += 32) s_perm[f] = (uint16_t)f;
            __syncwarp();
            warp_draw_primary(s_perm, d, m, abs_begin, depth, gid, g.k0, g.k1, lane);
            uint4 rnd = make_uint4(0, 0, 0, 0);
            for (int pos0 = 0; pos0 < d;) {
                if (pos0 >= m && found) break;
                const int limit = pos0 < m ? m : d;
                const int nfc = min((int)nf, limit - pos0);
                if (pos0 >= m) {  // fallback: extend the draw lazily (rare)
                    if (lane == 0) {
                        for (int i = pos0; i < pos0 + nfc; ++i) {
                            if (i == pos0 || (i & 3) == 0)
                                rnd = philox4x32_10(abs_begin, depth, gid, TST_PURPOSE_FEAT | ((uint32_t)(i >> 2) << 8), g.k0, g.k1);
                            const int j = i + (int)mulhi32(pick4(rnd, i & 3), (uint32_t)(d - i));
                            const uint16_t fj = s_perm[j];
                            s_perm[j] = s_perm[i];
                            s_perm[i] = fj;
                        }
                        if (pos0 == m) st.fallback++;
                    }
                    __syncwarp();
                }
                const bool pvalid = (int)gi < nfc;
                const uint32_t pos = (uint32_t)pos0 + (pvalid ? gi : 0u);
                const uint32_t f = s_perm[pos];
                const bool lv = gvalid && pvalid;
                const uint32_t b = lv ? (uint32_t)rowp[f] : 0u;
                const unsigned long long P = cw_g | ((unsigned long long)b << 52);
                __syncwarp();
                s_P[lane] = P;
                __syncwarp();
                const uint32_t cmp = (b << 20) | 0xFFFFFu;
                unsigned long long acc = 0;
                const ulonglong2* gp = (const ulonglong2*)(s_P + gbase);
                for (uint32_t kk = 0; kk < rr; kk += 4) {
                    const ulonglong2 a = gp[kk >> 1], c = gp[(kk >> 1) + 1];
                    if ((uint32_t)(a.x >> 32) <= cmp) acc += a.x;
                    if ((uint32_t)(a.y >> 32) <= cmp) acc += a.y;
                    if ((uint32_t)(c.x >> 32) <= cmp) acc += c.x;
                    if ((uint32_t)(c.y >> 32) <= cmp) acc += c.y;
                }
                const uint32_t L0 = F13(acc, 0), L1 = F13(acc, 1), L2 = F13(acc, 2), L3 = F13(acc, 3);
                const uint32_t NL = L0 + L1 + L2 + L3, NR = N - NL;
                const uint32_t R0 = t0 - L0, R1 = t1 - L1, R2 = t2 - L2, R3 = t3 - L3;
                const bool ok = lv && NL >= g.msl && NR >= g.msl;
                const uint32_t SL = L0 * L0 + L1 * L1 + L2 * L2 + L3 * L3, SR = R0 * R0 + R1 * R1 + R2 * R2 + R3 * R3;
                double gn = -1.0;
                uint32_t key = 0xFFFFFFFFu;
                if (pos0 < m) {
                    const float ge = ok ? gain_estimate(__uint2float_rn(SL), __uint2float_rn(NL), __uint2float_rn(SR), __uint2float_rn(NR)) : 0.0f;
                    float thr = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(ge)));
                    if (found) thr = fmaxf(thr, __double2float_rd(best_gain));
                    thr *= TST_EST_MARGIN;
                    if (ok && ge >= thr) {
                        gn = gini_gain(SL, NL, SR, NR);
                        key = (pos << 8) | b;  // ties among the positions evaluated side by side: the feature drawn first, then lowest bin
                    }
                    warp_argmax_redux(gn, key);
                    if (key != 0xFFFFFFFFu && gn > best_gain) {
                        best_gain = gn; best_key = ((uint32_t)s_perm[key >> 8] << 8) | (key & 0xFFu); found = true;
                    }
                } else {
                    if (ok) {
                        gn = gini_gain(SL, NL, SR, NR);
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
            st.leaves++;
            if (pure) {
                const uint32_t pc = t0 == N ? 0u : (t1 == N ? 1u : (t2 == N ? 2u : 3u));
                if (lane == 0) write_pure_leaf(nodes, slot, N, pc);
            } else {
                if ((int)lane < C) leafc[(uint64_t)(abs_begin >> 1) * C + lane] = F13(tot, lane);
                if (lane == 0) write_mixed_leaf(nodes, slot, abs_begin);
            }
            if (sp == 0) break;
            --sp;
            const uint32_t lh = s_stack[sp * 3];
            lo = lh & 0xFFu; hi = lh >> 8; depth = s_stack[sp * 3 + 1]; slot = s_stack[sp * 3 + 2];
            __syncwarp();
            continue;
        }
        // partition inside the warp (stable)
        const uint32_t bf = best_key >> 8, bb = best_key & 0xFFu;
        const bool goL = active && rows8[li * pitch4 + bf] <= bb;
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
        li = __shfl_sync(0xffffffffu, li, src);
        const uint32_t right_slot = slot + 2u * nLd;
        if (lane == 0) {
            write_internal(nodes, slot, right_slot, bf, bb);
            imp_add(g, imp_row, bf, best_gain, imp_parent_gini((unsigned long long)(t0 * t0 + t1 * t1 + t2 * t2 + t3 * t3), N));
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

// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ size_t sub_smem_bytes(int warps, int rcap, uint32_t dstride) {
    const size_t pitch = (dstride >> 2) + 1u;
    return (size_t)rcap * pitch * 4 + (size_t)warps * SUB_HW * 4 + (size_t)rcap * 4 /* rowid */ + (size_t)rcap * 2 /* wc */ +
           2 * (size_t)rcap * 2 /* idx ping-pong */ + (size_t)warps * 64 * 2 /* perm */ + (size_t)(rcap / 2) * 8 /* tiny list */ + SUB_STACK * 16;
}

template <int WARPS, int RCAP, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB) k_grow_sub(GrowCtx g, const Task* __restrict__ tasks, uint32_t ntasks) {
    static_assert(RCAP % 32 == 0 && RCAP <= 4096, "row capacity");
    constexpr int THREADS = WARPS * 32;
    extern __shared__ __align__(16) uint32_t s_mem[];
    __shared__ uint4 s_wide[SUB_MAXW];
    __shared__ unsigned long long s_tot64;
    __shared__ uint32_t s_nwide, s_top, s_busy, s_lock, s_ntiny, s_tinybase;
    __shared__ uint32_t s_st[6];  // visits, nodes, leaves, searched, fallback, maxd
    __shared__ double s_fgain[WARPS];
    __shared__ uint32_t s_fbin[WARPS], s_fP0[WARPS], s_fP1[WARPS], s_cnt[WARPS];
    __shared__ double s_bgain;
    __shared__ uint32_t s_bkey, s_bP0, s_bP1, s_found, s_leafroot;
    __shared__ SubNode s_cur;

    if (blockIdx.x >= ntasks) return;
    const Task t = tasks[blockIdx.x];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    const int d = g.d, m = g.m;
    const TreeDesc td = g.trees[t.tree];
    const Entry* __restrict__ ent_in = ((t.depth & 1u) ? g.entB : g.entA) + (uint64_t)t.tree * g.n_cap;
    Entry* __restrict__ ent_park = ((t.depth & 1u) ? g.entA : g.entB) + (uint64_t)t.tree * g.n_cap;
    uint2* nodes = g.nodes + (uint64_t)td.arena * g.node_stride;
    uint32_t* leafc = g.leafc + (uint64_t)td.arena * g.leafc_stride;
    const uint32_t dstride = g.dstride, pitch = (dstride >> 2) + 1u, pitch4 = pitch * 4u;
    const uint32_t r = t.end - t.begin;

    uint32_t* s_rows = s_mem;
    uint32_t* s_histall = s_rows + RCAP * pitch;
    uint32_t* s_rowid = s_histall + WARPS * SUB_HW;
    uint16_t* s_wc = (uint16_t*)(s_rowid + RCAP);
    uint16_t* s_idx0 = s_wc + RCAP;
    uint16_t* s_idx1 = s_idx0 + RCAP;
    uint16_t* s_permall = s_idx1 + RCAP;
    uint2* s_tiny = (uint2*)(s_permall + WARPS * 64);
    uint4* s_stack = (uint4*)(s_tiny + RCAP / 2);
    const uint8_t* rows8 = (const uint8_t*)s_rows;
    uint32_t* s_hist = s_histall + wid * SUB_HW;
    uint16_t* s_perm = s_permall + wid * 64;

#ifdef SUB_PROFILE
    long long pc0 = clock64(), pc1 = 0, pc2 = 0, pw_narrow = 0, pw_small = 0, pw_idle = 0;
#endif
    if (tid == 0) {
        s_tot64 = 0ull; s_nwide = 0; s_top = 0; s_busy = 0; s_lock = 0; s_ntiny = 0; s_leafroot = 0;
        for (int i = 0; i < 6; ++i) s_st[i] = 0;
    }
    __syncthreads();

    // ---- phase 0: the subtree's rows, once
    {
        unsigned long long cw = 0;
        for (uint32_t base = 0; base < r; base += THREADS) {
            const uint32_t i = base + tid;
            const Entry e = i < r ? ent_load(ent_in + t.begin + i) : make_uint2(0u, 0u);
            fetch_rows32(g.bins_rm, dstride, e.x, e.y != 0u, s_rows + (base + wid * 32u) * pitch, lane);
            if (i < r) {
                s_rowid[i] = e.x;
                s_wc[i] = (uint16_t)e.y;
                s_idx0[i] = (uint16_t)i;
                cw += (unsigned long long)ent_w(e) << (16u * ent_cls(e));
            }
        }
        cw = warp_sum64(cw);
        if (lane == 0 && cw) atomicAdd(&s_tot64, cw);
    }
    __syncthreads();
    if (tid == 0) {
        SubNode root;
        root.lo = 0; root.hi = r; root.depth = t.depth; root.slot = t.slot; root.T0 = (uint32_t)s_tot64; root.T1 = (uint32_t)(s_tot64 >> 32);
        const uint32_t N = sub_weight(root.T0, root.T1);
        if (stop_rule(g, N, t.depth, sub_pure_class(root.T0, root.T1, N) >= 0)) {  // only a tree root can arrive unchecked
            SubStats st = {0, 0, 0, 0, 0, 0};
            sub_leaf(g, nodes, leafc, t.begin, root, false, st, true);
            atomicAdd(&g.stats->nodes, 1ull); atomicAdd(&g.stats->leaves, 1ull); atomicMax(&g.stats->max_depth, t.depth);
            s_leafroot = 1;
        } else if (r > SUB_WIDE) { s_wide[0] = sub_pack(root, t.slot, t.depth); s_nwide = 1; }
        else { s_stack[0] = sub_pack(root, t.slot, t.depth); s_top = 1; }
    }
    __syncthreads();
    if (s_leafroot) return;

#ifdef SUB_PROFILE
    pc1 = clock64();
#endif
    // ---- phase 1: nodes of > SUB_WIDE rows, the whole CTA per node
    while (true) {
        __syncthreads();
        if (s_nwide == 0) break;
        __syncthreads();
        if (tid == 0) {
            s_cur = sub_unpack(s_wide[--s_nwide], t.slot, t.depth);
            s_found = 0; s_bgain = -1.0; s_bkey = 0xFFFFFFFFu;
        }
        uint16_t* perm0 = s_permall;  // warp 0's permutation is the node's
        if (wid == 0) {
            for (int f = lane; f < d; f += 32) perm0[f] = (uint16_t)f;
            __syncwarp();
        }
        __syncthreads();
        const SubNode nd = s_cur;
        const uint32_t nr = nd.hi - nd.lo, abs_begin = t.begin + nd.lo;
        const uint16_t* idx_in = ((nd.depth - t.depth) & 1u) ? s_idx1 : s_idx0;
        uint16_t* idx_out = ((nd.depth - t.depth) & 1u) ? s_idx0 : s_idx1;
        if (wid == 0) warp_draw_primary(perm0, d, m, abs_begin, nd.depth, td.gid, g.k0, g.k1, lane);
        __syncthreads();
        uint4 rnd = make_uint4(0, 0, 0, 0);
        int pos = 0;
        while (pos < d) {
            if (pos >= m && s_found) break;
            const bool primary = pos < m;
            const int NF = primary ? min(m - pos, WARPS) : 1;
            if (!primary) {  // fallback: one more position of the draw per pass
                if (tid == 0) {
                    rnd = philox4x32_10(abs_begin, nd.depth, td.gid, TST_PURPOSE_FEAT | ((uint32_t)(pos >> 2) << 8), g.k0, g.k1);
                    const int j = pos + (int)mulhi32(pick4(rnd, pos & 3), (uint32_t)(d - pos));
                    const uint16_t fj = perm0[j];
                    perm0[j] = perm0[pos];
                    perm0[pos] = fj;
                    if (pos == m) s_st[4] += 1u;
                }
            }
            if ((int)wid < NF) {
                uint4* hz = (uint4*)(s_histall + wid * SUB_HW + lane * 20u);
                hz[0] = hz[1] = hz[2] = hz[3] = make_uint4(0u, 0u, 0u, 0u);
            }
            __syncthreads();
            {   // G warps share the rows of one feature; every feature has its own histogram
                const int G = WARPS / NF, j = (int)wid / G, sub = (int)wid % G;
                if (j < NF) {
                    const uint32_t f = perm0[pos + j];
                    uint32_t* hj = s_histall + j * SUB_HW;
                    for (uint32_t i = nd.lo + (uint32_t)sub * 32u + lane; i < nd.hi; i += 32u * (uint32_t)G) {
                        const uint32_t li = idx_in[i], wc = s_wc[li], b = rows8[li * pitch4 + f];
                        atomicAdd(&hj[(b >> 3) * 20u + (b & 7u) * 2u + (wc >> 9)], (wc & 0xFFu) << (16u * ((wc >> 8) & 1u)));
                    }
                }
            }
            __syncthreads();
            if ((int)wid < NF) {
                double bg; uint32_t bb, P0 = 0, P1 = 0, X0, X1;
                scan_feature_packed4(s_histall + wid * SUB_HW, g.msl, lane, bg, bb, P0, P1, X0, X1);
                if (lane == 0) { s_fgain[wid] = bg; s_fbin[wid] = bb; s_fP0[wid] = P0; s_fP1[wid] = P1; }
            }
            __syncthreads();
            if (tid == 0) {
                for (int j = 0; j < NF; ++j) {
                    if (s_fbin[j] == 0xFFFFFFFFu) continue;
                    const uint32_t key = ((uint32_t)perm0[pos + j] << 8) | s_fbin[j];
                    if (!primary || s_fgain[j] > s_bgain) {  // ties: the feature drawn first
                        s_bgain = s_fgain[j]; s_bkey = key; s_bP0 = s_fP0[j]; s_bP1 = s_fP1[j]; s_found = 1;
                        if (!primary) break;
                    }
                }
            }
            pos += NF;
            __syncthreads();
        }
        if (!s_found) {
            if (tid == 0) {
                SubStats st = {0, 0, 0, 0, 0, 0};
                sub_leaf(g, nodes, leafc, t.begin, nd, true, st, true);
                s_st[0] += st.visits; s_st[1] += st.nodes; s_st[2] += st.leaves; s_st[3] += st.searched; s_st[5] = max(s_st[5], st.maxd);
            }
            continue;
        }
        // CTA partition of idx_in[lo, hi) -> idx_out: every warp owns a contiguous block
        const uint32_t bf = s_bkey >> 8, bbin = s_bkey & 0xFFu;
        const uint32_t rpw = ((nr + THREADS - 1u) / THREADS) * 32u;
        const uint32_t blo = min(nd.hi, nd.lo + wid * rpw), bhi = min(nd.hi, blo + rpw);
        {
            uint32_t c = 0;
            for (uint32_t i0 = blo; i0 < bhi; i0 += 32u) {
                const uint32_t i = i0 + lane;
                const bool l = i < bhi && rows8[(uint32_t)idx_in[i] * pitch4 + bf] <= bbin;
                c += __popc(__ballot_sync(0xffffffffu, l));
            }
            if (lane == 0) s_cnt[wid] = c;
        }
        __syncthreads();
        uint32_t lbase = 0, nL = 0;
        for (int w = 0; w < WARPS; ++w) { const uint32_t c = s_cnt[w]; if (w < (int)wid) lbase += c; nL += c; }
        uint32_t rbase = (blo - nd.lo) - lbase;
        for (uint32_t i0 = blo; i0 < bhi; i0 += 32u) {
            const uint32_t i = i0 + lane;
            const bool v = i < bhi;
            const uint32_t li = v ? idx_in[i] : 0u;
            const bool l = v && rows8[li * pitch4 + bf] <= bbin;
            const uint32_t balL = __ballot_sync(0xffffffffu, l), balR = __ballot_sync(0xffffffffu, v && !l);
            if (l) idx_out[nd.lo + lbase + __popc(balL & lanemask_lt())] = (uint16_t)li;
            else if (v) idx_out[nd.lo + nL + rbase + __popc(balR & lanemask_lt())] = (uint16_t)li;
            lbase += __popc(balL);
            rbase += __popc(balR);
        }
        __syncthreads();
        if (tid == 0) {
            if (nL == 0 || nL == nr) g.stats->error = 4;  // cannot happen for a valid split
            else {
                SubNode cl, cr;
                uint32_t live;
                SubStats st = {0, 0, 0, 0, 0, 0};
                sub_emit(g, nodes, leafc, t.begin, nd, s_bkey, s_bP0, s_bP1, nL, cl, cr, live, st, true, td.imp_row, s_bgain);
                s_st[0] += st.visits; s_st[1] += st.nodes; s_st[2] += st.leaves; s_st[3] += st.searched; s_st[5] = max(s_st[5], st.maxd);
                for (int side = 0; side < 2; ++side)
                    if ((live >> side) & 1u) {
                        const SubNode& c = side ? cr : cl;
                        if (c.hi - c.lo > SUB_WIDE) s_wide[s_nwide++] = sub_pack(c, t.slot, t.depth);
                        else s_stack[s_top++] = sub_pack(c, t.slot, t.depth);
                    }
            }
        }
    }

#ifdef SUB_PROFILE
    pc2 = clock64();
#endif
    // ---- phase 2: nodes of <= SUB_WIDE rows, one warp each
    {
        SubStats st = {0, 0, 0, 0, 0, 0};
        while (true) {
#ifdef SUB_PROFILE
            const long long pi0 = clock64();
#endif
            uint4 pk = make_uint4(0u, 0u, 0u, 0u);
            uint32_t state = 0;  // 0 = nothing yet, 1 = got a node, 2 = all done
            if (lane == 0) {
                while (atomicCAS(&s_lock, 0u, 1u) != 0u) __nanosleep(32);
                __threadfence_block();
                const uint32_t top = *(volatile uint32_t*)&s_top;
                if (top > 0u) {
                    const volatile uint32_t* sp = (const volatile uint32_t*)(s_stack + (top - 1u));
                    pk = make_uint4(sp[0], sp[1], sp[2], sp[3]);
                    *(volatile uint32_t*)&s_top = top - 1u;
                    atomicAdd(&s_busy, 1u);  // before the lock is released: an idle warp never sees "empty and idle" while a chain is starting
                    state = 1;
                } else if (atomicAdd(&s_busy, 0u) == 0u) state = 2;
                __threadfence_block();
                atomicExch(&s_lock, 0u);
            }
            state = __shfl_sync(0xffffffffu, state, 0);
#ifdef SUB_PROFILE
            if (state != 1u) pw_idle += clock64() - pi0 + (state == 0u ? 256 * 2 : 0);
#endif
            if (state == 2u) break;
            if (state == 0u) { __nanosleep(256); continue; }
            pk.x = __shfl_sync(0xffffffffu, pk.x, 0); pk.y = __shfl_sync(0xffffffffu, pk.y, 0);
            pk.z = __shfl_sync(0xffffffffu, pk.z, 0); pk.w = __shfl_sync(0xffffffffu, pk.w, 0);
            SubNode nd = sub_unpack(pk, t.slot, t.depth);
            bool have = true;
            while (have) {
                const uint16_t* idx_in = ((nd.depth - t.depth) & 1u) ? s_idx1 : s_idx0;
                uint16_t* idx_out = ((nd.depth - t.depth) & 1u) ? s_idx0 : s_idx1;
#ifdef SUB_PROFILE
                const long long pn0 = clock64();
#endif
                if (nd.hi - nd.lo <= SMALL_MAX) {
                    sub_small_subtree(g, lane, rows8, pitch4, s_rowid, s_wc, s_hist, s_perm, idx_in, nd, t.begin, td.gid, ent_park, nodes, leafc,
                                      s_tiny, &s_ntiny, t.slot, t.depth, st, td.imp_row);
#ifdef SUB_PROFILE
                    pw_small += clock64() - pn0;
#endif
                    break;
                }
                uint32_t key = 0, P0 = 0, P1 = 0, nLd = 0, n_fb = 0;
                double gain = 0.0;
                const bool ok = sub_narrow_node(g, lane, rows8, pitch4, s_wc, s_hist, s_perm, idx_in, idx_out, nd, t.begin + nd.lo, td.gid, key, P0, P1,
                                                nLd, n_fb, gain);
#ifdef SUB_PROFILE
                pw_narrow += clock64() - pn0;
#endif
                st.fallback += n_fb;
                if (!ok) { sub_leaf(g, nodes, leafc, t.begin, nd, true, st, lane == 0); break; }
                if (nLd == 0 || nLd == nd.hi - nd.lo) { if (lane == 0) g.stats->error = 4; break; }
                SubNode cl, cr;
                uint32_t live;
                {
                    SubStats s1 = {0, 0, 0, 0, 0, 0};
                    sub_emit(g, nodes, leafc, t.begin, nd, key, P0, P1, nLd, cl, cr, live, s1, lane == 0, td.imp_row, gain);  // uniform arguments: every lane gets the children
                    st.visits += s1.visits; st.nodes += s1.nodes; st.leaves += s1.leaves; st.searched += s1.searched; st.maxd = max(st.maxd, s1.maxd);
                }
                have = false;
                if (live == 3u) {
                    const bool left_small = (cl.hi - cl.lo) <= (cr.hi - cr.lo);
                    const SubNode a = left_small ? cl : cr, b = left_small ? cr : cl;  // a = smaller
                    const uint16_t* cidx = idx_out;  // the children live in the buffer this node wrote
                    if (a.hi - a.lo <= SMALL_MAX) {
#ifdef SUB_PROFILE
                        const long long ps0 = clock64();
#endif
                        sub_small_subtree(g, lane, rows8, pitch4, s_rowid, s_wc, s_hist, s_perm, cidx, a, t.begin, td.gid, ent_park, nodes, leafc,
                                          s_tiny, &s_ntiny, t.slot, t.depth, st, td.imp_row);
#ifdef SUB_PROFILE
                        pw_small += clock64() - ps0;
#endif
                        nd = b; have = true;
                    } else {
                        if (lane == 0) {  // both > 32 rows: the larger one waits on the shared stack
                            while (atomicCAS(&s_lock, 0u, 1u) != 0u) __nanosleep(32);
                            __threadfence_block();
                            const uint32_t top = *(volatile uint32_t*)&s_top;
                            if (top < SUB_STACK) {
                                const uint4 pb = sub_pack(b, t.slot, t.depth);
                                volatile uint32_t* sp = (volatile uint32_t*)(s_stack + top);
                                sp[0] = pb.x; sp[1] = pb.y; sp[2] = pb.z; sp[3] = pb.w;
                                __threadfence_block();
                                *(volatile uint32_t*)&s_top = top + 1u;
                            }
                            else g.stats->error = 2;
                            __threadfence_block();
                            atomicExch(&s_lock, 0u);
                        }
                        __syncwarp();
                        nd = a; have = true;
                    }
                } else if (live) {
                    nd = (live & 1u) ? cl : cr;
                    have = true;
                }
            }
            __syncwarp();
            if (lane == 0) atomicSub(&s_busy, 1u);
        }
        if (lane == 0) {
            atomicAdd(&s_st[0], st.visits); atomicAdd(&s_st[1], st.nodes); atomicAdd(&s_st[2], st.leaves);
            atomicAdd(&s_st[3], st.searched); atomicAdd(&s_st[4], st.fallback); atomicMax(&s_st[5], st.maxd);
        }
    }
    __syncthreads();

#ifdef SUB_PROFILE
    {
        const long long pc3 = clock64();
        if (tid == 0) {
            atomicAdd(&g.stats->prof[0], (unsigned long long)(pc1 - pc0)); atomicAdd(&g.stats->prof[1], (unsigned long long)(pc2 - pc1));
            atomicAdd(&g.stats->prof[2], (unsigned long long)(pc3 - pc2)); atomicAdd(&g.stats->prof[6], 1ull);
        }
        if (lane == 0) {
            atomicAdd(&g.stats->prof[3], (unsigned long long)pw_narrow); atomicAdd(&g.stats->prof[4], (unsigned long long)pw_small);
            atomicAdd(&g.stats->prof[5], (unsigned long long)pw_idle);
        }
    }
#endif
    // ---- phase 3: hand the tiny subtrees to the thread-per-subtree tier with one queue reservation; flush the counters
    const uint32_t nt = s_ntiny;
    if (tid == 0) {
        s_tinybase = nt ? atomicAdd(g.qtiny_count, nt) : 0u;
        if (s_st[1]) {
            atomicAdd(&g.stats->row_visits_sub, (unsigned long long)s_st[0]);
            atomicAdd(&g.stats->nodes, (unsigned long long)s_st[1]);
            atomicAdd(&g.stats->leaves, (unsigned long long)s_st[2]);
            atomicAdd(&g.stats->searched, (unsigned long long)s_st[3]);
            if (s_st[4]) atomicAdd(&g.stats->fallback, (unsigned long long)s_st[4]);
            atomicMax(&g.stats->max_depth, s_st[5]);
        }
    }
    __syncthreads();
    if (nt) {
        const uint32_t base = s_tinybase, buf = (t.depth + 1u) & 1u;
        for (uint32_t i = tid; i < nt; i += THREADS) {
            const uint2 v = s_tiny[i];
            if (base + i < g.qtiny_cap)
                ((uint4*)g.qtiny)[base + i] = make_uint4(t.begin + (v.x & 0xFFFFu), t.slot + (v.y & 0xFFFFu), (t.depth + (v.y >> 16)) | (buf << 31), t.tree | ((v.x >> 16) << 16));
            else g.stats->error = 2;
        }
    }
}
