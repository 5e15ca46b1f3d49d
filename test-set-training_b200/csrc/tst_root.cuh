// tst_root.cuh -- dense histogram of the ROOT node of every big tree (north_star item 3, top level).
//
// The root holds all in-bag rows of its tree, so its histogram needs no row-id indirection: stream the
// column-major binned table.  Per tree and 2048-row tile the async proxy (TMA bulk copies, SASS UBLKCP)
// lands the m sampled feature columns + the tree's multiplicity bytes + the label bytes in shared memory
// through a 2-stage mbarrier pipeline; threads read 8 consecutive rows with 64-bit shared loads and feed
// a shared-memory histogram; partial histograms are merged into a per-tree global histogram with
// reductions.  Bytes moved per row: m + 2, all contiguous -- this is the one histogram kernel on the path
// that is bandwidth-bound rather than gather-latency-bound.
//
//   k_root_hist   grid (row chunks, trees)   -> ghist[tree][m][FS]
//   k_root_split  grid (trees)               -> RootSplit[tree] (best split of the root, or "use the generic path")
// k_grow_big then only partitions the pre-split roots (Task.pad0 == 1).
#pragma once
#include "tst_grow.cuh"

#define ROOT_TILE 2048u    // rows per TMA tile
#define ROOT_CHUNK 32768u  // rows per CTA

__device__ __forceinline__ void root_features(const GrowCtx& g, uint32_t gid, int d, int m, uint16_t* s_perm, uint32_t* s_feat) {
    // the node's partial Fisher-Yates draw for (begin = 0, depth = 0); one thread
    for (int f = 0; f < d; ++f) s_perm[f] = (uint16_t)f;
    uint4 rnd = make_uint4(0, 0, 0, 0);
    for (int i = 0; i < m; ++i) {
        if ((i & 3) == 0) rnd = philox4x32_10(0u, 0u, gid, TST_PURPOSE_FEAT | ((uint32_t)(i >> 2) << 8), g.k0, g.k1);
        int j = i + (int)mulhi32(pick4(rnd, i & 3), (uint32_t)(d - i));
        uint16_t fj = s_perm[j];
        s_perm[j] = s_perm[i];
        s_perm[i] = fj;
        s_feat[i] = fj;
    }
}

__global__ void __launch_bounds__(256) k_root_hist(GrowCtx g, const uint8_t* __restrict__ bins_cm, uint64_t ld, const uint8_t* __restrict__ y_pad,
                                                   const uint8_t* __restrict__ counts8, uint64_t n_rows, uint32_t* __restrict__ ghist) {
    extern __shared__ __align__(16) uint8_t s_raw[];  // hist [m][FS] u32 | 2 stages x (m + 2) x ROOT_TILE bytes
    __shared__ uint16_t s_perm[TST_MAX_FEATURES];
    __shared__ uint32_t s_feat[8];
    __shared__ __align__(8) uint64_t s_bar[2];
    const uint32_t tid = threadIdx.x;
    const int C = g.C, d = g.d, m = g.m;
    const TreeDesc td = g.trees[blockIdx.y];
    if (td.n_inbag <= g.big_min) return;  // not a big root
    uint32_t* s_hist = (uint32_t*)s_raw;
    uint8_t* s_stage = s_raw + (size_t)m * g.FS * 4;
    const uint32_t stage_bytes = (uint32_t)(m + 2) * ROOT_TILE;
    const uint64_t row_lo = (uint64_t)blockIdx.x * ROOT_CHUNK;
    if (row_lo >= n_rows) return;
    const uint64_t row_hi = min(n_rows, row_lo + ROOT_CHUNK);
    const uint32_t tiles = (uint32_t)((row_hi - row_lo + ROOT_TILE - 1) / ROOT_TILE);
    if (tid == 0) {
        root_features(g, td.gid, d, m, s_perm, s_feat);
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        mbar_fence_init();
    }
    for (int i = tid; i < m * g.FS; i += 256) s_hist[i] = 0;
    __syncthreads();
    const uint8_t* cnt = counts8 + (uint64_t)blockIdx.y * ld;  // this tree's multiplicity bytes (n_pad == ld)
    auto issue = [&](uint32_t s, uint32_t tile) {
        const uint64_t r0 = row_lo + (uint64_t)tile * ROOT_TILE;
        const uint32_t bytes = (uint32_t)min((uint64_t)ROOT_TILE, (ld - r0));  // ld is a multiple of 16; rows >= n carry multiplicity 0
        uint8_t* dst = s_stage + (size_t)s * stage_bytes;
        mbar_expect_tx(&s_bar[s], bytes * (uint32_t)(m + 2));
        for (int j = 0; j < m; ++j) tma_load_1d(dst + (size_t)j * ROOT_TILE, bins_cm + (uint64_t)s_feat[j] * ld + r0, bytes, &s_bar[s]);
        tma_load_1d(dst + (size_t)m * ROOT_TILE, cnt + r0, bytes, &s_bar[s]);
        tma_load_1d(dst + (size_t)(m + 1) * ROOT_TILE, y_pad + r0, bytes, &s_bar[s]);
    };
    if (tid == 0) {
        issue(0, 0);
        if (tiles > 1) issue(1, 1);
    }
    for (uint32_t i = 0; i < tiles; ++i) {
        const uint32_t s = i & 1u;
        mbar_wait(&s_bar[s], (i >> 1) & 1u);
        const uint8_t* st = s_stage + (size_t)s * stage_bytes;
        const uint64_t r0 = row_lo + (uint64_t)i * ROOT_TILE;
        const uint32_t valid = (uint32_t)min((uint64_t)ROOT_TILE, ld - r0);
        if (tid * 8u < valid) {
            const uint2 mw = *(const uint2*)(st + (size_t)m * ROOT_TILE + tid * 8u);  // 8 multiplicities
            if (mw.x | mw.y) {
                const uint2 yw = *(const uint2*)(st + (size_t)(m + 1) * ROOT_TILE + tid * 8u);
                for (int j = 0; j < m; ++j) {
                    const uint2 bw = *(const uint2*)(st + (size_t)j * ROOT_TILE + tid * 8u);  // 128-byte-coalesced column bytes
                    uint32_t* hj = s_hist + j * g.FS;
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const uint32_t w = ((q < 4 ? mw.x : mw.y) >> (8 * (q & 3))) & 0xFFu;
                        if (w) {
                            const uint32_t b = ((q < 4 ? bw.x : bw.y) >> (8 * (q & 3))) & 0xFFu;
                            const uint32_t cls = ((q < 4 ? yw.x : yw.y) >> (8 * (q & 3))) & 0xFFu;
                            atomicAdd(&hj[hist_index(b, cls, C)], w);
                        }
                    }
                }
            }
        }
        __syncthreads();  // every thread is done with stage s
        if (tid == 0 && i + 2 < tiles) issue(s, i + 2);
    }
    uint32_t* gh = ghist + (uint64_t)blockIdx.y * (uint64_t)(m * g.FS);
    for (int i = tid; i < m * g.FS; i += 256) {
        const uint32_t v = s_hist[i];
        if (v) atomicAdd(&gh[i], v);
    }
}

__global__ void __launch_bounds__(256) k_root_split(GrowCtx g, const uint32_t* __restrict__ ghist, RootSplit* __restrict__ out) {
    extern __shared__ uint32_t s_hist[];  // [m][FS]
    __shared__ uint16_t s_perm[TST_MAX_FEATURES];
    __shared__ uint32_t s_feat[8];
    __shared__ double s_fgain[8];
    __shared__ uint32_t s_fbin[8];
    __shared__ uint32_t s_tot[TST_MAX_CLASSES];
    __shared__ uint32_t s_N, s_bj, s_bkey;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    const int C = g.C, d = g.d, m = g.m;
    const TreeDesc td = g.trees[blockIdx.x];
    RootSplit* rs = out + blockIdx.x;
    if (td.n_inbag <= g.big_min) { if (tid == 0) rs->state = 0; return; }
    const uint32_t* gh = ghist + (uint64_t)blockIdx.x * (uint64_t)(m * g.FS);
    for (int i = tid; i < m * g.FS; i += 256) s_hist[i] = gh[i];
    if (tid == 0) root_features(g, td.gid, d, m, s_perm, s_feat);
    __syncthreads();
    for (int j = wid; j < m; j += 8) {
        double bg; uint32_t bb;
        scan_feature(s_hist + j * g.FS, C, g.msl, lane, bg, bb, j == 0 ? s_tot : nullptr, j == 0 ? &s_N : nullptr, g.criterion);
        if (lane == 0) { s_fgain[j] = bg; s_fbin[j] = bb; }
    }
    __syncthreads();
    if (tid == 0) {
        bool pure = false;
        for (int c = 0; c < C; ++c) pure |= (s_tot[c] == s_N);
        uint32_t state = 3, key = 0xFFFFFFFFu, bj = 0;
        double best = -1.0;
        if (stop_rule(g, s_N, 0u, pure)) state = 2;
        else
            for (int j = 0; j < m; ++j) {
                if (s_fbin[j] == 0xFFFFFFFFu) continue;
                const uint32_t k2 = (s_feat[j] << 8) | s_fbin[j];
                if (s_fgain[j] > best) { best = s_fgain[j]; key = k2; bj = j; state = 1; }  // ties: the feature drawn first
            }
        rs->state = state; rs->key = key; rs->N = s_N;
        for (int c = 0; c < C; ++c) rs->tot[c] = s_tot[c];
        s_bj = bj; s_bkey = key;
    }
    __syncthreads();
    if (s_bkey != 0xFFFFFFFFu) {
        const uint32_t bb = s_bkey & 0xFFu;
        const uint32_t* hl = s_hist + s_bj * g.FS + lane * (8u * C + 1u);
        for (int c = wid; c < C; c += 8) {
            uint32_t v = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) v += (lane * 8u + q <= bb) ? hl[q * C + c] : 0u;
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
            if (lane == 0) rs->bestL[c] = v;
        }
    }
}
