// tst_quantile.cuh -- bin-edge selection on the device (north_star item 1, "quantile binning").
//
// Rule = DESIGN.md 2.1 (SK/ensemble/_hist_gradient_boosting/binning.py:31-111 with exact integer ranks):
// per feature, sort the non-missing values; <= B distinct -> midpoints of consecutive distinct values,
// otherwise the B-1 interior 'averaged_inverted_cdf' quantiles, duplicates removed.  Bit-identical to the
// host rule in tst/binning.py (order statistics are exact; a midpoint of two floats is exact in double).
//
//   k_keys_transpose  X [n][d] f32 row-major -> order-preserving uint32 keys, column-major [d][n]
//   k_sort_columns    one CTA per column: stable LSD radix sort, 4 passes x 8 bits, ping-pong in HBM.
//                     All d columns sort concurrently (d CTAs); the table is read 8x in total.
//   k_edges_from_sorted  one CTA per column: distinct count, then midpoints or integer-rank quantiles.
#pragma once
#include "tst_common.cuh"

__device__ __forceinline__ uint32_t f32_key(float x) {
    uint32_t u = __float_as_uint(x);
    if (x != x) return 0xFFFFFFFFu;          // NaN sorts last (ignored by the edge rule)
    if ((u << 1) == 0u) u = 0u;              // -0.0 == +0.0
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_f32(uint32_t k) {
    uint32_t u = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k;
    return __uint_as_float(u);
}

__global__ void __launch_bounds__(256) k_keys_transpose(const float* __restrict__ X, int64_t n, int d, uint32_t* __restrict__ keys,
                                                        int64_t ldk) {
    __shared__ uint32_t s_t[32][33];
    const int64_t r0 = (int64_t)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    for (int i = ty; i < 32; i += 8) {
        int64_t r = r0 + i; int c = c0 + tx;
        s_t[i][tx] = (r < n && c < d) ? f32_key(X[r * d + c]) : 0xFFFFFFFFu;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        int c = c0 + i; int64_t r = r0 + tx;
        if (c < d && r < n) keys[(int64_t)c * ldk + r] = s_t[tx][i];
    }
}

// Stable LSD radix sort of one column per CTA.  Per 8-bit pass: digit histogram -> exclusive scan ->
// ordered scatter; inside a 1024-key tile the stable rank of a key is (keys with the same digit in
// earlier warps) + (earlier lanes of its warp with the same digit, via __match_any_sync).
#define SORT_THREADS 1024
__global__ void __launch_bounds__(SORT_THREADS) k_sort_columns(uint32_t* __restrict__ a, uint32_t* __restrict__ b, int64_t n, int64_t ldk) {
    constexpr int NW = SORT_THREADS / 32;
    __shared__ uint32_t s_base[256];
    __shared__ uint16_t s_wc[256][NW + 2];  // per-digit counts of each warp in the current tile
    const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    uint32_t* src = a + (int64_t)blockIdx.x * ldk;
    uint32_t* dst = b + (int64_t)blockIdx.x * ldk;
    for (int shift = 0; shift < 32; shift += 8) {
        if (tid < 256) s_base[tid] = 0;
        __syncthreads();
        for (int64_t i = tid; i < n; i += SORT_THREADS) atomicAdd(&s_base[(src[i] >> shift) & 0xFFu], 1u);
        __syncthreads();
        if (wid == 0) {  // exclusive scan of 256 counters by one warp (8 per lane)
            uint32_t v[8], run = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) { v[q] = s_base[lane * 8 + q]; run += v[q]; }
            uint32_t incl = run;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= (uint32_t)off) incl += t; }
            uint32_t ex = incl - run;
#pragma unroll
            for (int q = 0; q < 8; ++q) { s_base[lane * 8 + q] = ex; ex += v[q]; }
        }
        __syncthreads();
        uint32_t key_next = tid < n ? src[tid] : 0u;  // the next tile's key is requested one tile ahead (the CTA is alone on its SM: nothing else hides the latency)
        for (int64_t t0 = 0; t0 < n; t0 += SORT_THREADS) {
            const int64_t i = t0 + tid;
            const bool ok = i < n;
            const uint32_t key = key_next;
            key_next = i + SORT_THREADS < n ? src[i + SORT_THREADS] : 0u;
            const uint32_t dg = (key >> shift) & 0xFFu;
            for (int q = tid; q < 256 * (NW + 2); q += SORT_THREADS) ((uint16_t*)s_wc)[q] = 0;
            __syncthreads();
            // lanes of this warp holding the same digit (invalid lanes match among themselves on a 9th bit)
            const uint32_t peers = __match_any_sync(0xffffffffu, ok ? dg : 0x100u);
            const uint32_t rank_in_warp = __popc(peers & lanemask_lt());
            if (ok && rank_in_warp == 0) s_wc[dg][wid] = (uint16_t)__popc(peers);
            __syncthreads();
            // prefix over warps per digit (thread = digit), total appended at [NW]
            if (tid < 256) {
                uint32_t acc = 0;
                for (int w = 0; w < NW; ++w) { uint32_t c = s_wc[tid][w]; s_wc[tid][w] = (uint16_t)acc; acc += c; }
                s_wc[tid][NW] = (uint16_t)acc;
            }
            __syncthreads();
            if (ok) dst[s_base[dg] + s_wc[dg][wid] + rank_in_warp] = key;
            __syncthreads();
            if (tid < 256) s_base[tid] += s_wc[tid][NW];
            __syncthreads();
        }
        uint32_t* t = src; src = dst; dst = t;
    }
    // 4 passes: the sorted column is back in `a`
}

// One CTA per column over the sorted keys.
__global__ void __launch_bounds__(256) k_edges_from_sorted(const uint32_t* __restrict__ keys, int64_t n, int64_t ldk, int n_bins,
                                                           double* __restrict__ edges, int* __restrict__ n_edges,
                                                           uint32_t* __restrict__ scratch /* [d][256] distinct keys */) {
    __shared__ unsigned long long s_cnt;
    __shared__ uint32_t s_nd;
    const int col = blockIdx.x;
    const uint32_t* v = keys + (int64_t)col * ldk;
    double* e = edges + (size_t)col * TST_MAX_EDGES;
    uint32_t* dk = scratch + (size_t)col * 256;
    const int B = n_bins - 1;
    const uint32_t tid = threadIdx.x;
    if (tid == 0) { s_cnt = 0; s_nd = 0; }
    __syncthreads();
    // number of non-missing values (NaN keys are 0xFFFFFFFF and sort last)
    unsigned long long c = 0;
    for (int64_t i = tid; i < n; i += 256) c += v[i] != 0xFFFFFFFFu;
    atomicAdd(&s_cnt, c);
    __syncthreads();
    const int64_t nv = (int64_t)s_cnt;
    // distinct values: boundaries in the sorted order; the first B+1 are kept in order of position
    for (int64_t i0 = 0; i0 < nv && s_nd <= (uint32_t)B; i0 += 256) {  // s_nd only grows; reading it racily just delays the exit
        const int64_t i = i0 + tid;
        const bool isb = i < nv && (i == 0 || v[i] != v[i - 1]);
        // ordered append: ballot per warp, warp totals through shared memory
        __shared__ uint32_t s_w[8];
        const uint32_t bal = __ballot_sync(0xffffffffu, isb);
        if ((tid & 31u) == 0) s_w[tid >> 5] = __popc(bal);
        __syncthreads();
        uint32_t pre = s_nd, tot = 0;
        for (uint32_t w = 0; w < 8; ++w) { if (w < (tid >> 5)) pre += s_w[w]; tot += s_w[w]; }
        const uint32_t at = pre + __popc(bal & lanemask_lt());
        if (isb && at < 256) dk[at] = v[i];
        __syncthreads();
        if (tid == 0) s_nd += tot;
        __syncthreads();
    }
    const uint32_t nd = s_nd;
    if (nv == 0 || nd <= 1) { if (tid == 0) n_edges[col] = 0; return; }
    if (nd <= (uint32_t)B) {
        for (uint32_t i = tid; i + 1 < nd; i += 256) e[i] = fmin(((double)key_f32(dk[i]) + (double)key_f32(dk[i + 1])) * 0.5, 1e300);
        if (tid == 0) n_edges[col] = (int)nd - 1;
        return;
    }
    // quantiles: t = i * nv, j = t / B; v[j] if t % B else (v[j-1] + v[j]) / 2; then drop duplicates in order
    __shared__ double s_e[256];
    __shared__ uint32_t s_keep[8];
    double mine = 0.0;
    const int i = (int)tid + 1;
    if (i < B) {
        const long long t = (long long)i * nv, j = t / B, rem = t % B;
        mine = rem ? (double)key_f32(v[j]) : ((double)key_f32(v[j - 1]) + (double)key_f32(v[j])) * 0.5;
        s_e[tid] = mine;
    }
    __syncthreads();
    const bool keep = i < B && (tid == 0 || s_e[tid - 1] != mine);
    const uint32_t bal = __ballot_sync(0xffffffffu, keep);
    if ((tid & 31u) == 0) s_keep[tid >> 5] = __popc(bal);
    __syncthreads();
    uint32_t pre = 0, tot = 0;
    for (uint32_t w = 0; w < 8; ++w) { if (w < (tid >> 5)) pre += s_keep[w]; tot += s_keep[w]; }
    if (keep) e[pre + __popc(bal & lanemask_lt())] = fmin(mine, 1e300);
    if (tid == 0) n_edges[col] = (int)tot;
}
