// tst_prep.cuh -- table ingest, fold/partition bookkeeping, Philox bootstrap, entry compaction,
// OOF inference, vote reduction, s(x), IBP selection.  (north_star items 1, 2, 5, 6, 7.)
#pragma once
#include "tst_common.cuh"

// ---------------------------------------------------------------------------------------------
// K1 bin map.  Tile = 128 rows x d features.  X is read coalesced (feature fastest), the row-major
// table is written coalesced, the column-major table is written through a shared-memory transpose
// as 128-byte runs.  bin = #edges strictly below x (SK/.../_binning.pyx:76-85); NaN -> 255.
// Algorithmic bytes: n*d*(4 + 1 + 1).
#define BIN_TILE 128
__global__ void __launch_bounds__(256) k_bin_map(const float* __restrict__ X, int64_t n, int d,
                                                 const double* __restrict__ edges, const int* __restrict__ n_edges,
                                                 uint8_t* __restrict__ bins_cm, int64_t ld, uint8_t* __restrict__ bins_rm,
                                                 uint32_t dstride) {
    extern __shared__ uint8_t s_tile[];  // [d][BIN_TILE + 4]
    const int pitch = BIN_TILE + 4;
    const int64_t row0 = (int64_t)blockIdx.x * BIN_TILE;
    const int rows = (int)min((int64_t)BIN_TILE, n - row0);
    for (int idx = threadIdx.x; idx < rows * d; idx += blockDim.x) {
        int r = idx / d, f = idx - r * d;
        float xv = X[(row0 + r) * d + f];
        uint32_t b;
        if (xv != xv) {
            b = 255;
        } else {
            double x = (double)xv;
            const double* e = edges + (size_t)f * TST_MAX_EDGES;
            int lo = 0, hi = n_edges[f];
            while (lo < hi) {
                int mid = (lo + hi) >> 1;
                if (__ldg(e + mid) < x) lo = mid + 1; else hi = mid;
            }
            b = (uint32_t)lo;
        }
        s_tile[f * pitch + r] = (uint8_t)b;
        bins_rm[(row0 + r) * dstride + f] = (uint8_t)b;
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < rows * d; idx += blockDim.x) {
        int f = idx / rows, r = idx - f * rows;
        bins_cm[(int64_t)f * ld + row0 + r] = s_tile[f * pitch + r];
    }
}

// column-major -> row-major (tst_set_bins)
__global__ void __launch_bounds__(256) k_cm_to_rm(const uint8_t* __restrict__ cm, int64_t n, int d, int64_t ld_in,
                                                  uint8_t* __restrict__ cm_out, int64_t ld, uint8_t* __restrict__ rm,
                                                  uint32_t dstride) {
    extern __shared__ uint8_t s_tile[];
    const int pitch = BIN_TILE + 4;
    const int64_t row0 = (int64_t)blockIdx.x * BIN_TILE;
    const int rows = (int)min((int64_t)BIN_TILE, n - row0);
    for (int idx = threadIdx.x; idx < rows * d; idx += blockDim.x) {
        int f = idx / rows, r = idx - f * rows;
        uint8_t b = cm[(int64_t)f * ld_in + row0 + r];
        s_tile[f * pitch + r] = b;
        cm_out[(int64_t)f * ld + row0 + r] = b;
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < rows * d; idx += blockDim.x) {
        int r = idx / d, f = idx - r * d;
        rm[(row0 + r) * dstride + f] = s_tile[f * pitch + r];
    }
}

__global__ void k_copy_cm(const uint8_t* __restrict__ cm, int64_t n, int d, int64_t ld, uint8_t* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n * d) { int64_t f = i / n, r = i - f * n; out[i] = cm[f * ld + r]; }
}

// ---------------------------------------------------------------------------------------------
// Stratified folds.  One CTA per group walks the rows in order; per chunk one ballot
// per class gives the in-warp rank, a per-warp count table gives the in-chunk rank.
template <int THREADS>
__global__ void __launch_bounds__(THREADS) k_assign_folds(const uint8_t* __restrict__ y, const uint8_t* __restrict__ active,
                                                          const int32_t* __restrict__ group, int64_t n, int k, int C,
                                                          uint8_t* __restrict__ fold) {
    constexpr int NW = THREADS / 32;
    __shared__ uint32_t s_cnt[TST_MAX_CLASSES];   // class totals, then class base
    __shared__ uint32_t s_run[TST_MAX_CLASSES];
    __shared__ uint32_t s_w[TST_MAX_CLASSES][NW];
    const int g = blockIdx.x;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    for (int c = tid; c < C; c += THREADS) { s_cnt[c] = 0; s_run[c] = 0; }
    __syncthreads();
    for (int64_t i = tid; i < n; i += THREADS)
        if (active[i] && (group ? group[i] : 0) == g) atomicAdd(&s_cnt[y[i]], 1u);
    __syncthreads();
    if (tid == 0) {
        uint32_t acc = 0;
        for (int c = 0; c < C; ++c) { uint32_t v = s_cnt[c]; s_cnt[c] = acc; acc += v; }
    }
    __syncthreads();
    for (int64_t c0 = 0; c0 < n; c0 += THREADS) {
        int64_t i = c0 + tid;
        bool mine = i < n && (group ? group[i] : 0) == g;
        int cl = (mine && active[i]) ? (int)y[i] : -1;
        uint32_t myrank = 0;
        for (int c = 0; c < C; ++c) {
            uint32_t bal = __ballot_sync(0xffffffffu, cl == c);
            if (lane == 0) s_w[c][wid] = __popc(bal);
            if (cl == c) myrank = __popc(bal & lanemask_lt());
        }
        __syncthreads();
        if (cl >= 0) {
            uint32_t p = s_cnt[cl] + s_run[cl] + myrank;
            for (uint32_t w = 0; w < wid; ++w) p += s_w[cl][w];
            fold[i] = (uint8_t)(p % (uint32_t)k);
        } else if (mine) {
            fold[i] = TST_FOLD_INACTIVE;
        }
        __syncthreads();
        if ((int)tid < C) {
            uint32_t tot = 0;
            for (int w = 0; w < NW; ++w) tot += s_w[tid][w];
            s_run[tid] += tot;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// Partitions: rows sorted by (group, fold slot, row id); slot = fold (0..k-1) or k for TRAIN_ONLY.
__device__ __forceinline__ int part_key(const uint8_t* fold, const int32_t* group, int64_t i, int k) {
    uint32_t f = fold[i];
    if (f == TST_FOLD_INACTIVE) return -1;
    int slot = (f == TST_FOLD_TRAIN_ONLY) ? k : (int)f;
    if (slot > k) return -1;
    return (group ? group[i] : 0) * (k + 1) + slot;
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS) k_part_count(const uint8_t* __restrict__ fold, const int32_t* __restrict__ group,
                                                        int64_t n, int k, uint32_t* __restrict__ part_count) {
    __shared__ uint32_t s_c;
    const int p = blockIdx.x;
    if (threadIdx.x == 0) s_c = 0;
    __syncthreads();
    uint32_t c = 0;
    for (int64_t i = threadIdx.x; i < n; i += THREADS) c += part_key(fold, group, i, k) == p;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
    if ((threadIdx.x & 31) == 0) atomicAdd(&s_c, c);
    __syncthreads();
    if (threadIdx.x == 0) part_count[p] = s_c;
}

__global__ void k_part_scan(const uint32_t* __restrict__ part_count, int np, uint32_t* __restrict__ part_start) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        uint32_t acc = 0;
        for (int p = 0; p < np; ++p) { part_start[p] = acc; acc += part_count[p]; }
        part_start[np] = acc;
    }
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS) k_part_scatter(const uint8_t* __restrict__ fold, const int32_t* __restrict__ group,
                                                          int64_t n, int k, const uint32_t* __restrict__ part_start,
                                                          uint32_t* __restrict__ sorted) {
    constexpr int NW = THREADS / 32;
    __shared__ uint32_t s_w[2][NW];
    const int p = blockIdx.x;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    uint32_t base = part_start[p];
    int par = 0;
    for (int64_t c0 = 0; c0 < n; c0 += THREADS, par ^= 1) {
        int64_t i = c0 + tid;
        bool hit = i < n && part_key(fold, group, i, k) == p;
        uint32_t bal = __ballot_sync(0xffffffffu, hit);
        if (lane == 0) s_w[par][wid] = __popc(bal);
        __syncthreads();
        uint32_t pre = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) { uint32_t a = s_w[par][w]; if (w < (int)wid) pre += a; tot += a; }
        if (hit) sorted[base + pre + __popc(bal & lanemask_lt())] = (uint32_t)i;
        base += tot;
    }
}

// ---------------------------------------------------------------------------------------------
// K2 bootstrap: draw j of tree t is word (j & 3) of Philox(ctr = (j >> 2, 0, gid, BOOT)), mapped to
// a training row by mulhi (no modulo, no rejection).  Multiplicities accumulate as bytes packed in
// 32-bit words (SK/ensemble/_forest.py:95-113,155-156: indices -> bincount).
__global__ void __launch_bounds__(256) k_boot_draw(const TreeDesc* __restrict__ trees, const uint32_t* __restrict__ sorted,
                                                   uint32_t* __restrict__ counts8, uint64_t n_pad, uint32_t k0, uint32_t k1) {
    const TreeDesc td = trees[blockIdx.y];
    uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;  // Philox block index
    if ((uint64_t)q * 4u >= td.n_train) return;
    uint4 r = philox4x32_10(q, 0u, td.gid, TST_PURPOSE_BOOT, k0, k1);
    uint32_t* cnt = counts8 + (uint64_t)blockIdx.y * (n_pad / 4);
#pragma unroll
    for (uint32_t i = 0; i < 4; ++i) {
        uint32_t j = q * 4u + i;
        if (j < td.n_train) {
            uint32_t u = mulhi32(pick4(r, i), td.n_train);
            uint32_t row = sorted[u < td.a ? td.gs + u : td.gs + u + td.fsz];
            atomicAdd(&cnt[row >> 2], 1u << (8u * (row & 3u)));
        }
    }
}

// Entry compaction: in-bag rows ascending, entry = (row, multiplicity | class << 8) (zero-weight rows are
// dropped as SK/tree/_splitter.pyx:167-171).  Chunk = 4096 rows: count -> per-tree scan -> scatter.
#define BOOT_CHUNK 4096
__global__ void __launch_bounds__(256) k_boot_count(const uint32_t* __restrict__ counts8, uint64_t n_pad, uint32_t n_chunks,
                                                    uint32_t* __restrict__ chunk_cnt, uint32_t* __restrict__ chunk_sum) {
    __shared__ uint32_t s_c, s_s;
    if (threadIdx.x == 0) { s_c = 0; s_s = 0; }
    __syncthreads();
    const uint4* src = (const uint4*)(counts8 + (uint64_t)blockIdx.y * (n_pad / 4));
    uint64_t v = (uint64_t)blockIdx.x * 256 + threadIdx.x;  // uint4 index
    uint32_t c = 0, s = 0;
    if (v * 16 < n_pad) {
        uint4 q = src[v];
        uint32_t ws[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int b = 0; b < 4; ++b) { uint32_t m = (ws[i] >> (8 * b)) & 0xFFu; c += m != 0; s += m; }
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) { c += __shfl_xor_sync(0xffffffffu, c, off); s += __shfl_xor_sync(0xffffffffu, s, off); }
    if ((threadIdx.x & 31) == 0) { atomicAdd(&s_c, c); atomicAdd(&s_s, s); }
    __syncthreads();
    if (threadIdx.x == 0) {
        chunk_cnt[(uint64_t)blockIdx.y * n_chunks + blockIdx.x] = s_c;
        chunk_sum[(uint64_t)blockIdx.y * n_chunks + blockIdx.x] = s_s;
    }
}

__global__ void __launch_bounds__(256) k_boot_scan(TreeDesc* __restrict__ trees, uint32_t n_chunks, uint32_t* __restrict__ chunk_cnt,
                                                   const uint32_t* __restrict__ chunk_sum, DevStats* __restrict__ stats) {
    // one CTA per tree: exclusive scan of the chunk counts (in place), weight check
    __shared__ uint32_t s_w[8], s_ws[8];
    __shared__ uint32_t s_base, s_sum;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    uint32_t* cc = chunk_cnt + (uint64_t)blockIdx.x * n_chunks;
    const uint32_t* cs = chunk_sum + (uint64_t)blockIdx.x * n_chunks;
    if (tid == 0) { s_base = 0; s_sum = 0; }
    __syncthreads();
    for (uint32_t c0 = 0; c0 < n_chunks; c0 += 256) {
        uint32_t i = c0 + tid;
        uint32_t v = i < n_chunks ? cc[i] : 0u, sv = i < n_chunks ? cs[i] : 0u;
        uint32_t incl = v, sincl = sv;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            uint32_t a = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= (uint32_t)off) incl += a;
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) sincl += __shfl_xor_sync(0xffffffffu, sincl, off);
        if (lane == 31) s_w[wid] = incl;
        if (lane == 0) s_ws[wid] = sincl;
        __syncthreads();
        uint32_t pre = s_base, tot = 0, stot = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) { if (w < (int)wid) pre += s_w[w]; tot += s_w[w]; stot += s_ws[w]; }
        if (i < n_chunks) cc[i] = pre + incl - v;
        __syncthreads();
        if (tid == 0) { s_base += tot; s_sum += stot; }
        __syncthreads();
    }
    if (tid == 0) {
        trees[blockIdx.x].n_inbag = s_base;
        atomicAdd(&stats->inbag_rows, (unsigned long long)s_base);
        if (s_sum != trees[blockIdx.x].n_train) stats->error = 1;  // a byte counter wrapped (multiplicity > 255)
    }
}

__global__ void __launch_bounds__(256) k_boot_scatter(const uint32_t* __restrict__ counts8, uint64_t n_pad, uint32_t n_chunks,
                                                      const uint32_t* __restrict__ chunk_off, const uint8_t* __restrict__ y,
                                                      const uint16_t* __restrict__ yq /* criterion 2: the 16-bit target rides where the class does */,
                                                      Entry* __restrict__ entA, uint64_t n_cap) {
    __shared__ uint32_t s_w[8];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    const uint4* src = (const uint4*)(counts8 + (uint64_t)blockIdx.y * (n_pad / 4));
    uint64_t v = (uint64_t)blockIdx.x * 256 + tid;
    uint32_t ws[4] = {0, 0, 0, 0};
    if (v * 16 < n_pad) { uint4 q = src[v]; ws[0] = q.x; ws[1] = q.y; ws[2] = q.z; ws[3] = q.w; }
    uint32_t c = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int b = 0; b < 4; ++b) c += ((ws[i] >> (8 * b)) & 0xFFu) != 0;
    uint32_t incl = c;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        uint32_t a = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= (uint32_t)off) incl += a;
    }
    if (lane == 31) s_w[wid] = incl;
    __syncthreads();
    uint32_t pos = chunk_off[(uint64_t)blockIdx.y * n_chunks + blockIdx.x] + incl - c;
    for (uint32_t w = 0; w < wid; ++w) pos += s_w[w];
    Entry* out = entA + (uint64_t)blockIdx.y * n_cap;
    uint32_t row = (uint32_t)(v * 16);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            uint32_t m = (ws[i] >> (8 * b)) & 0xFFu;
            if (m) { uint32_t rw = row + i * 4 + b; out[pos++] = make_uint2(rw, m | ((yq ? (uint32_t)yq[rw] : (uint32_t)y[rw]) << 8)); }
        }
}

// Root tasks: empty tree -> empty leaf; otherwise routed by size like any other node.
__global__ void k_init_roots(GrowCtx g, uint32_t n_trees) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_trees) return;
    const TreeDesc td = g.trees[t];
    if (td.n_inbag == 0) {
        g.nodes[(uint64_t)td.arena * g.node_stride] = make_uint2(0u, (KIND_PURE << 30));
        atomicAdd(&g.stats->nodes, 1ull);
        atomicAdd(&g.stats->leaves, 1ull);
        return;
    }
    Task k;
    k.tree = t; k.begin = 0; k.end = td.n_inbag; k.depth = 0; k.slot = 0; k.pad0 = k.pad1 = k.pad2 = 0;
    if (g.rsplit && td.n_inbag > g.big_min) k.pad0 = 1;  // pre-split by k_root_hist / k_root_split
    if (td.n_inbag <= g.tiny_max) tiny_push(g, t, 0u, td.n_inbag, 0u, 0u, 0u);
    else route_push(g, td.n_inbag, td.n_train, k);
}

// ---------------------------------------------------------------------------------------------
// K8 OOF inference (SK/tree/_tree.pyx:954-996 on bins).  Thread = one held-out row of one partition,
// looping over that partition's trees in the batch; each tree adds the Q0.16 vote
// (count_c << 16) / N of the reached leaf (SK/tree/_criterion.pyx:483-486, _forest.py:961-962).
struct InferPart { uint32_t fs, fsz, tree_lo, tree_hi; };

template <bool ATOMIC>
__device__ __forceinline__ void add_leaf_votes(uint2 rec, const uint32_t* __restrict__ leafc, int C, uint32_t* __restrict__ vrow, bool reg) {
    uint32_t kind = rec.y >> 30;
    if (reg) {
        // criterion 2: the leaf mean in Q16.16, floor((S << 16) / N), added as two 16-bit halves -- vrow[0] collects the low
        // halves, vrow[1] the high ones, so no carry ever crosses a word and rank / batch sums stay element-wise
        unsigned long long N, S;
        if (kind == KIND_PURE) { N = rec.x; S = N * (rec.y & 0xFFFFu); }
        else { N = rec.y & 0x3FFFFFFFu; const uint32_t* lc = leafc + (uint64_t)rec.x * 2; S = (unsigned long long)lc[0] | ((unsigned long long)lc[1] << 32); }
        if (N) {
            const uint32_t v = (uint32_t)((S << 16) / N);
            if (ATOMIC) { atomicAdd(&vrow[0], v & 0xFFFFu); atomicAdd(&vrow[1], v >> 16); } else { vrow[0] += v & 0xFFFFu; vrow[1] += v >> 16; }
        }
        return;
    }
    if (kind == KIND_PURE) {
        if (rec.x) { if (ATOMIC) atomicAdd(&vrow[rec.y & 0xFFu], 65536u); else vrow[rec.y & 0xFFu] += 65536u; }
    } else {
        const uint32_t* lc = leafc + (uint64_t)rec.x * C;
        unsigned long long N = 0;
        for (int c = 0; c < C; ++c) N += lc[c];
        if (N) for (int c = 0; c < C; ++c) {
            uint32_t v = (uint32_t)(((unsigned long long)lc[c] << 16) / N);
            if (v) { if (ATOMIC) atomicAdd(&vrow[c], v); else vrow[c] += v; }
        }
    }
}

// grid = (row blocks, partitions, tree chunks).  With one tree chunk a thread owns its row's votes; benchmark-sized
// tables have too few held-out rows to fill the GPU, so the trees are cut into chunks as well and the (integer,
// order-independent) votes are added atomically.
template <bool ATOMIC>
__global__ void __launch_bounds__(256) k_oof_infer(GrowCtx g, const InferPart* __restrict__ parts, const uint32_t* __restrict__ sorted,
                                                   uint32_t* __restrict__ votes, uint32_t trees_per_chunk) {
    const InferPart ip = parts[blockIdx.y];
    uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t tlo = ip.tree_lo + blockIdx.z * trees_per_chunk, thi = min(ip.tree_hi, tlo + trees_per_chunk);
    unsigned long long steps = 0;
    if (q < ip.fsz && tlo < thi) {
        const uint32_t row = sorted[ip.fs + q];
        const uint8_t* __restrict__ rp = g.bins_rm + (uint64_t)row * g.dstride;
        uint32_t* vrow = votes + (uint64_t)row * g.C;
        for (uint32_t t = tlo; t < thi; ++t) {
            const uint32_t arena = g.trees[t].arena;
            const uint2* __restrict__ nodes = g.nodes + (uint64_t)arena * g.node_stride;
            uint32_t slot = 0;
            uint2 rec = nodes[0];
            while ((rec.y >> 30) == KIND_INTERNAL) {
                uint32_t f = (rec.y >> 8) & 0xFFFFu, b = rec.y & 0xFFu;
                slot = rp[f] <= b ? slot + 1u : rec.x;
                rec = nodes[slot];
                steps++;
            }
            add_leaf_votes<ATOMIC>(rec, g.leafc + (uint64_t)arena * g.leafc_stride, g.C, vrow, g.criterion == TST_REG);
        }
    }
    // warp-aggregate the step counter
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) steps += __shfl_xor_sync(0xffffffffu, steps, off);
    if ((threadIdx.x & 31) == 0 && steps) atomicAdd(&g.stats->oof_steps, steps);
}

// Kept-forest predict on freshly binned rows (row-major [n][dstride]).
__global__ void __launch_bounds__(256) k_predict(const uint8_t* __restrict__ rm, uint32_t dstride, int64_t n, int C,
                                                 const uint2* __restrict__ nodes_all, uint64_t node_stride,
                                                 const uint32_t* __restrict__ leafc_all, uint64_t leafc_stride,
                                                 const uint32_t* __restrict__ arenas, uint32_t n_trees, uint32_t* __restrict__ votes, bool reg) {
    int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    const uint8_t* __restrict__ rp = rm + row * dstride;
    uint32_t* vrow = votes + row * C;
    for (uint32_t t = 0; t < n_trees; ++t) {
        const uint32_t arena = arenas[t];
        const uint2* __restrict__ nodes = nodes_all + (uint64_t)arena * node_stride;
        uint32_t slot = 0;
        uint2 rec = nodes[0];
        while ((rec.y >> 30) == KIND_INTERNAL) {
            uint32_t f = (rec.y >> 8) & 0xFFFFu, b = rec.y & 0xFFu;
            slot = rp[f] <= b ? slot + 1u : rec.x;
            rec = nodes[slot];
        }
        add_leaf_votes<false>(rec, leafc_all + (uint64_t)arena * leafc_stride, C, vrow, reg);
    }
}

// Feature importances: per-tree normalisation and the forest-level integer sums (one thread per tree).
// sums[f] += floor((double)imp_t[f] / (double)sum_f imp_t[f] * 2^32); sums[d] counts the trees with at least one split.
__global__ void k_imp_reduce(const unsigned long long* __restrict__ imp, int64_t n_trees, int d, unsigned long long* __restrict__ sums) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_trees) return;
    const unsigned long long* r = imp + t * d;
    unsigned long long tsum = 0;
    for (int f = 0; f < d; ++f) tsum += r[f];
    if (!tsum) return;  // a one-leaf tree has no importances (SK/ensemble/_forest.py: node_count > 1)
    const double ds = __ull2double_rn(tsum);
    for (int f = 0; f < d; ++f) {
        const unsigned long long v = r[f];
        if (v) atomicAdd(&sums[f], __double2ull_rd(__dmul_rn(__ddiv_rn(__ull2double_rn(v), ds), 4294967296.0)));
    }
    atomicAdd(&sums[d], 1ull);
}

// K9 rank-ordered vote reduction (integers: the order cannot matter, it is fixed anyway).
__global__ void k_reduce_votes(const uint32_t* __restrict__ gathered, int world, int64_t nv, uint32_t* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nv) return;
    uint32_t acc = 0;
    for (int r = 0; r < world; ++r) acc += gathered[(int64_t)r * nv + i];
    out[i] = acc;
}

// s(x) = votes[i][y_i] / (T * 65536) in double -> float; per-partition accuracy counts.
__global__ void k_finalize(const uint32_t* __restrict__ votes, const uint8_t* __restrict__ y, const uint8_t* __restrict__ fold,
                           const int32_t* __restrict__ group, int64_t n, int C, int k, double denom, float* __restrict__ s,
                           uint32_t* __restrict__ correct, uint32_t* __restrict__ total) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (fold[i] >= k) { if (s) s[i] = -1.0f; return; }
    const uint32_t* v = votes + i * C;
    if (s) s[i] = (float)__ddiv_rn((double)v[y[i]], denom);
    int best = 0;
    for (int c = 1; c < C; ++c) if (v[c] > v[best]) best = c;
    int part = (group ? group[i] : 0) * k + fold[i];
    if (total) atomicAdd(&total[part], 1u);
    if (correct && best == (int)y[i]) atomicAdd(&correct[part], 1u);
}

// ---------------------------------------------------------------------------------------------
// K10 IBP selection, one CTA: radix select (4 x 8 bits, MSB first) of the batch-th largest key among
// candidates, then an ordered pass that clears rows above the threshold and the first `rem` ties.
template <int THREADS>
__global__ void __launch_bounds__(THREADS) k_ibp_select(const uint32_t* __restrict__ votes, const uint8_t* __restrict__ y,
                                                        const uint8_t* __restrict__ fold, int64_t n, int C, int k,
                                                        unsigned long long vmin, long long batch, uint8_t* __restrict__ active,
                                                        long long* __restrict__ removed) {
    constexpr int NW = THREADS / 32;
    __shared__ uint32_t s_h[256];
    __shared__ uint32_t s_w[2][NW];
    __shared__ unsigned long long s_m;
    __shared__ uint32_t s_prefix, s_need;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    auto key_of = [&](int64_t i, bool& cand) -> uint32_t {
        cand = false;
        if (i >= n || !active[i] || fold[i] >= k) return 0u;
        uint32_t v = votes[i * C + y[i]];
        cand = (unsigned long long)v >= vmin;
        return v;
    };
    if (tid == 0) s_m = 0;
    __syncthreads();
    unsigned long long m = 0;
    for (int64_t i = tid; i < n; i += THREADS) { bool c; key_of(i, c); m += c; }
    atomicAdd(&s_m, m);
    __syncthreads();
    const unsigned long long n_cand = s_m;
    if (n_cand == 0 || batch <= 0) { if (tid == 0) *removed = 0; return; }
    if (n_cand <= (unsigned long long)batch) {
        for (int64_t i = tid; i < n; i += THREADS) { bool c; key_of(i, c); if (c) active[i] = 0; }
        if (tid == 0) *removed = (long long)n_cand;
        return;
    }
    // find v* = batch-th largest key: walk digits from the top
    if (tid == 0) { s_prefix = 0; s_need = (uint32_t)batch; }
    for (int shift = 24; shift >= 0; shift -= 8) {
        for (int b = tid; b < 256; b += THREADS) s_h[b] = 0;
        __syncthreads();
        const uint32_t prefix = s_prefix;
        const uint32_t himask = shift == 24 ? 0u : (0xFFFFFFFFu << (shift + 8));
        for (int64_t i = tid; i < n; i += THREADS) {
            bool c; uint32_t v = key_of(i, c);
            if (c && (v & himask) == prefix) atomicAdd(&s_h[(v >> shift) & 0xFFu], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            uint32_t need = s_need;
            int dgt = 255;
            for (; dgt > 0; --dgt) { if (s_h[dgt] >= need) break; need -= s_h[dgt]; }
            s_need = need;  // still to take inside this digit's bucket
            s_prefix = prefix | ((uint32_t)dgt << shift);
        }
        __syncthreads();
    }
    const uint32_t vstar = s_prefix;
    const uint32_t rem = s_need;  // ties at v* to take, lowest row ids first
    uint32_t taken = 0;
    int par = 0;
    for (int64_t c0 = 0; c0 < n; c0 += THREADS, par ^= 1) {
        int64_t i = c0 + tid;
        bool c; uint32_t v = key_of(i, c);
        bool tie = c && v == vstar;
        uint32_t bal = __ballot_sync(0xffffffffu, tie);
        if (lane == 0) s_w[par][wid] = __popc(bal);
        __syncthreads();
        uint32_t pre = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) { uint32_t a = s_w[par][w]; if (w < (int)wid) pre += a; tot += a; }
        if (c && v > vstar) active[i] = 0;
        else if (tie && taken + pre + __popc(bal & lanemask_lt()) < rem) active[i] = 0;
        taken += tot;
    }
    if (tid == 0) *removed = batch;
}
