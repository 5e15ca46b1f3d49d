/*
 * oracle/hist_rf.c -- CPU restatement of the TsT histogram Random-Forest diagnostic.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing on the product path (test-set-training_b200/)
 * may import, link or execute this file; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs do, and there only as the checker
 * or the timed CPU baseline.
 *
 * PARITY STATUS: "parity unpinned".  The reference repository
 * (vision-x-nyu/test-set-training) ships no code, tests or golden vectors
 * (REF/README.md:38-39 "Coming soon!").  The behaviour restated here is the
 * paper's "lightweight Random Forest-based diagnostic" (REF/README.md:35) in the
 * histogram formulation fixed by DESIGN.md section 2 ("normative spec"); the
 * arithmetic follows scikit-learn 1.9.0 (third-party, not vendored in the
 * reference, version = the one in this image).  What pins this file instead:
 *   - Random123 Philox4x32-10 known-answer vectors            (tests/test_oracle_kat.py)
 *   - scikit-learn's own integer KATs for histograms and bin edges
 *     (SK/ensemble/_hist_gradient_boosting/tests/test_histogram.py:20-50,
 *      .../tests/test_binning.py:28-50,96-115)                 (tests/test_oracle_kat.py)
 *   - statistical agreement with sklearn's exact-split RandomForestClassifier
 *     (oracle/sklearn_tst.py; tests/test_oracle_vs_sklearn.py)
 *
 * SK/ below = site-packages/sklearn (1.9.0).
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_FOLD_TRAIN_ONLY 254
#define ORC_FOLD_INACTIVE 255
#define ORC_PURPOSE_BOOT 1u
#define ORC_PURPOSE_FEAT 2u
#define ORC_NBINS 256
#define ORC_MAX_EDGES 255

typedef struct {
    int32_t max_depth;         /* <= 0: unlimited  (SK/tree/_classes.py:315 max_depth=None) */
    int32_t min_samples_split; /* >= 2            (SK/tree/_classes.py:317-328) */
    int32_t min_samples_leaf;  /* >= 1 */
    int32_t max_features;      /* m in [1, d]; "sqrt" resolved by the caller (SK/tree/_classes.py:330-345) */
    int32_t criterion;         /* 0 = Gini (classification); 1 = squared error on ordinal levels: the labels 0..C-1 are
                                  the levels of a uniformly quantised numeric target (SK/tree/_criterion.pyx:707-1000);
                                  2 = squared error on a continuous target in 16-bit fixed point (orc_set_target): per-bin
                                  (count, sum y) histograms, leaves hold (N, sum y), votes are Q16.16 leaf means */
} orc_params;

typedef struct {
    uint64_t row_visits;   /* V: sum over searched nodes of distinct in-bag rows */
    uint64_t nodes;        /* all nodes of all trees built by this call */
    uint64_t leaves;
    uint64_t searched;     /* nodes on which a split search ran */
    uint64_t fallback;     /* searched nodes that needed features beyond the first m */
    uint64_t inbag_rows;   /* sum over trees of distinct in-bag rows */
    uint64_t trees;
    uint64_t oof_steps;    /* node hops during OOF inference */
    uint32_t max_depth;
    uint32_t pad;
} orc_stats;

typedef struct {
    int32_t left, right;   /* child node index, -1 for leaves */
    int32_t feature;       /* original feature index, -1 for leaves */
    int32_t bin;           /* row goes left iff bin[row][feature] <= bin */
    uint32_t begin, end;   /* segment of the tree's in-bag row list */
    uint32_t depth;
    uint32_t n_weighted;   /* sum of bootstrap multiplicities in the node */
} orc_node;

typedef struct {
    int64_t n_nodes;
    orc_node* nodes;   /* processing order; export re-sorts canonically */
    uint32_t* counts;  /* [n_nodes][C] */
    int32_t part;      /* partition = group*(k) + fold */
    int32_t tree;      /* tree index within the partition */
} orc_tree;

typedef struct {
    int32_t n_trees_total;
    int32_t C;
    orc_tree* trees; /* indexed by part*n_trees + t; n_nodes == 0 if not built by this rank */
} orc_forest;

/* ---------------------------------------------------------------- Philox4x32-10
 * Random123 / cuRAND constants (CUDA/curand_philox4x32_x.h:88-91,160-191). */
static inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                 uint32_t k1, uint32_t out[4]) {
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
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void orc_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1], out);
}

static inline uint32_t mulhi32(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }

/* ---------------------------------------------------------------- binning
 * bin = number of edges strictly below x, i.e. edges[b-1] < x <= edges[b]
 * (SK/ensemble/_hist_gradient_boosting/_binning.pyx:50-85); NaN -> 255.
 * edges: [d][255] float64, n_edges[d] valid entries per feature.
 * out: column-major [d][n]. */
int orc_bin(const float* X, int64_t n, int32_t d, const double* edges, const int32_t* n_edges,
            uint8_t* bins_cm) {
    for (int32_t f = 0; f < d; ++f) {
        const double* e = edges + (size_t)f * ORC_MAX_EDGES;
        int32_t ne = n_edges[f];
        if (ne < 0 || ne > ORC_MAX_EDGES) return -1;
        for (int64_t i = 0; i < n; ++i) {
            float xv = X[(size_t)i * d + f];
            uint8_t b;
            if (xv != xv) {
                b = 255;
            } else {
                double x = (double)xv;
                int32_t lo = 0, hi = ne; /* first index with e[idx] >= x */
                while (lo < hi) {
                    int32_t mid = (lo + hi) >> 1;
                    if (e[mid] < x) lo = mid + 1; else hi = mid;
                }
                b = (uint8_t)lo;
            }
            bins_cm[(size_t)f * n + i] = b;
        }
    }
    return 0;
}

/* ---------------------------------------------------------------- stratified folds
 * within each group, position p = (#active rows of lower class) +
 * (rank of the row among active rows of its class, ascending row id); fold = p mod k.
 * Same per-class allocation counts as sklearn's StratifiedKFold
 * (SK/model_selection/_split.py:774-840), dealt round-robin instead of in blocks. */
int orc_assign_folds(const uint8_t* y, const uint8_t* active, const int32_t* group, int32_t n_groups,
                     int64_t n, int32_t k, int32_t C, uint8_t* fold) {
    if (k < 1 || k > 250 || C < 1 || C > 256 || n_groups < 1) return -1;
    size_t m = (size_t)n_groups * C;
    uint32_t* cnt = (uint32_t*)calloc(m, 4);
    uint32_t* base = (uint32_t*)calloc(m, 4);
    if (!cnt || !base) return -2;
    for (int64_t i = 0; i < n; ++i) {
        if (!active[i]) continue;
        int32_t g = group ? group[i] : 0;
        if (g < 0 || g >= n_groups || y[i] >= C) { free(cnt); free(base); return -1; }
        cnt[(size_t)g * C + y[i]]++;
    }
    for (int32_t g = 0; g < n_groups; ++g) {
        uint32_t acc = 0;
        for (int32_t c = 0; c < C; ++c) { base[(size_t)g * C + c] = acc; acc += cnt[(size_t)g * C + c]; }
    }
    memset(cnt, 0, m * 4);
    for (int64_t i = 0; i < n; ++i) {
        if (!active[i]) { fold[i] = ORC_FOLD_INACTIVE; continue; }
        int32_t g = group ? group[i] : 0;
        size_t q = (size_t)g * C + y[i];
        uint32_t p = base[q] + cnt[q]++;
        fold[i] = (uint8_t)(p % (uint32_t)k);
    }
    free(cnt); free(base);
    return 0;
}

/* ---------------------------------------------------------------- partitions
 * Rows sorted by (group, fold slot, row id); fold slot = fold for 0..k-1, k for
 * TRAIN_ONLY; inactive rows are dropped.  part_start has n_groups*(k+1)+1 entries. */
typedef struct {
    uint32_t* sorted;
    uint32_t* part_start;
    int32_t n_groups, k;
} orc_parts;

static int build_parts(const uint8_t* fold, const int32_t* group, int32_t n_groups, int64_t n,
                       int32_t k, orc_parts* P) {
    int32_t np = n_groups * (k + 1);
    P->n_groups = n_groups; P->k = k;
    P->part_start = (uint32_t*)calloc((size_t)np + 1, 4);
    P->sorted = (uint32_t*)malloc((size_t)(n > 0 ? n : 1) * 4);
    if (!P->part_start || !P->sorted) return -2;
    for (int64_t i = 0; i < n; ++i) {
        uint8_t f = fold[i];
        if (f == ORC_FOLD_INACTIVE) continue;
        int32_t slot = (f == ORC_FOLD_TRAIN_ONLY) ? k : f;
        if (slot > k) return -1;
        int32_t g = group ? group[i] : 0;
        if (g < 0 || g >= n_groups) return -1;
        P->part_start[(size_t)g * (k + 1) + slot + 1]++;
    }
    for (int32_t p = 0; p < np; ++p) P->part_start[p + 1] += P->part_start[p];
    uint32_t* cur = (uint32_t*)malloc((size_t)np * 4);
    memcpy(cur, P->part_start, (size_t)np * 4);
    for (int64_t i = 0; i < n; ++i) {
        uint8_t f = fold[i];
        if (f == ORC_FOLD_INACTIVE) continue;
        int32_t slot = (f == ORC_FOLD_TRAIN_ONLY) ? k : f;
        int32_t g = group ? group[i] : 0;
        P->sorted[cur[(size_t)g * (k + 1) + slot]++] = (uint32_t)i;
    }
    free(cur);
    return 0;
}

/* Continuous target of criterion 2: yq[i] in [0, 65535] (host: round((y - lo) / (hi - lo) * 65535)).  Set before orc_fit_oof;
 * the uint8 labels passed to orc_fit_oof are then only the stratification levels of the fold assignment. */
static const uint16_t* g_target16 = NULL;
void orc_set_target(const uint16_t* yq) { g_target16 = yq; }

/* ---------------------------------------------------------------- one tree */
typedef struct { uint32_t row; uint32_t w; } entry_t;

typedef struct {
    uint32_t begin, end, depth;
    int32_t parent; int32_t is_right;
} stack_item;

typedef struct {
    /* growable node store */
    orc_node* nodes; uint32_t* counts; int64_t n_nodes, cap;
} node_store;

static int64_t store_push(node_store* S, int32_t C) {
    if (S->n_nodes == S->cap) {
        int64_t nc = S->cap ? S->cap * 2 : 1024;
        S->nodes = (orc_node*)realloc(S->nodes, (size_t)nc * sizeof(orc_node));
        S->counts = (uint32_t*)realloc(S->counts, (size_t)nc * C * 4);
        S->cap = nc;
    }
    return S->n_nodes++;
}

/* gain = SL/NL + SR/NR with exact uint64 numerators: sklearn's Gini proxy
 * (SK/tree/_criterion.pyx:147-163,661-687) minus a per-node constant; two IEEE
 * divisions and one add, so no FMA contraction is possible (compile with
 * -ffp-contract=off anyway). */
static inline double gain_of(uint64_t SL, uint64_t NL, uint64_t SR, uint64_t NR) {
    volatile double a = (double)SL / (double)NL;
    volatile double b = (double)SR / (double)NR;
    return a + b;
}

/* squared-error proxy (SK/tree/_criterion.pyx MSE.proxy_impurity_improvement): sum_left^2 / n_left +
 * sum_right^2 / n_right with S = sum over levels c of c * count_c (exact uint64); one multiply, one divide per
 * side, one add, each rounded once. */
static inline double gain_mse(uint64_t SL, uint64_t NL, uint64_t SR, uint64_t NR) {
    volatile double sl = (double)SL, sr = (double)SR;
    volatile double pl = sl * sl, pr = sr * sr;
    volatile double a = pl / (double)NL;
    volatile double b = pr / (double)NR;
    return a + b;
}

/* Impurity decrease of one split in Q24 fixed point (feature importances, SK/tree/_tree.pyx:1244-1280):
 * N*imp - N_L*imp_L - N_R*imp_R = gain - parent_term with the proxy the search maximised; floor(x * 2^24), negative -> 0. */
static inline uint64_t imp_q24(double gain, double parent_term) {
    volatile double c = gain - parent_term;
    if (!(c > 0.0)) return 0;
    volatile double scaled = c * 16777216.0;
    return (uint64_t)floor(scaled);
}

static int grow_tree(const uint8_t* bins_cm, int64_t n, int32_t d, const uint8_t* y, int32_t C,
                     const orc_params* prm, uint32_t k0, uint32_t k1, uint32_t gid, entry_t* ent,
                     entry_t* tmp, uint32_t n_inbag, node_store* S, orc_stats* st, uint64_t* imp /* [d] or NULL */) {
    int32_t m = prm->max_features;
    uint32_t msl = (uint32_t)prm->min_samples_leaf, mss = (uint32_t)prm->min_samples_split;
    uint32_t* hist = (uint32_t*)calloc((size_t)ORC_NBINS * C, 4);
    uint64_t* tot = (uint64_t*)malloc((size_t)C * 8);
    uint64_t* cum = (uint64_t*)malloc((size_t)C * 8);
    int32_t* perm = (int32_t*)malloc((size_t)d * 4);
    size_t scap = 64, sp = 0;
    stack_item* stk = (stack_item*)malloc(scap * sizeof(stack_item));
    stk[sp++] = (stack_item){0, n_inbag, 0, -1, 0};
    while (sp) {
        stack_item it = stk[--sp];
        int64_t id = store_push(S, C);
        orc_node* nd = &S->nodes[id];
        uint32_t* cnts = &S->counts[(size_t)id * C];
        if (it.parent >= 0) {
            if (it.is_right) S->nodes[it.parent].right = (int32_t)id; else S->nodes[it.parent].left = (int32_t)id;
        }
        /* node totals: SK/tree/_criterion.pyx:302-370 (weighted class counts) */
        for (int32_t c = 0; c < C; ++c) tot[c] = 0;
        uint64_t N = 0;
        for (uint32_t i = it.begin; i < it.end; ++i) { tot[y[ent[i].row]] += ent[i].w; N += ent[i].w; }
        int pure = 0;
        for (int32_t c = 0; c < C; ++c) { cnts[c] = (uint32_t)tot[c]; if (tot[c] == N) pure = 1; }
        nd->left = nd->right = -1; nd->feature = -1; nd->bin = -1;
        nd->begin = it.begin; nd->end = it.end; nd->depth = it.depth; nd->n_weighted = (uint32_t)N;
        st->nodes++;
        if (it.depth > st->max_depth) st->max_depth = it.depth;
        /* stop rules: SK/tree/_tree.pyx:230-240 */
        int leaf = (prm->max_depth > 0 && it.depth >= (uint32_t)prm->max_depth) || N < mss ||
                   N < 2ull * msl || pure;
        int found = 0; int32_t bf = 0, bb = 0; double bg = 0.0;
        if (!leaf) {
            st->searched++;
            st->row_visits += it.end - it.begin;
            for (int32_t f = 0; f < d; ++f) perm[f] = f;
            /* Fisher-Yates draw without replacement keyed by (begin, depth, tree):
             * SK/tree/_splitter.pyx:323-389 restated with a counter-based stream. */
            uint32_t rnd[4] = {0, 0, 0, 0};
            for (int32_t i = 0; i < d; ++i) {
                if (i >= m && found) break; /* fallback stops at the first valid feature */
                if ((i & 3) == 0)
                    philox4x32_10(it.begin, it.depth, gid, ORC_PURPOSE_FEAT | ((uint32_t)(i >> 2) << 8), k0, k1, rnd);
                int32_t j = i + (int32_t)mulhi32(rnd[i & 3], (uint32_t)(d - i));
                int32_t f = perm[j]; perm[j] = perm[i]; perm[i] = f;
                if (i == m) st->fallback++;
                const uint8_t* col = bins_cm + (size_t)f * n;
                int32_t lo = 255, hi = 0;
                for (uint32_t r = it.begin; r < it.end; ++r) {
                    uint8_t b = col[ent[r].row];
                    hist[(size_t)b * C + y[ent[r].row]] += ent[r].w;
                    if (b < lo) lo = b;
                    if (b > hi) hi = b;
                }
                for (int32_t c = 0; c < C; ++c) cum[c] = 0;
                uint64_t NL = 0;
                int f_found = 0; int32_t f_bin = 0; double f_gain = 0.0;
                for (int32_t b = lo; b < hi; ++b) { /* b == hi leaves the right side empty */
                    uint64_t inb = 0;
                    for (int32_t c = 0; c < C; ++c) { uint32_t h = hist[(size_t)b * C + c]; cum[c] += h; inb += h; }
                    NL += inb;
                    if (!inb) continue; /* same partition as the previous occupied bin: loses the tie */
                    uint64_t NR = N - NL;
                    if (NL < msl || NR < msl) continue;
                    uint64_t SL = 0, SR = 0;
                    double g;
                    if (prm->criterion == 1) {
                        for (int32_t c = 0; c < C; ++c) { uint64_t l = cum[c], r = tot[c] - l; SL += (uint64_t)c * l; SR += (uint64_t)c * r; }
                        g = gain_mse(SL, NL, SR, NR);
                    } else {
                        for (int32_t c = 0; c < C; ++c) { uint64_t l = cum[c], r = tot[c] - l; SL += l * l; SR += r * r; }
                        g = gain_of(SL, NL, SR, NR);
                    }
                    if (!f_found || g > f_gain) { f_found = 1; f_gain = g; f_bin = b; }
                }
                for (int32_t b = lo; b <= hi; ++b) for (int32_t c = 0; c < C; ++c) hist[(size_t)b * C + c] = 0;
                if (f_found) {
                    if (i < m) {
                        /* max gain; ties -> the feature drawn first, then the lowest bin: sklearn keeps the first best in draw
                         * order (strict >, SK/tree/_splitter.pyx:445).  A lowest-feature-INDEX rule biases the many exact ties of
                         * small nodes, and with them the importances, toward low-index features (+-30 % measured). */
                        if (!found || f_gain > bg) { found = 1; bg = f_gain; bf = f; bb = f_bin; }
                    } else {
                        found = 1; bg = f_gain; bf = f; bb = f_bin;
                    }
                }
            }
        }
        if (leaf || !found) { st->leaves++; continue; }
        if (imp) {
            uint64_t Sp = 0;
            double parent;
            if (prm->criterion == 1) {
                for (int32_t c = 0; c < C; ++c) Sp += (uint64_t)c * tot[c];
                volatile double sp = (double)Sp; volatile double sq = sp * sp;
                parent = sq / (double)N;
            } else {
                for (int32_t c = 0; c < C; ++c) Sp += tot[c] * tot[c];
                parent = (double)Sp / (double)N;
            }
            imp[bf] += imp_q24(bg, parent);
        }
        /* stable partition: SK/tree/_partitioner.pyx:248-279 restated out-of-place */
        const uint8_t* col = bins_cm + (size_t)bf * n;
        uint32_t nl = 0, nr = 0;
        for (uint32_t r = it.begin; r < it.end; ++r) {
            if (col[ent[r].row] <= bb) ent[it.begin + nl++] = ent[r]; else tmp[nr++] = ent[r];
        }
        memcpy(ent + it.begin + nl, tmp, (size_t)nr * sizeof(entry_t));
        nd = &S->nodes[id];
        nd->feature = bf; nd->bin = bb;
        if (sp + 2 > scap) { scap *= 2; stk = (stack_item*)realloc(stk, scap * sizeof(stack_item)); }
        stk[sp++] = (stack_item){it.begin + nl, it.end, it.depth + 1, (int32_t)id, 1};
        stk[sp++] = (stack_item){it.begin, it.begin + nl, it.depth + 1, (int32_t)id, 0};
    }
    free(hist); free(tot); free(cum); free(perm); free(stk);
    return 0;
}

/* Criterion 2: the same growth loop on (count, sum y) histograms.  A node is a leaf when all its targets are
 * equal (sklearn: impurity <= EPSILON, SK/tree/_tree.pyx:236), N < min_samples_split, N < 2 * min_samples_leaf, depth limit,
 * or no valid split.  gain = S_L^2 / N_L + S_R^2 / N_R (MSE.proxy_impurity_improvement, SK/tree/_criterion.pyx:1074-1096) with
 * S = sum of w * yq exact in uint64.  counts[node] = (S low 32 bits, S high 32 bits); n_weighted = N. */
static int grow_tree_reg(const uint8_t* bins_cm, int64_t n, int32_t d, const uint16_t* yq, const orc_params* prm, uint32_t k0,
                         uint32_t k1, uint32_t gid, entry_t* ent, entry_t* tmp, uint32_t n_inbag, node_store* S, orc_stats* st,
                         uint64_t* imp) {
    int32_t m = prm->max_features;
    uint32_t msl = (uint32_t)prm->min_samples_leaf, mss = (uint32_t)prm->min_samples_split;
    uint64_t* hcnt = (uint64_t*)calloc(ORC_NBINS, 8);
    uint64_t* hsum = (uint64_t*)calloc(ORC_NBINS, 8);
    int32_t* perm = (int32_t*)malloc((size_t)d * 4);
    size_t scap = 64, sp = 0;
    stack_item* stk = (stack_item*)malloc(scap * sizeof(stack_item));
    stk[sp++] = (stack_item){0, n_inbag, 0, -1, 0};
    while (sp) {
        stack_item it = stk[--sp];
        int64_t id = store_push(S, 2);
        orc_node* nd = &S->nodes[id];
        if (it.parent >= 0) {
            if (it.is_right) S->nodes[it.parent].right = (int32_t)id; else S->nodes[it.parent].left = (int32_t)id;
        }
        uint64_t N = 0, Sy = 0;
        uint32_t ymin = 0xFFFFFFFFu, ymax = 0;
        for (uint32_t i = it.begin; i < it.end; ++i) {
            uint32_t v = yq[ent[i].row];
            N += ent[i].w; Sy += (uint64_t)ent[i].w * v;
            if (v < ymin) ymin = v;
            if (v > ymax) ymax = v;
        }
        S->counts[(size_t)id * 2] = (uint32_t)Sy; S->counts[(size_t)id * 2 + 1] = (uint32_t)(Sy >> 32);
        nd->left = nd->right = -1; nd->feature = -1; nd->bin = -1;
        nd->begin = it.begin; nd->end = it.end; nd->depth = it.depth; nd->n_weighted = (uint32_t)N;
        st->nodes++;
        if (it.depth > st->max_depth) st->max_depth = it.depth;
        int pure = ymin == ymax;
        int leaf = (prm->max_depth > 0 && it.depth >= (uint32_t)prm->max_depth) || N < mss || N < 2ull * msl || pure;
        int found = 0; int32_t bf = 0, bb = 0; double bg = 0.0;
        if (!leaf) {
            st->searched++;
            st->row_visits += it.end - it.begin;
            for (int32_t f = 0; f < d; ++f) perm[f] = f;
            uint32_t rnd[4] = {0, 0, 0, 0};
            for (int32_t i = 0; i < d; ++i) {
                if (i >= m && found) break;
                if ((i & 3) == 0)
                    philox4x32_10(it.begin, it.depth, gid, ORC_PURPOSE_FEAT | ((uint32_t)(i >> 2) << 8), k0, k1, rnd);
                int32_t j = i + (int32_t)mulhi32(rnd[i & 3], (uint32_t)(d - i));
                int32_t f = perm[j]; perm[j] = perm[i]; perm[i] = f;
                if (i == m) st->fallback++;
                const uint8_t* col = bins_cm + (size_t)f * n;
                int32_t lo = 255, hi = 0;
                for (uint32_t r = it.begin; r < it.end; ++r) {
                    uint8_t b = col[ent[r].row];
                    hcnt[b] += ent[r].w; hsum[b] += (uint64_t)ent[r].w * yq[ent[r].row];
                    if (b < lo) lo = b;
                    if (b > hi) hi = b;
                }
                uint64_t NL = 0, SL = 0;
                int f_found = 0; int32_t f_bin = 0; double f_gain = 0.0;
                for (int32_t b = lo; b < hi; ++b) {
                    if (!hcnt[b]) continue;
                    NL += hcnt[b]; SL += hsum[b];
                    uint64_t NR = N - NL;
                    if (NL < msl || NR < msl) continue;
                    double g = gain_mse(SL, NL, Sy - SL, NR);
                    if (!f_found || g > f_gain) { f_found = 1; f_gain = g; f_bin = b; }
                }
                for (int32_t b = lo; b <= hi; ++b) { hcnt[b] = 0; hsum[b] = 0; }
                if (f_found) {
                    if (i < m) { if (!found || f_gain > bg) { found = 1; bg = f_gain; bf = f; bb = f_bin; } }
                    else { found = 1; bg = f_gain; bf = f; bb = f_bin; }
                }
            }
        }
        if (leaf || !found) { st->leaves++; continue; }
        if (imp) {
            volatile double sp2 = (double)Sy; volatile double sq = sp2 * sp2;
            imp[bf] += imp_q24(bg, sq / (double)N);
        }
        const uint8_t* col = bins_cm + (size_t)bf * n;
        uint32_t nl = 0, nr = 0;
        for (uint32_t r = it.begin; r < it.end; ++r) {
            if (col[ent[r].row] <= bb) ent[it.begin + nl++] = ent[r]; else tmp[nr++] = ent[r];
        }
        memcpy(ent + it.begin + nl, tmp, (size_t)nr * sizeof(entry_t));
        nd = &S->nodes[id];
        nd->feature = bf; nd->bin = bb;
        if (sp + 2 > scap) { scap *= 2; stk = (stack_item*)realloc(stk, scap * sizeof(stack_item)); }
        stk[sp++] = (stack_item){it.begin + nl, it.end, it.depth + 1, (int32_t)id, 1};
        stk[sp++] = (stack_item){it.begin, it.begin + nl, it.depth + 1, (int32_t)id, 0};
    }
    free(hcnt); free(hsum); free(perm); free(stk);
    return 0;
}

/* ---------------------------------------------------------------- fit + OOF votes
 * Trees are identified by (part = group*k + fold, t).  RNG id gid = gid_base +
 * part*n_trees + t.  This rank builds trees with (part*n_trees + t) % world == rank,
 * t in [t0, t1).  Votes: per tree and held-out row, class c receives
 * (count_c << 16) / N  (Q0.16 soft vote; SK/tree/_criterion.pyx:483-486,
 * SK/ensemble/_forest.py:961-962), accumulated as uint32. */
int orc_fit_oof(const uint8_t* bins_cm, int64_t n, int32_t d, const uint8_t* y, const uint8_t* fold,
                const int32_t* group, int32_t n_groups, int32_t k, int32_t n_trees, int32_t C,
                const orc_params* prm, uint64_t seed, uint32_t gid_base, int32_t t0, int32_t t1,
                int32_t rank, int32_t world, uint32_t* votes, orc_stats* stats, orc_forest** keep,
                int32_t nthreads, uint64_t* imp_sums /* [d + 1] or NULL: sum over trees of floor((double)imp_t[f] / (double)sum_f imp_t[f] * 2^32); [d] = trees with a split */) {
    if (n < 0 || n >= 0x7FFFFFF0LL || d < 1 || C < 1 || C > 256 || k < 1 || k > 250 || n_trees < 1 ||
        n_trees > 65535 || t0 < 0 || t1 > n_trees || t0 > t1 || world < 1 || rank < 0 || rank >= world ||
        prm->max_features < 1 || prm->max_features > d || prm->min_samples_leaf < 1 || prm->min_samples_split < 2 ||
        prm->criterion < 0 || prm->criterion > 2 || (prm->criterion == 2 && (!g_target16 || C != 2)))
        return -1;
    orc_parts P;
    int rc = build_parts(fold, group, n_groups, n, k, &P);
    if (rc) return rc;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    int32_t n_parts = n_groups * k;
    orc_forest* F = NULL;
    if (keep) {
        F = (orc_forest*)calloc(1, sizeof(orc_forest));
        F->n_trees_total = n_parts * n_trees; F->C = C;
        F->trees = (orc_tree*)calloc((size_t)F->n_trees_total, sizeof(orc_tree));
        *keep = F;
    }
    orc_stats total; memset(&total, 0, sizeof(total));
    int err = 0;
    int64_t n_tasks = (int64_t)n_parts * (t1 - t0);
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
    {
        orc_stats st; memset(&st, 0, sizeof(st));
        uint32_t* mult = (uint32_t*)calloc((size_t)(n > 0 ? n : 1), 4);
        entry_t* ent = (entry_t*)malloc((size_t)(n > 0 ? n : 1) * sizeof(entry_t));
        entry_t* tmp = (entry_t*)malloc((size_t)(n > 0 ? n : 1) * sizeof(entry_t));
        uint64_t* imp = imp_sums ? (uint64_t*)malloc((size_t)d * 8) : NULL;
        uint64_t* imp_loc = imp_sums ? (uint64_t*)calloc((size_t)d + 1, 8) : NULL;
#pragma omp for schedule(dynamic, 1)
        for (int64_t task = 0; task < n_tasks; ++task) {
            int32_t part = (int32_t)(task / (t1 - t0));
            int32_t t = t0 + (int32_t)(task % (t1 - t0));
            uint32_t lid = (uint32_t)part * (uint32_t)n_trees + (uint32_t)t;
            if ((int32_t)(lid % (uint32_t)world) != rank) continue;
            uint32_t gid = gid_base + lid;
            int32_t g = part / k, f = part % k;
            uint32_t gs = P.part_start[(size_t)g * (k + 1)], ge = P.part_start[(size_t)(g + 1) * (k + 1)];
            uint32_t fs = P.part_start[(size_t)g * (k + 1) + f], fe = P.part_start[(size_t)g * (k + 1) + f + 1];
            uint32_t n_train = (ge - gs) - (fe - fs);
            uint32_t a = fs - gs, fsz = fe - fs;
            /* bootstrap: SK/ensemble/_forest.py:95-113,155-156 with a Philox stream */
            uint32_t rnd[4];
            for (uint32_t j = 0; j < n_train; ++j) {
                if ((j & 3) == 0) philox4x32_10(j >> 2, 0, gid, ORC_PURPOSE_BOOT, k0, k1, rnd);
                uint32_t u = mulhi32(rnd[j & 3], n_train);
                uint32_t row = P.sorted[u < a ? gs + u : gs + u + fsz];
                mult[row]++;
            }
            /* in-bag rows ascending (zero-weight rows dropped: SK/tree/_splitter.pyx:167-171) */
            uint32_t ni = 0;
            int bad = 0;
            for (int64_t i = 0; i < n; ++i) {
                if (mult[i]) {
                    if (mult[i] > 255) bad = 1;
                    ent[ni].row = (uint32_t)i; ent[ni].w = mult[i]; ni++; mult[i] = 0;
                }
            }
            if (bad) {
#pragma omp atomic write
                err = -3;
                continue;
            }
            node_store S; memset(&S, 0, sizeof(S));
            if (imp) memset(imp, 0, (size_t)d * 8);
            if (prm->criterion == 2) grow_tree_reg(bins_cm, n, d, g_target16, prm, k0, k1, gid, ent, tmp, ni, &S, &st, imp);
            else grow_tree(bins_cm, n, d, y, C, prm, k0, k1, gid, ent, tmp, ni, &S, &st, imp);
            st.trees++; st.inbag_rows += ni;
            if (imp) {  /* per-tree normalisation, quantised to Q32: the forest-level figure is an integer sum of these, in any order */
                uint64_t tsum = 0;
                for (int32_t f2 = 0; f2 < d; ++f2) tsum += imp[f2];
                if (tsum) {
                    for (int32_t f2 = 0; f2 < d; ++f2) {
                        volatile double q = (double)imp[f2] / (double)tsum;  /* both < 2^53: exact conversions, one IEEE division */
                        volatile double sc = q * 4294967296.0;
                        imp_loc[f2] += (uint64_t)floor(sc);
                    }
                    imp_loc[d]++;
                }
            }
            /* OOF inference: SK/tree/_tree.pyx:954-996 on bins */
            if (votes) {
                for (uint32_t q = fs; q < fe; ++q) {
                    uint32_t row = P.sorted[q];
                    int32_t nd = 0;
                    while (S.nodes[nd].feature >= 0) {
                        uint8_t b = bins_cm[(size_t)S.nodes[nd].feature * n + row];
                        nd = (b <= S.nodes[nd].bin) ? S.nodes[nd].left : S.nodes[nd].right;
                        st.oof_steps++;
                    }
                    uint64_t N = S.nodes[nd].n_weighted;
                    if (!N) continue;
                    if (prm->criterion == 2) {  /* Q16.16 leaf mean, added as two 16-bit halves: no carries between the words */
                        uint64_t Sy = (uint64_t)S.counts[(size_t)nd * 2] | ((uint64_t)S.counts[(size_t)nd * 2 + 1] << 32);
                        uint32_t v = (uint32_t)((Sy << 16) / N);
#pragma omp atomic
                        votes[(size_t)row * 2] += v & 0xFFFFu;
#pragma omp atomic
                        votes[(size_t)row * 2 + 1] += v >> 16;
                        continue;
                    }
                    for (int32_t c = 0; c < C; ++c) {
                        uint32_t v = (uint32_t)(((uint64_t)S.counts[(size_t)nd * C + c] << 16) / N);
                        if (v) {
#pragma omp atomic
                            votes[(size_t)row * C + c] += v;
                        }
                    }
                }
            }
            if (F) {
                orc_tree* T = &F->trees[lid];
                T->n_nodes = S.n_nodes; T->nodes = S.nodes; T->counts = S.counts; T->part = part; T->tree = t;
            } else {
                free(S.nodes); free(S.counts);
            }
        }
#pragma omp critical
        {
            total.row_visits += st.row_visits; total.nodes += st.nodes; total.leaves += st.leaves;
            total.searched += st.searched; total.fallback += st.fallback; total.inbag_rows += st.inbag_rows;
            total.trees += st.trees; total.oof_steps += st.oof_steps;
            if (st.max_depth > total.max_depth) total.max_depth = st.max_depth;
            if (imp_sums) for (int32_t f2 = 0; f2 <= d; ++f2) imp_sums[f2] += imp_loc[f2];
        }
        free(mult); free(ent); free(tmp); free(imp); free(imp_loc);
    }
    if (stats) *stats = total;
    free(P.sorted); free(P.part_start);
    return err;
}

/* ---------------------------------------------------------------- canonical export
 * Nodes sorted by (depth, begin) = breadth-first, left before right. */
static int cmp_node_idx(const void* a, const void* b, void* ctx) {
    const orc_node* nodes = (const orc_node*)ctx;
    const orc_node* x = &nodes[*(const int64_t*)a];
    const orc_node* y = &nodes[*(const int64_t*)b];
    if (x->depth != y->depth) return x->depth < y->depth ? -1 : 1;
    if (x->begin != y->begin) return x->begin < y->begin ? -1 : 1;
    return 0;
}

int64_t orc_tree_size(const orc_forest* F, int32_t part, int32_t n_trees, int32_t t) {
    return F->trees[(size_t)part * n_trees + t].n_nodes;
}

int orc_export_tree(const orc_forest* F, int32_t part, int32_t n_trees, int32_t t, orc_node* out,
                    uint32_t* counts_out) {
    const orc_tree* T = &F->trees[(size_t)part * n_trees + t];
    int64_t m = T->n_nodes;
    if (!m) return -1;
    int64_t* order = (int64_t*)malloc((size_t)m * 8);
    int64_t* inv = (int64_t*)malloc((size_t)m * 8);
    for (int64_t i = 0; i < m; ++i) order[i] = i;
    qsort_r(order, (size_t)m, 8, cmp_node_idx, T->nodes);
    for (int64_t i = 0; i < m; ++i) inv[order[i]] = i;
    for (int64_t i = 0; i < m; ++i) {
        orc_node nd = T->nodes[order[i]];
        if (nd.left >= 0) { nd.left = (int32_t)inv[nd.left]; nd.right = (int32_t)inv[nd.right]; }
        out[i] = nd;
        memcpy(counts_out + (size_t)i * F->C, T->counts + (size_t)order[i] * F->C, (size_t)F->C * 4);
    }
    free(order); free(inv);
    return 0;
}

void orc_free_forest(orc_forest* F) {
    if (!F) return;
    for (int32_t i = 0; i < F->n_trees_total; ++i) { free(F->trees[i].nodes); free(F->trees[i].counts); }
    free(F->trees); free(F);
}

/* Apply a kept forest (k == 1 style predict): votes over all trees for arbitrary binned rows. */
int orc_predict(const orc_forest* F, const uint8_t* bins_cm, int64_t n, uint32_t* votes) {
    for (int32_t ti = 0; ti < F->n_trees_total; ++ti) {
        const orc_tree* T = &F->trees[ti];
        if (!T->n_nodes) continue;
        for (int64_t row = 0; row < n; ++row) {
            int32_t nd = 0;
            while (T->nodes[nd].feature >= 0) {
                uint8_t b = bins_cm[(size_t)T->nodes[nd].feature * n + row];
                nd = (b <= T->nodes[nd].bin) ? T->nodes[nd].left : T->nodes[nd].right;
            }
            uint64_t N = T->nodes[nd].n_weighted;
            if (!N) continue;
            for (int32_t c = 0; c < F->C; ++c)
                votes[(size_t)row * F->C + c] += (uint32_t)(((uint64_t)T->counts[(size_t)nd * F->C + c] << 16) / N);
        }
    }
    return 0;
}

/* ---------------------------------------------------------------- s(x) and accuracy
 * s_i = votes[i][y_i] / (T * 65536) computed in double, stored as float; inactive -> -1.
 * correct/total per partition are integer counts (argmax ties -> lowest class). */
int orc_finalize(const uint32_t* votes, const uint8_t* y, const uint8_t* fold, const int32_t* group,
                 int64_t n, int32_t C, int32_t k, int32_t n_trees, float* s, uint32_t* part_correct,
                 uint32_t* part_total) {
    double denom = (double)n_trees * 65536.0;
    for (int64_t i = 0; i < n; ++i) {
        if (fold[i] >= k) { if (s) s[i] = -1.0f; continue; }
        const uint32_t* v = votes + (size_t)i * C;
        if (s) s[i] = (float)((double)v[y[i]] / denom);
        int32_t best = 0;
        for (int32_t c = 1; c < C; ++c) if (v[c] > v[best]) best = c;
        int32_t part = (group ? group[i] : 0) * k + fold[i];
        if (part_total) part_total[part]++;
        if (part_correct && best == y[i]) part_correct[part]++;
    }
    return 0;
}

/* ---------------------------------------------------------------- IBP round selection
 * candidates = active rows with v_i = votes[i][y_i] >= vmin; remove the
 * `batch` largest by (v desc, row asc).  Returns the number removed. */
typedef struct { uint32_t v; uint32_t row; } cand_t;
static int cmp_cand(const void* a, const void* b) {
    const cand_t* x = (const cand_t*)a; const cand_t* y = (const cand_t*)b;
    if (x->v != y->v) return x->v > y->v ? -1 : 1;
    if (x->row != y->row) return x->row < y->row ? -1 : 1;
    return 0;
}

int64_t orc_ibp_select(const uint32_t* votes, const uint8_t* y, const uint8_t* fold, int64_t n, int32_t C,
                       int32_t k, uint64_t vmin, int64_t batch, uint8_t* active) {
    cand_t* c = (cand_t*)malloc((size_t)(n > 0 ? n : 1) * sizeof(cand_t));
    int64_t m = 0;
    for (int64_t i = 0; i < n; ++i) {
        if (!active[i] || fold[i] >= k) continue;
        uint32_t v = votes[(size_t)i * C + y[i]];
        if ((uint64_t)v >= vmin) { c[m].v = v; c[m].row = (uint32_t)i; m++; }
    }
    qsort(c, (size_t)m, sizeof(cand_t), cmp_cand);
    int64_t take = m < batch ? m : batch;
    for (int64_t i = 0; i < take; ++i) active[c[i].row] = 0;
    free(c);
    return take;
}

/* Full IBP loop (REF/README.md:35 "Iterative Bias Pruning"). */
int orc_ibp(const uint8_t* bins_cm, int64_t n, int32_t d, const uint8_t* y, uint8_t* active,
            const int32_t* group, int32_t n_groups, int32_t k, int32_t n_trees, int32_t C,
            const orc_params* prm, uint64_t seed, int32_t rounds, double tau, int64_t batch,
            int64_t* removed_per_round, float* s_per_round /* [rounds][n] or NULL */, int32_t nthreads) {
    uint8_t* fold = (uint8_t*)malloc((size_t)(n > 0 ? n : 1));
    uint32_t* votes = (uint32_t*)malloc((size_t)(n > 0 ? n : 1) * C * 4);
    uint64_t vmin = (uint64_t)ceil(tau * (double)n_trees * 65536.0);
    uint32_t per_round = (uint32_t)n_groups * (uint32_t)k * (uint32_t)n_trees;
    int rc = 0;
    for (int32_t r = 0; r < rounds; ++r) {
        removed_per_round[r] = 0;
        if (s_per_round) for (int64_t i = 0; i < n; ++i) s_per_round[(size_t)r * n + i] = -1.0f;
    }
    for (int32_t r = 0; r < rounds; ++r) {
        rc = orc_assign_folds(y, active, group, n_groups, n, k, C, fold);
        if (rc) break;
        memset(votes, 0, (size_t)n * C * 4);
        rc = orc_fit_oof(bins_cm, n, d, y, fold, group, n_groups, k, n_trees, C, prm, seed,
                         (uint32_t)r * per_round, 0, n_trees, 0, 1, votes, NULL, NULL, nthreads, NULL);
        if (rc) break;
        if (s_per_round) orc_finalize(votes, y, fold, group, n, C, k, n_trees, s_per_round + (size_t)r * n, NULL, NULL);
        int64_t took = orc_ibp_select(votes, y, fold, n, C, k, vmin, batch, active);
        removed_per_round[r] = took;
        if (!took) break;
    }
    free(fold); free(votes);
    return rc;
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
