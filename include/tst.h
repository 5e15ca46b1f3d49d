/*
 * include/tst.h -- C ABI of libtst.so, the B200-native TsT Random-Forest diagnostic.
 *
 * Drop-in boundary.  The reference repository ships no code
 * (REF/README.md:38-39 "Code: Coming soon!"), so there is no reference FFI to mirror;
 * every entry point below cites the sentence of REF/README.md that names the step and
 * the scikit-learn 1.9.0 code (SK/ = site-packages/sklearn) whose arithmetic it replaces.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no torch / C++ types cross the boundary.
 *   - *_dev pointers are caller-owned device memory (e.g. a torch tensor's data_ptr())
 *     valid for the duration of the call; *_host pointers are host memory.
 *   - every call returns 0 or a negative tst_status; the message is available from the
 *     thread-local tst_last_error().  No C++ exception crosses the ABI, CUDA errors are
 *     mapped, nothing calls abort().
 *   - work is enqueued on the caller's cudaStream_t (`stream`, may be NULL = default
 *     stream).  tst_fit_oof / tst_ibp synchronise that stream internally at tree-level
 *     boundaries (they read back queue counters), so they return with the work done.
 *   - a handle is not thread-safe: one handle per (process, device).
 *   - there is NO CPU fallback: every call fails with TST_ERR_CUDA when no device is present.
 */
#ifndef TST_H_
#define TST_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TST_ABI_VERSION 3
#define TST_MAX_EDGES 255        /* per feature; bins 0..254 + bin 255 for NaN */
#define TST_MAX_CLASSES 32
#define TST_MAX_FEATURES 1024
#define TST_MAX_ROWS 0x7FFFFFF0  /* row ids are 32-bit; the tree-entry is (row, multiplicity | class << 8) */
#define TST_FOLD_TRAIN_ONLY 254  /* row trains every tree, is never held out (plain fit) */
#define TST_FOLD_INACTIVE 255    /* row pruned by IBP / not in the table */

typedef enum {
    TST_OK = 0,
    TST_ERR_INVALID = -1,     /* bad argument */
    TST_ERR_CUDA = -2,        /* CUDA runtime error (message has the cudaError string) */
    TST_ERR_NOMEM = -3,       /* device allocation failed / forest too large to keep */
    TST_ERR_STATE = -4,       /* call order (e.g. fit before bin) */
    TST_ERR_OVERFLOW = -5,    /* bootstrap multiplicity > 255 or a work queue overflowed */
    TST_ERR_UNSUPPORTED = -6
} tst_status;

typedef struct tst_handle tst_handle;

/* Growth parameters; names and defaults follow RandomForestClassifier
 * (SK/ensemble/_forest.py:1515-1537, SK/tree/_classes.py:315-345). */
typedef struct {
    int32_t max_depth;          /* <= 0: unlimited */
    int32_t min_samples_split;  /* >= 2, on bootstrap-weighted counts */
    int32_t min_samples_leaf;   /* >= 1, on bootstrap-weighted counts */
    int32_t max_features;       /* m in [1, d]; host resolves "sqrt" -> max(1, floor(sqrt(d))) */
    int32_t trees_in_flight;    /* <= 0: sized from free HBM */
    int32_t keep_forest;        /* != 0: node arrays stay resident for tst_export_tree / tst_predict */
    int32_t criterion;          /* TST_GINI (0); TST_SQUARED_ERROR_LEVELS (1): the labels 0..C-1 are the ordinal levels of a
                                   uniformly quantised numeric target and splits maximise S_L^2/N_L + S_R^2/N_R with
                                   S = sum_c c * count_c; TST_SQUARED_ERROR (2): the same proxy (SK/tree/_criterion.pyx:1074-1096,
                                   MSE.proxy_impurity_improvement) on a CONTINUOUS target in 16-bit fixed point (tst_set_target):
                                   per-bin (weight, sum y) histograms with exact integer sums, leaves hold (N, sum y), n_classes
                                   must be 2 and votes_dev [n, 2] collects the low / high 16-bit halves of the Q16.16 leaf means
                                   (prediction = (votes[1] * 65536 + votes[0]) / (n_trees * 65536)); y_dev then only stratifies folds */
    int32_t reserved;
} tst_params;

enum { TST_GINI = 0, TST_SQUARED_ERROR_LEVELS = 1, TST_SQUARED_ERROR = 2 };

typedef struct {
    uint64_t row_visits_big;    /* distinct in-bag rows summed over nodes searched by the per-node CTA kernel */
    uint64_t row_visits_small;  /* ... by the warp-per-subtree kernel */
    uint64_t nodes, leaves, searched, fallback;
    uint64_t inbag_rows, trees, oof_steps;
    uint64_t kernel_launches;   /* kernels launched by this handle since the last tst_reset_stats */
    uint64_t levels;            /* level-synchronous rounds */
    uint32_t max_depth;
    uint32_t pad;
    double ms_grow_big;         /* CUDA-event time spent in the dominant kernel (k_grow_big) */
    double ms_grow_small;
    double ms_prep;             /* bootstrap + compaction */
    double ms_infer;
    uint64_t launches_grow_big;
    uint64_t launches_grow_small;
    double ms_grow_mid;         /* warp-per-node kernel (k_grow_mid) */
    uint64_t launches_grow_mid;
    uint64_t row_visits_mid;    /* ... by the warp-per-node kernel */
    double ms_grow_tiny;        /* thread-per-subtree kernel (k_grow_tiny4); its visits are counted in row_visits_small */
    uint64_t launches_grow_tiny;
    uint64_t launches_root;     /* dense root kernels (k_root_hist + k_root_split), timed inside ms_prep */
    uint64_t row_visits_sub;    /* ... by the CTA-per-subtree kernel (k_grow_sub: rows cached in shared memory) */
    double ms_grow_sub;
    uint64_t launches_grow_sub;
    uint64_t row_visits_wide;   /* ... by the four-warps-per-node kernel (k_grow_wide: 257 .. 1024-row nodes, rows staged once in shared memory, warp-privatised histograms) */
    double ms_grow_wide;
    uint64_t launches_grow_wide;
} tst_stats;

/* Canonical tree export: nodes in breadth-first order, i.e. sorted by (depth, begin). */
typedef struct {
    int32_t left, right;   /* node indices, -1 for leaves */
    int32_t feature;       /* original feature index, -1 for leaves */
    int32_t bin;           /* row goes left iff bin[row][feature] <= bin; raw threshold = edges[feature][bin] */
    uint32_t begin, end;   /* segment of the tree's in-bag row list */
    uint32_t depth;
    uint32_t n_weighted;   /* sum of bootstrap multiplicities */
} tst_node;

int tst_abi_version(void);
const char* tst_last_error(void);

/* Multi-GPU: one handle per (process, device).  nccl_unique_id (128 bytes from tst_nccl_unique_id on rank 0,
 * distributed by the host) makes the handle join an NCCL communicator of `world` ranks (libnccl.so.2 is
 * dlopen'ed: no link-time dependency); all ranks must call tst_create together.  With NULL the handle has no
 * communicator and the host does the vote all-gather itself before tst_reduce_votes. */
int tst_nccl_unique_id(void* out128);
int tst_create(tst_handle** out, int device, int rank, int world, const void* nccl_unique_id);
/* (C1 + K9) In place: ONE ncclAllGather of the [n_values] uint32 vote tensor over NVLink, then the rank-ordered
 * integer sum.  Collective: every rank calls it.  No-op for world == 1. */
int tst_allreduce_votes(tst_handle* h, uint32_t* votes_dev, int64_t n_values, void* stream);
int tst_destroy(tst_handle* h);

/* (1) Quantile binning -- "hand-crafted features" table ingest (REF/README.md:35); replaces
 * SK/ensemble/_hist_gradient_boosting/_binning.pyx:50-85 (map) given edges computed as
 * binning.py:31-111.  X_dev: [n, d] row-major float32.  edges_host: [d][TST_MAX_EDGES] float64.
 * Builds both the column-major [d][ld] and row-major [n][stride] uint8 tables in HBM. */
int tst_bin(tst_handle* h, const float* X_dev, int64_t n, int32_t d, const double* edges_host,
            const int32_t* n_edges_host, void* stream);
/* Same, with the bin edges selected on the device (per-column radix sort + integer-rank quantiles, the
 * rule of DESIGN.md 2.1 / SK/.../binning.py:31-111): nothing but X crosses PCIe.  n_bins in [3, 256].
 * edges_out_host [d][TST_MAX_EDGES] / n_edges_out_host [d] receive the chosen edges (may be NULL). */
int tst_bin_auto(tst_handle* h, const float* X_dev, int64_t n, int32_t d, int32_t n_bins, double* edges_out_host,
                 int32_t* n_edges_out_host, void* stream);
/* The bin edges of the current table (as chosen by tst_bin_auto or given to tst_bin): edges_out_host [d][TST_MAX_EDGES],
 * n_edges_out_host [d].  Synchronises with the binning.  tst_bin_auto(..., NULL, NULL, stream) + this call later lets the host
 * queue the fit while the table is still being copied and binned on another stream: tst_fit_oof waits for the binned table
 * on the device, after its bootstrap kernels (which only need labels and folds). */
int tst_get_edges(tst_handle* h, double* edges_out_host, int32_t* n_edges_out_host);
/* Same, from an already binned column-major table [d][n] on the device (tests, pre-binned data). */
int tst_set_bins(tst_handle* h, const uint8_t* bins_cm_dev, int64_t n, int32_t d, void* stream);
/* Copy the column-major table out: out_dev [d][n]. */
int tst_get_bins(tst_handle* h, uint8_t* out_dev, void* stream);

/* Numeric answers: the continuous target of criterion TST_SQUARED_ERROR, yq_dev [n] uint16 =
 * round((y - min) / (max - min) * 65535), copied into the handle; valid until the table is re-binned.  Replaces the float64 y
 * of RandomForestRegressor.fit (SK/ensemble/_forest.py:303-522, SK/tree/_criterion.pyx:707-1000). */
int tst_set_target(tst_handle* h, const uint16_t* yq_dev, int64_t n, void* stream);

/* (A3) Stratified k-fold assignment on the device -- "k-fold cross-validation" (REF/README.md:35);
 * replaces SK/model_selection/_split.py:774-840.  group_dev may be NULL (one group). */
int tst_assign_folds(tst_handle* h, const uint8_t* y_dev, const uint8_t* active_dev, const int32_t* group_dev,
                     int32_t n_groups, int32_t k, int32_t n_classes, uint8_t* fold_dev, void* stream);

/* (2)-(6) Fit this rank's share of the k-fold x n_trees forests and accumulate the out-of-fold
 * Q0.16 votes of the trees it built into votes_dev [n, C] uint32 (not zeroed here).  Replaces
 * RandomForestClassifier.fit / predict_proba per fold (SK/ensemble/_forest.py:303-522,921-967,
 * SK/tree/_tree.pyx:139-337,939-996, _splitter.pyx:262-504, _criterion.pyx:147-199,302-486).
 * Trees are (part = group*k + fold, t), t in [t0, t1); RNG id = gid_base + part*n_trees + t; the
 * rank owns the trees with (part*n_trees + t) % world == rank. */
int tst_fit_oof(tst_handle* h, const uint8_t* y_dev, const uint8_t* fold_dev, const int32_t* group_dev,
                int32_t n_groups, int32_t k, int32_t n_trees, int32_t n_classes, const tst_params* params,
                uint64_t seed, uint32_t gid_base, int32_t t0, int32_t t1, uint32_t* votes_dev, void* stream);

/* The whole diagnostic in one call on device buffers (the shape SURVEY 8(b) sketched for tst_fit_oof: active mask in, votes and
 * s(x) out): stratified folds of the active rows -> fold_out_dev [n], this rank's share of the k x n_trees forests, the vote
 * all-gather when the handle owns a communicator, votes_dev [n, C] (zeroed here), s_dev [n] (may be NULL).  Collective when
 * world > 1.  Equivalent to tst_assign_folds + tst_fit_oof(t0 = 0, t1 = n_trees, gid_base = 0) + tst_allreduce_votes + tst_finalize. */
int tst_run_oof(tst_handle* h, const uint8_t* y_dev, const uint8_t* active_dev, const int32_t* group_dev, int32_t n_groups,
                int32_t k, int32_t n_trees, int32_t n_classes, const tst_params* params, uint64_t seed, uint8_t* fold_out_dev,
                uint32_t* votes_dev, float* s_dev, void* stream);

/* (C1/K9) Sum `world` gathered vote tensors [world][n*C] in rank order into votes_out_dev. */
int tst_reduce_votes(tst_handle* h, const uint32_t* gathered_dev, int32_t world, int64_t n_values,
                     uint32_t* votes_out_dev, void* stream);

/* (6) s(x) -- "assign each sample a bias score s(x)" (REF/README.md:35); replaces
 * SK/ensemble/_forest.py:961-962.  s_dev [n] float32 (-1 for rows not held out);
 * correct_dev / total_dev [n_groups*k] uint32 are zeroed then filled (may be NULL). */
int tst_finalize(tst_handle* h, const uint32_t* votes_dev, const uint8_t* y_dev, const uint8_t* fold_dev,
                 const int32_t* group_dev, int32_t n_groups, int32_t k, int32_t n_trees, int32_t n_classes,
                 float* s_dev, uint32_t* correct_dev, uint32_t* total_dev, void* stream);

/* (7) One IBP prune step on the device -- "filtering high-bias samples" (REF/README.md:35):
 * clear from active_dev the `batch` rows with the largest v = votes[i][y_i] >= vmin
 * (ties -> lowest row id); *removed_host receives the count. */
int tst_ibp_select(tst_handle* h, const uint32_t* votes_dev, const uint8_t* y_dev, const uint8_t* fold_dev,
                   int32_t k, int32_t n_classes, uint64_t vmin, int64_t batch, uint8_t* active_dev,
                   int64_t* removed_host, void* stream);

/* (7) Whole IBP loop kept on the device: rounds x (folds, fit, [vote all-reduce], s, prune).  With world > 1 the
 * handle needs a communicator and every rank calls it; all ranks end with the same active mask.
 * removed_host [rounds]; s_per_round_dev [rounds][n] float32 or NULL. */
int tst_ibp(tst_handle* h, const uint8_t* y_dev, uint8_t* active_dev, const int32_t* group_dev, int32_t n_groups,
            int32_t k, int32_t n_trees, int32_t n_classes, const tst_params* params, uint64_t seed,
            int32_t rounds, double tau, int64_t batch, int64_t* removed_host, float* s_per_round_dev,
            void* stream);

/* Kept forest (params.keep_forest): votes of every kept tree for new rows X_dev [n, d]
 * (binned with the handle's edges), accumulated into votes_dev [n, C].  Replaces
 * RandomForestClassifier.predict_proba (SK/ensemble/_forest.py:921-967). */
int tst_predict(tst_handle* h, const float* X_dev, int64_t n, uint32_t* votes_dev, void* stream);

/* Kept forest: canonical export of tree (part, t) for bit-exact comparison with the oracle.
 * nodes_host [cap], counts_host [cap][C]; *n_nodes receives the node count (also when cap is too small,
 * in which case TST_ERR_INVALID is returned and nothing is written). */
int tst_export_tree(tst_handle* h, int32_t part, int32_t t, tst_node* nodes_host, uint32_t* counts_host,
                    int64_t cap, int64_t* n_nodes);

/* (f3) "interpretable auditing" (REF/README.md:35): feature importances (mean decrease in impurity) of the trees this rank
 * built in the LAST tst_fit_oof / tst_ibp round, accumulated on the device while the trees grew -- no kept forest needed, so
 * they exist for the out-of-fold forests too.  Replaces SK/tree/_tree.pyx:1244-1280 + SK/ensemble/_forest.py feature_importances_.
 * sums_host [d]: sum over those trees of floor((double)imp_t[f] / (double)sum_f imp_t[f] * 2^32) with imp_t in Q24 fixed point -- integers, so
 * the sums of several ranks / calls add up in any order; forest importance = sums[f] / sum_f sums[f].  *n_counted = trees
 * with at least one split.  raw_host (optional, [raw_cap_trees][d]): the per-tree Q24 impurity decreases, in the order the
 * trees were enumerated (part-major, then t). */
int tst_feature_importances(tst_handle* h, uint64_t* sums_host, int64_t* n_counted, uint64_t* raw_host, int64_t raw_cap_trees);

/* Tuning / diagnostic knobs of one handle (INTEGRATION.md "knobs").  Defaults come from the environment variables
 * TST_<NAME>, read once in tst_create; names are case-insensitive, with or without the TST_ prefix. */
int tst_set_option(tst_handle* h, const char* name, int64_t value);

int tst_get_stats(tst_handle* h, tst_stats* out);
int tst_reset_stats(tst_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* TST_H_ */
