// tst_api.cu -- host side of libtst.so: the C ABI declared in include/tst.h.
// One translation unit; kernels live in the .cuh files next to it.  sm_100a only, no CPU fallback.
#include <dlfcn.h>
#include <nvtx3/nvToolsExt.h>  // header-only: ranges show up in Nsight Systems / Compute timelines, no-ops without a tool attached

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <chrono>
#include <vector>

#include "tst_common.cuh"
#include "tst_grow.cuh"
#include "tst_grow_small.cuh"
#include "tst_grow_mid.cuh"
#include "tst_grow_tiny.cuh"
#include "tst_grow_sub.cuh"
#include "tst_grow_wide.cuh"
#include "tst_root.cuh"
#include "tst_prep.cuh"
#include "tst_quantile.cuh"

// --------------------------------------------------------------------------------------------- errors
static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CK(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess)                                                                            \
            return fail(TST_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_) + " @" + __FILE__ + ":" + std::to_string(__LINE__)); \
    } while (0)
#define CKL() CK(cudaGetLastError())

struct NvtxRange {  // one host-side range per phase of a fit (tracing)
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

struct NvtxPhase {  // consecutive phases inside one scope; closes the open one on every return path
    bool open = false;
    void next(const char* name) { if (open) nvtxRangePop(); nvtxRangePushA(name); open = true; }
    ~NvtxPhase() { if (open) nvtxRangePop(); }
};

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return 0;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        size_t want = bytes + (bytes >> 3) + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            e = cudaMalloc(&p, bytes);
            want = bytes;
            if (e != cudaSuccess) { cudaGetLastError(); p = nullptr; return -1; }
        }
        cap = want;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return (T*)p; }
};
#define ENSURE(buf, bytes)                                                                  \
    do {                                                                                    \
        if ((buf).ensure(bytes)) return fail(TST_ERR_NOMEM, std::string("cudaMalloc of ") + std::to_string((size_t)(bytes)) + " bytes for " #buf " failed"); \
    } while (0)

struct KeptTree { int32_t part, t; uint32_t arena, n_inbag; };

// NCCL through dlopen: the library is optional at link time and shared with the host's own copy (torch) when present.
typedef struct { char internal[128]; } tst_nccl_id;
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(tst_nccl_id*) = nullptr;
    int (*CommInitRank)(void**, int, tst_nccl_id, int) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*CommGetAsyncError)(void*, int*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;
static int load_nccl() {
    if (g_nccl.lib) return 0;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* lib = nullptr;
    for (const char* nm : names) { lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (lib) break; }
    if (!lib) return -1;
    g_nccl.GetUniqueId = (int (*)(tst_nccl_id*))dlsym(lib, "ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(void**, int, tst_nccl_id, int))dlsym(lib, "ncclCommInitRank");
    g_nccl.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(lib, "ncclAllGather");
    g_nccl.CommDestroy = (int (*)(void*))dlsym(lib, "ncclCommDestroy");
    g_nccl.CommGetAsyncError = (int (*)(void*, int*))dlsym(lib, "ncclCommGetAsyncError");
    g_nccl.GetErrorString = (const char* (*)(int))dlsym(lib, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllGather || !g_nccl.CommDestroy) return -1;
    g_nccl.lib = lib;
    return 0;
}
#define NCCL_UINT32 3  /* ncclUint32 */

struct tst_handle {
    int device = 0, rank = 0, world = 1;
    int sm_count = 148;
    size_t l2_persist_max = 0, l2_window_max = 0;
    bool l2_persist = true;  // keep the row-major binned table resident in L2 while entry lists stream through (TST_L2_PERSIST=0 disables)
    // binned table
    int64_t n = 0, ld = 0;
    int32_t d = 0;
    uint32_t dstride = 0;
    bool have_edges = false;
    DevBuf bins_cm, bins_rm, edges, n_edges;
    // per-fit scratch
    DevBuf part_count, part_start, sorted, trees, counts8, chunk_cnt, chunk_sum, entA, entB, nodes, leafc, qbig0, qbig1,
        qmid0, qmid1, qmidc[2][2], qbigl[2], qwide0, qwide1, qsmall, qsub, qtiny, qtiny2, imp, imp_sums, counters, dstats, infer_parts, ghist, rsplit, y_pad, stash;
    // kept forest
    DevBuf keep_nodes, keep_leafc, keep_arenas;
    std::vector<KeptTree> kept;
    int32_t kept_C = 0, kept_n_trees = 0, kept_criterion = 0;
    DevBuf target16;          // criterion 2: the 16-bit fixed-point target (tst_set_target)
    int64_t target_n = 0;
    int64_t imp_trees = 0;  // trees of the last fit whose importances sit in `imp` ([imp_trees][d] uint64, Q24)
    int32_t imp_d = 0;
    uint64_t kept_node_stride = 0, kept_leafc_stride = 0;
    // ibp / predict scratch
    DevBuf ibp_fold, ibp_votes, ibp_removed, pred_cm, pred_rm, sort_a, sort_b, sort_dk, gather;
    void* comm = nullptr;  // ncclComm_t
    uint32_t* h_pinned = nullptr;  // 16 words
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, evm0 = nullptr, evm1 = nullptr, evs0 = nullptr, evs1 = nullptr;
    cudaEvent_t ev_lvl[8][4] = {};  // (the two launches of the warp-per-node tier are timed together)
    //  // per level of a group of blind levels: before big, after big, after wide, after mid
    cudaEvent_t evi0 = nullptr, evi1 = nullptr;  // OOF inference
    cudaEvent_t evk0 = nullptr, evk1 = nullptr;  // CTA-per-subtree tier
    // The table may be binned on another stream than the one the fit runs on (the host overlaps the H2D copy + binning with the
    // bootstrap of the first batch): every consumer of the binned table waits for this event on its own stream first.
    cudaEvent_t ev_bins = nullptr;
    cudaStream_t bins_stream = nullptr;
    bool bins_recorded = false;
    size_t l2_carve_set = 0, l2_carve_saved = 0;
    bool l2_carve_touched = false;
    // tuning / diagnostic knobs, read from the environment once in tst_create
    struct Knobs {
        bool trace = false, stash = false, big_packed = true;
        int root_dense = -1;          // -1 = automatic, 0 = off, 2 = on for any table size
        int tiny_max = -1, mid_max = -1, stash_max = -1;
        int blind_levels = 4;
        long thin_nodes = 2048;
        int mid_classes = -1;         // size classes of the warp-per-node tier with their own queue + kernel instantiation: bit 0 = 33..64 rows
                                      // (2 row slots per lane), bit 1 = ..128 rows (4 slots); 0 = one queue for the whole tier; default: 3 for batches of >= 2^24 (tree, row) pairs, else 0
        int small_multi = -1;         // the k_grow_small4 instantiation with 2 / 4 positions of the draw per lane group and pass: 1 on, 0 off, default: deep batches and criterion 2
        int big_lite = -1;            // light CTA-tier nodes (packed Gini) in their own queue + the LITE instantiation of k_grow_big: 1 on, 0 off, default: on for deep batches
        int wide_max = -1;            // largest node of the four-warps-per-node tier (0 = tier off; default 2048)
        int sub_rows = -1;            // row capacity of the CTA-per-subtree tier (<= 0 = tier off, the default)
    } knobs;
    tst_stats stats;
    bool attrs_set = false;
};

#ifndef BIG_THREADS_N
#define BIG_THREADS_N 256
#endif
static const int BIG_THREADS = BIG_THREADS_N;
static const int SMALL_WARPS = 4;
#ifndef TINY_MAX_ROWS
#define TINY_MAX_ROWS 6
#endif
static const int TINY_MAX = TINY_MAX_ROWS;   // thread-per-subtree tier: rows held in registers by one thread
#ifndef MID_WARPS_N
#define MID_WARPS_N 4  // with MID_MINB 5: 94 registers, no spills, 20 warps per SM: 192.6 ms per 640 trees vs 196.2 (8 warps x 3 CTAs, 80 registers, spills)
#endif
static const int MID_WARPS = MID_WARPS_N;
#ifndef MID_R_N
#define MID_R_N 8
#endif
static const int MID_R = MID_R_N;  // warp-per-node tier holds up to 32 * MID_R rows in registers
// size classes of that tier (GrowCtx::midc_max): 33..64 rows are every second node of it, ..128 another third; with 2 / 4 row slots per
// lane instead of 8 the kernel needs fewer registers and more warps are resident (it is latency bound: profiles/r02_k_grow_mid.md)
#ifndef MIDC0_MINB_N
#define MIDC0_MINB_N 8
#endif
#ifndef MIDC1_MINB_N
#define MIDC1_MINB_N 6
#endif
static const int MIDC_R[2] = {2, 4};
static const size_t BIG_SMEM_BUDGET = 64 * 1024;

extern "C" int tst_abi_version(void) { return TST_ABI_VERSION; }
extern "C" const char* tst_last_error(void) { return g_err.c_str(); }

extern "C" int tst_nccl_unique_id(void* out128) {
    if (!out128) return fail(TST_ERR_INVALID, "tst_nccl_unique_id: null argument");
    if (load_nccl()) return fail(TST_ERR_UNSUPPORTED, "tst_nccl_unique_id: libnccl.so.2 not found");
    int rc = g_nccl.GetUniqueId((tst_nccl_id*)out128);
    if (rc) return fail(TST_ERR_CUDA, std::string("ncclGetUniqueId: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error"));
    return TST_OK;
}

static int set_option(tst_handle* h, const std::string& nm, long v) {
    tst_handle::Knobs& kn = h->knobs;
    if (nm == "L2_PERSIST") h->l2_persist = v != 0;
    else if (nm == "TRACE") kn.trace = v != 0;
    else if (nm == "STASH") kn.stash = v != 0;
    else if (nm == "BIG_PACKED") kn.big_packed = v != 0;
    else if (nm == "ROOT_DENSE") kn.root_dense = (int)v;
    else if (nm == "TINY_MAX") kn.tiny_max = (int)v;
    else if (nm == "MID_MAX") kn.mid_max = (int)v;
    else if (nm == "STASH_MAX") kn.stash_max = (int)v;
    else if (nm == "BLIND_LEVELS") kn.blind_levels = (int)std::max(1l, std::min(8l, v));
    else if (nm == "THIN_NODES") kn.thin_nodes = std::max(0l, v);
    else if (nm == "SUB_ROWS") kn.sub_rows = (int)v;
    else if (nm == "WIDE_MAX") kn.wide_max = (int)v;
    else if (nm == "MID_CLASSES") kn.mid_classes = (int)v;
    else if (nm == "BIG_LITE") kn.big_lite = (int)v;
    else if (nm == "SMALL_MULTI") kn.small_multi = (int)v;
    else return -1;
    return 0;
}

extern "C" int tst_set_option(tst_handle* h, const char* name, int64_t value) {
    if (!h || !name) return fail(TST_ERR_INVALID, "tst_set_option: null argument");
    std::string nm(name);
    for (auto& ch : nm) ch = (char)toupper((unsigned char)ch);
    if (nm.rfind("TST_", 0) == 0) nm = nm.substr(4);
    if (set_option(h, nm, (long)value)) return fail(TST_ERR_INVALID, std::string("tst_set_option: unknown option ") + name);
    return TST_OK;
}

static int create_resources(tst_handle* h, int world, int rank, const void* nccl_unique_id) {
    CK(cudaMallocHost((void**)&h->h_pinned, 256));
    CK(cudaEventCreate(&h->ev0));
    CK(cudaEventCreate(&h->ev1));
    CK(cudaEventCreate(&h->ev2));
    CK(cudaEventCreate(&h->evm0)); CK(cudaEventCreate(&h->evm1)); CK(cudaEventCreate(&h->evs0)); CK(cudaEventCreate(&h->evs1));
    CK(cudaEventCreate(&h->evi0)); CK(cudaEventCreate(&h->evi1)); CK(cudaEventCreate(&h->evk0)); CK(cudaEventCreate(&h->evk1));
    CK(cudaEventCreateWithFlags(&h->ev_bins, cudaEventDisableTiming));
    for (auto& lv : h->ev_lvl) for (auto& e : lv) CK(cudaEventCreate(&e));
    if (nccl_unique_id && world > 1) {
        if (load_nccl()) return fail(TST_ERR_UNSUPPORTED, "tst_create: libnccl.so.2 not found");
        tst_nccl_id id;
        memcpy(&id, nccl_unique_id, sizeof(id));
        int rc = g_nccl.CommInitRank(&h->comm, world, id, rank);
        if (rc) { h->comm = nullptr; return fail(TST_ERR_CUDA, std::string("ncclCommInitRank: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error")); }
    }
    return 0;
}

extern "C" int tst_create(tst_handle** out, int device, int rank, int world, const void* nccl_unique_id) {
    if (!out || world < 1 || rank < 0 || rank >= world) return fail(TST_ERR_INVALID, "tst_create: bad arguments");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(TST_ERR_CUDA, std::string("tst_create: no CUDA device (") + cudaGetErrorString(e) + "); libtst has no CPU fallback");
    }
    if (device < 0 || device >= count) return fail(TST_ERR_INVALID, "tst_create: device out of range");
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(TST_ERR_UNSUPPORTED, "tst_create: libtst is built for sm_100a (B200) only");
    tst_handle* h = new tst_handle();
    h->device = device; h->rank = rank; h->world = world; h->sm_count = prop.multiProcessorCount;
    h->l2_persist_max = (size_t)prop.persistingL2CacheMaxSize; h->l2_window_max = (size_t)prop.accessPolicyMaxWindowSize;
    // tuning / diagnostic knobs: defaults from the environment (TST_<NAME>), read once here; tst_set_option changes them later
    for (const char* nm : {"L2_PERSIST", "TRACE", "STASH", "BIG_PACKED", "ROOT_DENSE", "TINY_MAX", "MID_MAX", "STASH_MAX", "BLIND_LEVELS",
                           "THIN_NODES", "SUB_ROWS", "WIDE_MAX", "MID_CLASSES", "BIG_LITE", "SMALL_MULTI"}) {
        const std::string env = std::string("TST_") + nm;
        if (const char* v = getenv(env.c_str())) set_option(h, nm, atol(v));
    }
    memset(&h->stats, 0, sizeof(h->stats));
    int rc = create_resources(h, world, rank, nccl_unique_id);
    if (rc) { tst_destroy(h); return rc; }  // frees whatever was created before the failure
    *out = h;
    return TST_OK;
}

extern "C" int tst_destroy(tst_handle* h) {
    if (!h) return TST_OK;
    cudaSetDevice(h->device);
    DevBuf* bufs[] = {&h->bins_cm, &h->bins_rm, &h->edges, &h->n_edges, &h->part_count, &h->part_start, &h->sorted, &h->trees,
                      &h->counts8, &h->chunk_cnt, &h->chunk_sum, &h->entA, &h->entB, &h->nodes, &h->leafc, &h->qbig0, &h->qbig1,
                      &h->qmid0, &h->qmid1, &h->qmidc[0][0], &h->qmidc[0][1], &h->qmidc[1][0], &h->qmidc[1][1], &h->qbigl[0], &h->qbigl[1], &h->qwide0, &h->qwide1, &h->qsmall, &h->qsub, &h->imp, &h->imp_sums, &h->target16, &h->qtiny, &h->qtiny2, &h->ghist, &h->rsplit, &h->y_pad, &h->stash, &h->counters, &h->dstats, &h->infer_parts, &h->keep_nodes, &h->keep_leafc, &h->keep_arenas,
                      &h->ibp_fold, &h->ibp_votes, &h->ibp_removed, &h->pred_cm, &h->pred_rm, &h->sort_a, &h->sort_b, &h->sort_dk, &h->gather};
    for (DevBuf* b : bufs) b->release();
    if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
    if (h->h_pinned) cudaFreeHost(h->h_pinned);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->ev2) cudaEventDestroy(h->ev2);
    for (cudaEvent_t e : {h->evm0, h->evm1, h->evs0, h->evs1, h->evi0, h->evi1, h->evk0, h->evk1, h->ev_bins}) if (e) cudaEventDestroy(e);
    // cudaLimitPersistingL2CacheSize is device-wide: hand it back the way the host application (torch) had it
    if (h->l2_carve_touched) { cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, h->l2_carve_saved); cudaGetLastError(); }
    for (auto& lv : h->ev_lvl) for (auto& e : lv) if (e) cudaEventDestroy(e);
    delete h;
    return TST_OK;
}

static int set_kernel_attrs(tst_handle* h) {
    if (h->attrs_set) return 0;
    CK(cudaFuncSetAttribute(k_grow_big<BIG_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BIG_SMEM_BUDGET + 40 * 1024));
    CK(cudaFuncSetAttribute(k_grow_big<BIG_THREADS, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BIG_SMEM_BUDGET + 40 * 1024));
    CK(cudaFuncSetAttribute(k_grow_big<BIG_THREADS, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BIG_SMEM_BUDGET + 40 * 1024));
    CK(cudaFuncSetAttribute(k_grow_small<SMALL_WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    CK(cudaFuncSetAttribute(k_grow_small4<SMALL_WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    CK(cudaFuncSetAttribute(k_grow_small4<SMALL_WARPS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    CK(cudaFuncSetAttribute(k_grow_small4<SMALL_WARPS, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    CK(cudaFuncSetAttribute(k_grow_small4<SMALL_WARPS, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    CK(cudaFuncSetAttribute(k_grow_mid4<MID_WARPS, MID_R, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(k_grow_mid4<MID_WARPS, 2, false, MIDC0_MINB_N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    CK(cudaFuncSetAttribute(k_grow_mid4<MID_WARPS, 2, true, MIDC0_MINB_N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    CK(cudaFuncSetAttribute(k_grow_mid4<MID_WARPS, 4, false, MIDC1_MINB_N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    CK(cudaFuncSetAttribute(k_grow_mid4<MID_WARPS, 4, true, MIDC1_MINB_N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    CK(cudaFuncSetAttribute(k_grow_big<BIG_THREADS, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BIG_SMEM_BUDGET + 40 * 1024));
    CK(cudaFuncSetAttribute(k_grow_mid<MID_WARPS, MID_R>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(k_grow_mid<4, MID_R>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(k_grow_mid<2, MID_R>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(k_grow_mid4<MID_WARPS, MID_R>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(k_grow_sub<16, 2048, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sub_smem_bytes(16, 2048, 64)));
    CK(cudaFuncSetAttribute(k_grow_sub<8, 1024, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sub_smem_bytes(8, 1024, 64)));
    CK(cudaFuncSetAttribute(k_grow_wide<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wide_smem_bytes(4, TST_MAX_FEATURES)));
    CK(cudaFuncSetAttribute(k_root_hist, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(k_root_split, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(k_bin_map, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    CK(cudaFuncSetAttribute(k_cm_to_rm, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    h->attrs_set = true;
    return 0;
}

// --------------------------------------------------------------------------------------------- binning
static int bins_done(tst_handle* h, cudaStream_t st) {  // the producer's last step
    CK(cudaEventRecord(h->ev_bins, st));
    h->bins_stream = st; h->bins_recorded = true;
    return 0;
}
static int wait_bins(tst_handle* h, cudaStream_t st) {  // a consumer's first step (no-op on the producer's own stream)
    if (h->bins_recorded && h->bins_stream != st) CK(cudaStreamWaitEvent(st, h->ev_bins, 0));
    return 0;
}

static int alloc_table(tst_handle* h, int64_t n, int32_t d) {
    if (n < 1 || n >= TST_MAX_ROWS) return fail(TST_ERR_INVALID, "table: need 1 <= n < 2^31 - 16 rows");
    if (d < 1 || d > TST_MAX_FEATURES) return fail(TST_ERR_INVALID, "table: need 1 <= d <= 1024 features");
    h->n = n; h->d = d;
    h->target_n = 0;  // a target belongs to one table
    h->ld = (n + 15) / 16 * 16;
    // rows of <= 128 bytes are padded to a power of two so the cooperative row fetch can shift instead of divide
    h->dstride = (uint32_t)((d + 15) / 16 * 16);
    if (d <= 128) { uint32_t p2 = 16; while (p2 < (uint32_t)d) p2 <<= 1; h->dstride = p2; }
    ENSURE(h->bins_cm, (size_t)h->ld * d);
    ENSURE(h->bins_rm, (size_t)n * h->dstride);
    return 0;
}

extern "C" int tst_bin(tst_handle* h, const float* X_dev, int64_t n, int32_t d, const double* edges_host,
                       const int32_t* n_edges_host, void* stream) {
    if (!h || !X_dev || !edges_host || !n_edges_host) return fail(TST_ERR_INVALID, "tst_bin: null argument");
    CK(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    int rc = alloc_table(h, n, d);
    if (rc) return rc;
    if ((rc = set_kernel_attrs(h))) return rc;
    for (int f = 0; f < d; ++f)
        if (n_edges_host[f] < 0 || n_edges_host[f] > TST_MAX_EDGES) return fail(TST_ERR_INVALID, "tst_bin: n_edges out of range");
    ENSURE(h->edges, (size_t)d * TST_MAX_EDGES * sizeof(double));
    ENSURE(h->n_edges, (size_t)d * sizeof(int32_t));
    CK(cudaMemcpyAsync(h->edges.p, edges_host, (size_t)d * TST_MAX_EDGES * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(h->n_edges.p, n_edges_host, (size_t)d * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(h->bins_rm.p, 0, (size_t)n * h->dstride, st));
    unsigned grid = (unsigned)((n + BIN_TILE - 1) / BIN_TILE);
    size_t smem = (size_t)d * (BIN_TILE + 4);
    k_bin_map<<<grid, 256, smem, st>>>(X_dev, n, d, h->edges.as<double>(), h->n_edges.as<int>(), h->bins_cm.as<uint8_t>(), h->ld,
                                       h->bins_rm.as<uint8_t>(), h->dstride);
    CKL();
    h->stats.kernel_launches++;
    h->have_edges = true;
    return bins_done(h, st);
}

extern "C" int tst_bin_auto(tst_handle* h, const float* X_dev, int64_t n, int32_t d, int32_t n_bins, double* edges_out_host,
                            int32_t* n_edges_out_host, void* stream) {
    if (!h || !X_dev) return fail(TST_ERR_INVALID, "tst_bin_auto: null argument");
    if (n_bins < 3 || n_bins > 256) return fail(TST_ERR_INVALID, "tst_bin_auto: n_bins must be in [3, 256]");
    CK(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    int rc = alloc_table(h, n, d);
    if (rc) return rc;
    if ((rc = set_kernel_attrs(h))) return rc;
    const int64_t ldk = (n + 3) / 4 * 4;
    ENSURE(h->sort_a, (size_t)ldk * d * 4);
    ENSURE(h->sort_b, (size_t)ldk * d * 4);
    ENSURE(h->sort_dk, (size_t)d * 256 * 4);
    ENSURE(h->edges, (size_t)d * TST_MAX_EDGES * sizeof(double));
    ENSURE(h->n_edges, (size_t)d * sizeof(int32_t));
    CK(cudaMemsetAsync(h->edges.p, 0, (size_t)d * TST_MAX_EDGES * sizeof(double), st));
    dim3 gt((unsigned)((n + 31) / 32), (unsigned)((d + 31) / 32));
    k_keys_transpose<<<gt, 256, 0, st>>>(X_dev, n, d, h->sort_a.as<uint32_t>(), ldk);
    CKL();
    k_sort_columns<<<d, SORT_THREADS, 0, st>>>(h->sort_a.as<uint32_t>(), h->sort_b.as<uint32_t>(), n, ldk);
    CKL();
    k_edges_from_sorted<<<d, 256, 0, st>>>(h->sort_a.as<uint32_t>(), n, ldk, n_bins, h->edges.as<double>(), h->n_edges.as<int>(),
                                          h->sort_dk.as<uint32_t>());
    CKL();
    CK(cudaMemsetAsync(h->bins_rm.p, 0, (size_t)n * h->dstride, st));
    unsigned grid = (unsigned)((n + BIN_TILE - 1) / BIN_TILE);
    size_t smem = (size_t)d * (BIN_TILE + 4);
    k_bin_map<<<grid, 256, smem, st>>>(X_dev, n, d, h->edges.as<double>(), h->n_edges.as<int>(), h->bins_cm.as<uint8_t>(), h->ld,
                                       h->bins_rm.as<uint8_t>(), h->dstride);
    CKL();
    h->stats.kernel_launches += 4;
    h->have_edges = true;
    if (int rc2 = bins_done(h, st)) return rc2;
    if (edges_out_host) CK(cudaMemcpyAsync(edges_out_host, h->edges.p, (size_t)d * TST_MAX_EDGES * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (n_edges_out_host) CK(cudaMemcpyAsync(n_edges_out_host, h->n_edges.p, (size_t)d * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    if (edges_out_host || n_edges_out_host) CK(cudaStreamSynchronize(st));
    return TST_OK;
}

extern "C" int tst_get_edges(tst_handle* h, double* edges_out_host, int32_t* n_edges_out_host) {
    if (!h || !edges_out_host || !n_edges_out_host) return fail(TST_ERR_INVALID, "tst_get_edges: null argument");
    if (!h->n || !h->have_edges) return fail(TST_ERR_STATE, "tst_get_edges: the table was not binned by tst_bin / tst_bin_auto");
    CK(cudaSetDevice(h->device));
    if (h->bins_recorded) CK(cudaEventSynchronize(h->ev_bins));
    CK(cudaMemcpy(edges_out_host, h->edges.p, (size_t)h->d * TST_MAX_EDGES * sizeof(double), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(n_edges_out_host, h->n_edges.p, (size_t)h->d * sizeof(int32_t), cudaMemcpyDeviceToHost));
    return TST_OK;
}

extern "C" int tst_set_bins(tst_handle* h, const uint8_t* bins_cm_dev, int64_t n, int32_t d, void* stream) {
    if (!h || !bins_cm_dev) return fail(TST_ERR_INVALID, "tst_set_bins: null argument");
    CK(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    int rc = alloc_table(h, n, d);
    if (rc) return rc;
    if ((rc = set_kernel_attrs(h))) return rc;
    CK(cudaMemsetAsync(h->bins_rm.p, 0, (size_t)n * h->dstride, st));
    unsigned grid = (unsigned)((n + BIN_TILE - 1) / BIN_TILE);
    size_t smem = (size_t)d * (BIN_TILE + 4);
    k_cm_to_rm<<<grid, 256, smem, st>>>(bins_cm_dev, n, d, n, h->bins_cm.as<uint8_t>(), h->ld, h->bins_rm.as<uint8_t>(), h->dstride);
    CKL();
    h->stats.kernel_launches++;
    h->have_edges = false;
    return bins_done(h, st);
}

extern "C" int tst_get_bins(tst_handle* h, uint8_t* out_dev, void* stream) {
    if (!h || !out_dev) return fail(TST_ERR_INVALID, "tst_get_bins: null argument");
    if (!h->n) return fail(TST_ERR_STATE, "tst_get_bins: no table");
    CK(cudaSetDevice(h->device));
    if (int rcw = wait_bins(h, (cudaStream_t)stream)) return rcw;
    int64_t tot = h->n * h->d;
    k_copy_cm<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(h->bins_cm.as<uint8_t>(), h->n, h->d, h->ld, out_dev);
    CKL();
    h->stats.kernel_launches++;
    return TST_OK;
}

extern "C" int tst_set_target(tst_handle* h, const uint16_t* yq_dev, int64_t n, void* stream) {
    if (!h || !yq_dev) return fail(TST_ERR_INVALID, "tst_set_target: null argument");
    if (!h->n || n != h->n) return fail(TST_ERR_STATE, "tst_set_target: bin the table first; the target must have one value per row");
    CK(cudaSetDevice(h->device));
    ENSURE(h->target16, (size_t)n * 2);
    CK(cudaMemcpyAsync(h->target16.p, yq_dev, (size_t)n * 2, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    h->target_n = n;
    return TST_OK;
}

// --------------------------------------------------------------------------------------------- folds
extern "C" int tst_assign_folds(tst_handle* h, const uint8_t* y_dev, const uint8_t* active_dev, const int32_t* group_dev,
                                int32_t n_groups, int32_t k, int32_t C, uint8_t* fold_dev, void* stream) {
    if (!h || !y_dev || !active_dev || !fold_dev) return fail(TST_ERR_INVALID, "tst_assign_folds: null argument");
    if (!h->n) return fail(TST_ERR_STATE, "tst_assign_folds: bin the table first");
    if (k < 1 || k > 250 || C < 1 || C > TST_MAX_CLASSES || n_groups < 1) return fail(TST_ERR_INVALID, "tst_assign_folds: bad k / classes / groups");
    CK(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaMemsetAsync(fold_dev, TST_FOLD_INACTIVE, (size_t)h->n, st));
    k_assign_folds<1024><<<n_groups, 1024, 0, st>>>(y_dev, active_dev, group_dev, h->n, k, C, fold_dev);
    CKL();
    h->stats.kernel_launches++;
    return TST_OK;
}

// --------------------------------------------------------------------------------------------- fit + OOF
struct HostTree { uint32_t part, t; };

// L2 residency: the growth kernels gather rows of the 64-byte-per-row binned table at random while hundreds of
// MB of per-tree entry lists stream through the 126 MB L2.  An access-policy window marks the table persisting
// so the stream cannot evict it (ncu: L2 hit rate 55 % -> see profiles/).
static void set_l2_window(tst_handle* h, cudaStream_t st, bool on) {
    if (!h->l2_persist || !h->l2_persist_max || !h->l2_window_max) return;
    const size_t bytes = (size_t)h->n * h->dstride;
    cudaStreamAttrValue attr;
    memset(&attr, 0, sizeof(attr));
    if (on) {
        const size_t carve = std::min(bytes, h->l2_persist_max);
        if (h->l2_carve_set != carve) {  // device-wide and slow: only when it changes; the previous value is restored by tst_destroy
            if (!h->l2_carve_touched) { cudaDeviceGetLimit(&h->l2_carve_saved, cudaLimitPersistingL2CacheSize); h->l2_carve_touched = true; }
            cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve);
            h->l2_carve_set = carve;
        }
        attr.accessPolicyWindow.base_ptr = h->bins_rm.p;
        attr.accessPolicyWindow.num_bytes = std::min(bytes, h->l2_window_max);
        attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)carve / (double)attr.accessPolicyWindow.num_bytes);
        attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    } else {
        attr.accessPolicyWindow.num_bytes = 0;
    }
    cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr);
    cudaGetLastError();
}
struct L2Window {  // RAII: every return path of tst_fit_oof clears the window
    tst_handle* h; cudaStream_t st;
    L2Window(tst_handle* h_, cudaStream_t st_) : h(h_), st(st_) { set_l2_window(h, st, true); }
    ~L2Window() { set_l2_window(h, st, false); }
};

static int pull_stats(tst_handle* h, cudaStream_t st) {
    DevStats ds;
    CK(cudaMemcpyAsync(&ds, h->dstats.p, sizeof(ds), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    h->stats.row_visits_big += ds.row_visits_big; h->stats.row_visits_mid += ds.row_visits_mid; h->stats.row_visits_small += ds.row_visits_small;
    h->stats.row_visits_sub += ds.row_visits_sub; h->stats.row_visits_wide += ds.row_visits_wide;
#ifdef SUB_PROFILE
    fprintf(stderr, "[sub profile] CTAs %llu: cycles/CTA load %.0f wide %.0f phase2 %.0f | per CTA summed over warps: narrow %.0f small %.0f idle %.0f\n", ds.prof[6],
            (double)ds.prof[0] / std::max(1ull, ds.prof[6]), (double)ds.prof[1] / std::max(1ull, ds.prof[6]), (double)ds.prof[2] / std::max(1ull, ds.prof[6]),
            (double)ds.prof[3] / std::max(1ull, ds.prof[6]), (double)ds.prof[4] / std::max(1ull, ds.prof[6]), (double)ds.prof[5] / std::max(1ull, ds.prof[6]));
#endif
    h->stats.nodes += ds.nodes; h->stats.leaves += ds.leaves; h->stats.searched += ds.searched; h->stats.fallback += ds.fallback;
    h->stats.inbag_rows += ds.inbag_rows; h->stats.oof_steps += ds.oof_steps;
    h->stats.max_depth = std::max(h->stats.max_depth, ds.max_depth);
    CK(cudaMemsetAsync(h->dstats.p, 0, sizeof(DevStats), st));
    if (ds.error == 1) return fail(TST_ERR_OVERFLOW, "bootstrap multiplicity exceeded 255 (byte counter wrapped)");
    if (ds.error == 2) return fail(TST_ERR_OVERFLOW, "node work queue overflowed");
    if (ds.error == 4) return fail(TST_ERR_STATE, "internal: a chosen split left one side empty");
    if (ds.error) return fail(TST_ERR_OVERFLOW, "device error flag " + std::to_string(ds.error));
    return 0;
}

extern "C" int tst_fit_oof(tst_handle* h, const uint8_t* y_dev, const uint8_t* fold_dev, const int32_t* group_dev,
                           int32_t n_groups, int32_t k, int32_t n_trees, int32_t C, const tst_params* prm, uint64_t seed,
                           uint32_t gid_base, int32_t t0, int32_t t1, uint32_t* votes_dev, void* stream) {
    if (!h || !y_dev || !fold_dev || !prm) return fail(TST_ERR_INVALID, "tst_fit_oof: null argument");
    if (!h->n) return fail(TST_ERR_STATE, "tst_fit_oof: bin the table first (tst_bin / tst_set_bins)");
    const int64_t n = h->n;
    const int32_t d = h->d;
    if (k < 1 || k > 250 || n_groups < 1 || n_trees < 1 || n_trees > 65535 || C < 1 || C > TST_MAX_CLASSES || t0 < 0 || t1 > n_trees || t0 > t1)
        return fail(TST_ERR_INVALID, "tst_fit_oof: bad k / n_trees / classes / tree range");
    if (prm->max_features < 1 || prm->max_features > d || prm->min_samples_leaf < 1 || prm->min_samples_split < 2)
        return fail(TST_ERR_INVALID, "tst_fit_oof: bad growth parameters");
    if (prm->criterion != TST_GINI && prm->criterion != TST_SQUARED_ERROR_LEVELS && prm->criterion != TST_SQUARED_ERROR)
        return fail(TST_ERR_INVALID, "tst_fit_oof: unknown criterion");
    const bool reg = prm->criterion == TST_SQUARED_ERROR;
    if (reg && C != 2) return fail(TST_ERR_INVALID, "tst_fit_oof: the squared-error criterion takes n_classes = 2 (votes are the two 16-bit halves of the Q16.16 leaf means)");
    if (reg && h->target_n != n) return fail(TST_ERR_STATE, "tst_fit_oof: the squared-error criterion needs the 16-bit target of this table (tst_set_target)");
    if (reg && n >= (1ll << 24) / 255) return fail(TST_ERR_UNSUPPORTED, "tst_fit_oof: squared-error criterion: a node must weigh < 2^24");
    if ((uint64_t)n_groups * k * n_trees + gid_base > 0xFFFFFFFFull) return fail(TST_ERR_INVALID, "tst_fit_oof: tree id overflows 32 bits");
    CK(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    const tst_handle::Knobs& kn = h->knobs;
    const bool trace = kn.trace;  // host wall-clock phases of every batch on stderr
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms_since = [](std::chrono::steady_clock::time_point a) { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - a).count(); };
    const auto t_entry = now();
    if (trace) cudaStreamSynchronize(st);  // work queued by the caller before this call is not ours
    const double tr_wait = ms_since(t_entry);
    int rc = set_kernel_attrs(h);
    if (rc) return rc;
    const double tr_attrs = ms_since(t_entry);

    // ---- partitions: rows sorted by (group, fold slot, row)
    const int np = n_groups * (k + 1);
    ENSURE(h->part_count, (size_t)np * 4);
    ENSURE(h->part_start, (size_t)(np + 1) * 4);
    ENSURE(h->sorted, (size_t)n * 4);
    ENSURE(h->dstats, sizeof(DevStats));
    ENSURE(h->counters, 256);
    CK(cudaMemsetAsync(h->dstats.p, 0, sizeof(DevStats), st));
    k_part_count<1024><<<np, 1024, 0, st>>>(fold_dev, group_dev, n, k, h->part_count.as<uint32_t>());
    CKL();
    k_part_scan<<<1, 32, 0, st>>>(h->part_count.as<uint32_t>(), np, h->part_start.as<uint32_t>());
    CKL();
    k_part_scatter<1024><<<np, 1024, 0, st>>>(fold_dev, group_dev, n, k, h->part_start.as<uint32_t>(), h->sorted.as<uint32_t>());
    CKL();
    h->stats.kernel_launches += 3;
    std::vector<uint32_t> ps(np + 1);
    CK(cudaMemcpyAsync(ps.data(), h->part_start.p, (size_t)(np + 1) * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));

    // ---- this rank's trees, ordered by partition
    std::vector<TreeDesc> all;
    std::vector<HostTree> all_id;
    uint32_t max_train = 0;
    for (int part = 0; part < n_groups * k; ++part) {
        int g = part / k, f = part % k;
        uint32_t gs = ps[(size_t)g * (k + 1)], ge = ps[(size_t)(g + 1) * (k + 1)];
        uint32_t fs = ps[(size_t)g * (k + 1) + f], fe = ps[(size_t)g * (k + 1) + f + 1];
        for (int t = t0; t < t1; ++t) {
            uint32_t lid = (uint32_t)part * (uint32_t)n_trees + (uint32_t)t;
            if ((int)(lid % (uint32_t)h->world) != h->rank) continue;
            TreeDesc td;
            memset(&td, 0, sizeof(td));
            td.gid = gid_base + lid; td.part = (uint32_t)part; td.gs = gs; td.a = fs - gs; td.fsz = fe - fs;
            td.n_train = (ge - gs) - (fe - fs); td.fs = fs;
            max_train = std::max(max_train, td.n_train);
            all.push_back(td);
            all_id.push_back({(uint32_t)part, (uint32_t)t});
        }
    }
    const size_t total_trees = all.size();
    h->kept.clear();
    h->imp_trees = 0; h->imp_d = d;
    if (total_trees == 0) return TST_OK;
    const double tr_parts = ms_since(t_entry);
    L2Window l2w(h, st);
    const double tr_l2 = ms_since(t_entry);

    // ---- sizes (everything for B trees in flight stays resident in HBM)
    const uint64_t n_cap = std::max<uint64_t>(4, ((uint64_t)max_train + 3) / 4 * 4);
    const uint64_t n_pad = (uint64_t)h->ld;
    const uint32_t n_chunks = (uint32_t)((n_pad + BOOT_CHUNK - 1) / BOOT_CHUNK);
    const uint64_t node_stride = 2 * n_cap;
    const uint64_t leafc_stride = (n_cap / 2 + 1) * (uint64_t)C;
    // queue slots per tree.  With the thread-per-subtree tier on, qsmall only sees subtrees of >= 7 rows (<= n/7 of them); with it
    // off (many classes, squared error, d > 64) every 2..32-row subtree lands there: up to n/2 tasks for a tree that peels rows off.
    const uint64_t tiny_eff = (((C <= 4 && prm->criterion == TST_GINI) || reg) && d <= TINY_FY_MAX) ? (uint64_t)(kn.tiny_max >= 0 ? std::min(kn.tiny_max, TINY_MAX) : TINY_MAX) : 0;
    const uint64_t wide_lo = (uint64_t)std::max(32, kn.mid_max >= 0 ? std::min(kn.mid_max, 32 * MID_R) : 32 * MID_R);
    const uint64_t qwide_per = n_cap / (wide_lo + 1) + 2;  // disjoint segments of more rows than the warp-per-node tier takes
    const uint64_t qbig_per = n_cap / 32 + 2, qsmall_per = n_cap / std::max<uint64_t>(2, tiny_eff + 1) + 2, qtiny_per = n_cap / 2 + 2;
    const size_t arena_bytes = node_stride * sizeof(uint2) + leafc_stride * 4;
    const bool want_stash = kn.stash;  // see use_stash below
    const size_t work_bytes = n_pad + 2 * (size_t)n_chunks * 4 + (want_stash ? 3 : 2) * n_cap * sizeof(Entry) + (11 * qbig_per + 2 * qwide_per + qsmall_per) * sizeof(Task) + 2 * qtiny_per * 16 + sizeof(TreeDesc);
    const bool keep = prm->keep_forest != 0;
    size_t reclaim = h->counts8.cap + h->chunk_cnt.cap + h->chunk_sum.cap + h->entA.cap + h->entB.cap + h->nodes.cap + h->leafc.cap +
                     h->qbig0.cap + h->qbig1.cap + h->qmid0.cap + h->qmid1.cap + h->qmidc[0][0].cap + h->qmidc[0][1].cap + h->qmidc[1][0].cap + h->qmidc[1][1].cap + h->qbigl[0].cap + h->qbigl[1].cap + h->qwide0.cap + h->qwide1.cap + h->qsmall.cap + h->qsub.cap + h->qtiny.cap + h->qtiny2.cap + h->keep_nodes.cap + h->keep_leafc.cap + h->stash.cap;
    const size_t per_tree = work_bytes + (keep ? 0 : arena_bytes);
    size_t B = std::min<size_t>(total_trees, 4096);
    if (prm->trees_in_flight > 0) B = std::min<size_t>(B, (size_t)prm->trees_in_flight);
    // cudaMemGetInfo costs 0.3 - 0.9 ms once tens of GB are allocated (a whole IBP round on a 5k-row table is 4 ms): only ask the
    // driver when this job's buffers do not already fit into what the handle holds from earlier calls
    if (B * per_tree + (keep ? total_trees * arena_bytes : 0) > reclaim) {
        size_t free_b = 0, total_b = 0;
        CK(cudaMemGetInfo(&free_b, &total_b));
        size_t budget = (size_t)((free_b + reclaim) * 0.6);
        if (keep) {
            size_t need = total_trees * arena_bytes;
            if (need > budget) return fail(TST_ERR_NOMEM, "keep_forest: node arrays of " + std::to_string(total_trees) + " trees (" + std::to_string(need >> 20) + " MiB) do not fit");
            budget -= need;
        }
        B = std::min<size_t>(B, std::max<size_t>(1, budget / per_tree));
    }

    ENSURE(h->trees, B * sizeof(TreeDesc));
    ENSURE(h->counts8, B * n_pad);
    ENSURE(h->chunk_cnt, B * (size_t)n_chunks * 4);
    ENSURE(h->chunk_sum, B * (size_t)n_chunks * 4);
    ENSURE(h->entA, B * n_cap * sizeof(Entry) + 4096);  // slack: TMA tiles are whole 256-byte chunks
    ENSURE(h->entB, B * n_cap * sizeof(Entry) + 4096);
    ENSURE(h->qbig0, B * qbig_per * sizeof(Task));
    ENSURE(h->qbig1, B * qbig_per * sizeof(Task));
    ENSURE(h->qmid0, B * qbig_per * sizeof(Task));
    ENSURE(h->qmid1, B * qbig_per * sizeof(Task));
    for (int c = 0; c < 2; ++c) for (int p = 0; p < 2; ++p) ENSURE(h->qmidc[c][p], B * qbig_per * sizeof(Task));
    for (int p = 0; p < 2; ++p) ENSURE(h->qbigl[p], B * qbig_per * sizeof(Task));
    ENSURE(h->qsmall, B * qsmall_per * sizeof(Task));
    ENSURE(h->qsub, B * qbig_per * sizeof(Task));
    ENSURE(h->qwide0, B * qwide_per * sizeof(Task));
    ENSURE(h->qwide1, B * qwide_per * sizeof(Task));
    // every warp of k_grow_small4 may leave up to TINY_RESERVE - 1 reserved slots unused
    const uint64_t qtiny_slack = (uint64_t)h->sm_count * 64u * SMALL_WARPS * TINY_RESERVE;
    ENSURE(h->qtiny, (B * qtiny_per + qtiny_slack) * 16);
    ENSURE(h->qtiny2, (B * qtiny_per + qtiny_slack) * 16);
    ENSURE(h->infer_parts, (size_t)n_groups * k * sizeof(InferPart));
    uint2* nodes_p; uint32_t* leafc_p;
    if (keep) {
        ENSURE(h->keep_nodes, total_trees * node_stride * sizeof(uint2));
        ENSURE(h->keep_leafc, total_trees * leafc_stride * 4);
        nodes_p = h->keep_nodes.as<uint2>(); leafc_p = h->keep_leafc.as<uint32_t>();
        h->kept_C = C; h->kept_criterion = prm->criterion; h->kept_n_trees = n_trees; h->kept_node_stride = node_stride; h->kept_leafc_stride = leafc_stride;
    } else {
        ENSURE(h->nodes, B * node_stride * sizeof(uint2));
        ENSURE(h->leafc, B * leafc_stride * 4);
        nodes_p = h->nodes.as<uint2>(); leafc_p = h->leafc.as<uint32_t>();
    }

    ENSURE(h->imp, total_trees * (size_t)d * 8);
    CK(cudaMemsetAsync(h->imp.p, 0, total_trees * (size_t)d * 8, st));
    GrowCtx g;
    memset(&g, 0, sizeof(g));
    g.imp = h->imp.as<unsigned long long>();
    g.bins_rm = h->bins_rm.as<uint8_t>(); g.dstride = h->dstride; g.d = d; g.C = C; g.m = prm->max_features;
    g.max_depth = prm->max_depth; g.mss = (uint32_t)prm->min_samples_split; g.msl = (uint32_t)prm->min_samples_leaf;
    g.criterion = (uint32_t)prm->criterion;
    g.k0 = (uint32_t)seed; g.k1 = (uint32_t)(seed >> 32);
    g.entA = h->entA.as<Entry>(); g.entB = h->entB.as<Entry>(); g.n_cap = n_cap;
    g.nodes = nodes_p; g.node_stride = node_stride; g.leafc = leafc_p; g.leafc_stride = leafc_stride;
    g.trees = h->trees.as<TreeDesc>();
    g.qsmall = h->qsmall.as<Task>();
    g.stats = h->dstats.as<DevStats>();
    g.FS = 32 * (8 * C + 1);
    g.MB = (int)std::min<size_t>(std::min<size_t>(32, (size_t)g.m), std::max<size_t>(1, BIG_SMEM_BUDGET / ((size_t)(reg ? BIGP_FS : g.FS) * 4)));
    g.big_tile = h->dstride <= 64 ? 1 : 0;  // cooperative whole-row fetch into a shared tile (rows of <= 64 bytes)
    const size_t big_tile_bytes = g.big_tile ? (size_t)BIG_THREADS * ((h->dstride >> 2) + 1) * 4 : 0;
    g.big_packed = kn.big_packed ? 1u : 0u;
    // Option (TST_STASH=1, off by default): the histogram pass of k_grow_big leaves the m sampled bins of every row in a
    // stash (8 bytes per row, coalesced) so that its partition pass does not gather the split feature's byte from the
    // table again.  Measured on B200 (1M x 64, 640 trees): 25 % fewer LSU wavefronts and -5 % kernel time on the full
    // levels, but the stash streams through HBM (+5.8 GB per level on top of 9.3 GB: 1.8x the algorithmic bytes instead
    // of 1.1x) for 0.9 % of the step -- not worth the traffic.
    const bool use_stash = want_stash && g.big_tile && g.m <= 8 && g.m <= g.MB;
    if (use_stash) {
        ENSURE(h->stash, B * n_cap * sizeof(uint2));
        g.stash = h->stash.as<uint2>();
        g.stash_max_rows = 0xFFFFFFFFu;
        if (kn.stash_max >= 0) g.stash_max_rows = (uint32_t)kn.stash_max;
    }
    // the packed kernels: <= 4 classes with the Gini criterion, or the continuous squared-error criterion (64-bit (weight, sum) cells)
    const bool packed_ok = (C <= 4 && prm->criterion == TST_GINI) || reg;
    // the packed 16-bit layout needs BIGP_FS = 640 words per feature, more than FS when C == 2 (544): size for the larger one
    g.HFS_max = (reg || (packed_ok && g.big_packed && g.big_tile && g.MB <= 8)) ? std::max<int>(g.FS, (int)BIGP_FS) : g.FS;
    const size_t big_smem = (size_t)g.MB * g.HFS_max * 4 + big_tile_bytes;
    if (big_smem > BIG_SMEM_BUDGET + big_tile_bytes) return fail(TST_ERR_UNSUPPORTED, "histogram for one feature exceeds the shared-memory budget");
    const bool small_packed = packed_ok;  // register-only warp kernel (tst_grow_small.cuh); generic one otherwise
    const uint32_t small_per_warp = small_packed ? small4_per_warp(d)
                                                 : (uint32_t)C * 32u + (uint32_t)C + 32u + 96u + (uint32_t)((d + 1) / 2);
    const size_t small_smem = (size_t)SMALL_WARPS * small_per_warp * 4;
    // [2p], [2p+1] big / mid queue counts of parity p, [4] small, [5] tiny, [6 + r] tiny size histogram (r <= 6), [16 + p] four-warps-per-node
    // queue of parity p, [22 + 2c + p] size class c of the warp-per-node tier, [26 + p] light nodes of the CTA tier, [20] subtree queue, [32 + r] tiny cursors
    uint32_t* counters = h->counters.as<uint32_t>();
    g.qsmall_count = counters + 4;
    g.qtiny = h->qtiny.p; g.qtiny_count = counters + 5;
    g.tiny_max = (packed_ok && d <= TINY_FY_MAX && B <= 65535) ? TINY_MAX : 0;  // packed 4-class thread-per-subtree tier
    if (kn.tiny_max >= 0) g.tiny_max = std::min<uint32_t>(g.tiny_max, (uint32_t)kn.tiny_max);
    const bool mid_packed = packed_ok;  // packed 16-bit class counts (k_grow_mid4)
    const uint32_t mid_per_warp = mid_packed ? mid4_per_warp(MID_R, d)
                                             : (uint32_t)g.FS + 2u * (uint32_t)C + (uint32_t)((d + 1) / 2);
    // warps per CTA of the warp-per-node tier: 8, or fewer when the unpacked histograms of many classes would not fit
    int mid_warps = MID_WARPS;
    if (!mid_packed) while (mid_warps > 2 && (size_t)mid_warps * mid_per_warp * 4 > 200 * 1024) mid_warps >>= 1;
    const size_t mid_smem = (size_t)mid_warps * mid_per_warp * 4;
    g.mid_max = mid_smem <= 200 * 1024 ? 32u * MID_R : SMALL_MAX;
    g.mid_packed = mid_packed ? 1u : 0u;
    if (kn.mid_max >= 0) g.mid_max = std::max<uint32_t>(SMALL_MAX, std::min<uint32_t>(32u * MID_R, (uint32_t)kn.mid_max));
    // Two more launches per level: they pay once the levels are many waves deep (measured crossover, tools/class_crossover.py:
    // ~10M in-bag rows per batch; 5 x 100 trees on 5k rows lose 0.5 ms of 3.3 to the extra launches), so "auto" looks at the batch.
    const bool deep_levels = (uint64_t)B * n_cap >= (1ull << 24);
    const int mid_classes = (mid_packed && g.mid_max > SMALL_MAX) ? (kn.mid_classes >= 0 ? kn.mid_classes & 3 : (deep_levels ? 3 : 0)) : 0;
    // light nodes of the CTA tier (weight < 2^16: levels 5+ of a 1M-row tree) run the packed-Gini-only instantiation, 4 CTAs per SM
    g.bigl = (!reg && packed_ok && g.big_packed && g.big_tile && (kn.big_lite >= 0 ? kn.big_lite != 0 : deep_levels)) ? 1u : 0u;
    const bool small_multi = kn.small_multi >= 0 ? kn.small_multi != 0 : (deep_levels || reg);
    const size_t bigl_smem = (8 * (size_t)BIGP_FS + (size_t)BIG_THREADS * ((h->dstride >> 2) + 1)) * 4;
    size_t midc_smem[2];
    for (int c = 0; c < 2; ++c) {
        g.midc_max[c] = (mid_classes >> c & 1) ? 32u * MIDC_R[c] : 0u;
        midc_smem[c] = (size_t)MID_WARPS * mid4_per_warp(MIDC_R[c], d) * 4;
    }
    // CTA-per-subtree tier (tst_grow_sub.cuh): packed 4-class Gini, rows of <= 64 bytes.  It takes every node of 33 .. sub_max
    // rows that weighs < 2^16 and finishes the whole subtree, so the warp-per-node tier is off and the warp-per-subtree kernel only
    // sees the rare <= 32-row children of CTA-tier nodes.  SUB_ROWS: 2048 (16 warps, 1 CTA/SM, default), 1024 (8 warps, 2 CTAs/SM), 0 = off.
    // Measured on B200 (profiles/r02_k_grow_sub.md): bit-exact but 3x slower than the level-synchronous tiers -- the row cache leaves
    // 16 warps per SM for code that is a chain of dependent steps -- so it is OFF by default (SUB_ROWS=2048 / 1024 turns it on).
    int sub_rows = 0;
    if (kn.sub_rows > 0 && packed_ok && !reg && h->dstride <= 64 && B <= 65535) sub_rows = kn.sub_rows >= 2048 ? 2048 : (kn.sub_rows >= 1024 ? 1024 : 0);
    g.sub_max = (uint32_t)sub_rows;
    g.qsub = h->qsub.as<Task>(); g.qsub_count = counters + 20;
    if (sub_rows) g.mid_max = SMALL_MAX;
    // four-warps-per-node tier (tst_grow_wide.cuh): the packed Gini nodes of mid_max + 1 .. 2048 rows that k_grow_big would take
    g.wide_max = (packed_ok && !reg && !sub_rows && g.mid_max > SMALL_MAX) ? WIDE_MAXR : 0u;
    if (kn.wide_max >= 0 && g.wide_max) g.wide_max = (uint32_t)std::min<int>(kn.wide_max, (int)WIDE_MAXR) > g.mid_max ? (uint32_t)std::min<int>(kn.wide_max, (int)WIDE_MAXR) : 0u;
    const size_t wide_smem = wide_smem_bytes(4, d);
    g.big_min = sub_rows ? (uint32_t)sub_rows : (g.wide_max ? g.wide_max : g.mid_max);
    const size_t sub_smem = sub_rows ? sub_smem_bytes(sub_rows == 2048 ? 16 : 8, sub_rows, h->dstride) : 0;
    // dense root histograms (tst_root.cuh): big tables only; needs m <= 8 and the histogram + 2 TMA stages in shared memory
    const size_t root_smem = (size_t)g.m * g.FS * 4 + 2 * (size_t)(g.m + 2) * ROOT_TILE;
    bool root_dense = n >= 65536 && g.m <= 8 && root_smem <= 200 * 1024 && max_train > g.big_min;
    if (kn.root_dense == 0) root_dense = false;  // 0 = off, 2 = on for any table size (tests)
    if (kn.root_dense == 2) root_dense = g.m <= 8 && root_smem <= 200 * 1024 && max_train > g.big_min;
    if (reg) root_dense = false;  // the dense root kernels histogram class counts
    if (root_dense) {
        ENSURE(h->ghist, B * (size_t)g.m * g.FS * 4);
        ENSURE(h->rsplit, B * sizeof(RootSplit));
        ENSURE(h->y_pad, (size_t)h->ld);
        CK(cudaMemsetAsync(h->y_pad.p, 0, (size_t)h->ld, st));
        CK(cudaMemcpyAsync(h->y_pad.p, y_dev, (size_t)n, cudaMemcpyDeviceToDevice, st));
    }

    if (trace) {
        cudaStreamSynchronize(st);
        fprintf(stderr, "[tst trace] setup %.3f ms: caller's queued work %.3f, kernel attributes %.3f, partitions %.3f, L2 window %.3f, buffers %.3f\n", ms_since(t_entry),
                tr_wait, tr_attrs - tr_wait, tr_parts - tr_attrs, tr_l2 - tr_parts, ms_since(t_entry) - tr_l2);
    }
    NvtxRange nvtx_fit("tst_fit_oof");
    for (size_t b0 = 0; b0 < total_trees; b0 += B) {
        NvtxRange nvtx_batch("tst batch of trees");
        const auto t_batch = now();
        double tr_boot = 0, tr_levels = 0, tr_tail = 0;
        const uint32_t nb = (uint32_t)std::min(B, total_trees - b0);
        for (uint32_t i = 0; i < nb; ++i) { all[b0 + i].arena = keep ? (uint32_t)(b0 + i) : i; all[b0 + i].imp_row = (uint32_t)(b0 + i); }
        CK(cudaMemcpyAsync(h->trees.p, all.data() + b0, (size_t)nb * sizeof(TreeDesc), cudaMemcpyHostToDevice, st));
        g.qbig_cap = (uint32_t)std::min<uint64_t>(0xFFFFFFFFull, nb * qbig_per);
        g.qwide_cap = (uint32_t)std::min<uint64_t>(0xFFFFFFFFull, nb * qwide_per);
        g.qsmall_cap = (uint32_t)std::min<uint64_t>(0xFFFFFFFFull, nb * qsmall_per);
        g.qtiny_cap = (uint32_t)std::min<uint64_t>(0xFFFFFFFFull, nb * qtiny_per + qtiny_slack);

        // ---- bootstrap + compaction
        NvtxPhase phase;
        phase.next("tst bootstrap + roots");
        CK(cudaEventRecord(h->ev0, st));
        CK(cudaMemsetAsync(h->counts8.p, 0, (size_t)nb * n_pad, st));
        CK(cudaMemsetAsync(counters, 0, 256, st));
        {
            uint32_t blocks4 = (max_train + 3) / 4;
            dim3 grid((blocks4 + 255) / 256, nb);
            if (grid.x) k_boot_draw<<<grid, 256, 0, st>>>(g.trees, h->sorted.as<uint32_t>(), h->counts8.as<uint32_t>(), n_pad, g.k0, g.k1);
            CKL();
            dim3 gridc(n_chunks, nb);
            k_boot_count<<<gridc, 256, 0, st>>>(h->counts8.as<uint32_t>(), n_pad, n_chunks, h->chunk_cnt.as<uint32_t>(), h->chunk_sum.as<uint32_t>());
            CKL();
            k_boot_scan<<<nb, 256, 0, st>>>(g.trees, n_chunks, h->chunk_cnt.as<uint32_t>(), h->chunk_sum.as<uint32_t>(), g.stats);
            CKL();
            k_boot_scatter<<<gridc, 256, 0, st>>>(h->counts8.as<uint32_t>(), n_pad, n_chunks, h->chunk_cnt.as<uint32_t>(), y_dev, reg ? h->target16.as<uint16_t>() : nullptr, g.entA, n_cap);
            CKL();
            g.rsplit = nullptr;
            // everything above only needed labels and folds; from here on the kernels read the binned table
            if ((rc = wait_bins(h, st))) return rc;
            if (root_dense) {
                CK(cudaMemsetAsync(h->ghist.p, 0, (size_t)nb * g.m * g.FS * 4, st));
                dim3 gr((unsigned)((n + ROOT_CHUNK - 1) / ROOT_CHUNK), nb);
                k_root_hist<<<gr, 256, root_smem, st>>>(g, h->bins_cm.as<uint8_t>(), (uint64_t)h->ld, h->y_pad.as<uint8_t>(), h->counts8.as<uint8_t>(),
                                                       (uint64_t)n, h->ghist.as<uint32_t>());
                CKL();
                k_root_split<<<nb, 256, (size_t)g.m * g.FS * 4, st>>>(g, h->ghist.as<uint32_t>(), h->rsplit.as<RootSplit>());
                CKL();
                g.rsplit = h->rsplit.as<RootSplit>();
                h->stats.kernel_launches += 2; h->stats.launches_root += 2;
            }
            g.qnext = h->qbig0.as<Task>(); g.qnext_count = counters + 0;
            g.qmid_next = h->qmid0.as<Task>(); g.qmid_next_count = counters + 1;
            g.qwide_next = h->qwide0.as<Task>(); g.qwide_next_count = counters + 16;
            for (int c = 0; c < 2; ++c) { g.qmidc_next[c] = h->qmidc[c][0].as<Task>(); g.qmidc_next_count[c] = counters + 22 + 2 * c; }
            g.qbigl_next = h->qbigl[0].as<Task>(); g.qbigl_next_count = counters + 26;
            k_init_roots<<<(nb + 127) / 128, 128, 0, st>>>(g, nb);
            CKL();
            h->stats.kernel_launches += 5;
        }
        CK(cudaEventRecord(h->ev1, st));
        CK(cudaMemcpyAsync(h->h_pinned, counters, 128, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        { float ms = 0; cudaEventElapsedTime(&ms, h->ev0, h->ev1); h->stats.ms_prep += ms; }
        tr_boot = ms_since(t_batch);
        phase.next("tst level-synchronous growth");

        // ---- level-synchronous growth: CTA per node (> mid_max rows) and warp per node (33 .. mid_max rows).
        // One host round trip per level sizes the grids exactly.  Once a level is thin (the long tail of an unbalanced
        // tree: a few nodes per level for dozens of levels) up to 4 levels are launched back to back with grids that are
        // upper bounds (children <= 2 x parents) and the kernels read the real counts on the device.
        int cur = 0;
        uint64_t level_guard = 0;
        bool small_started = false, tiny_timed = false, sub_timed = false;
        const int blind_levels = kn.blind_levels;
        const uint64_t thin_nodes = (uint64_t)kn.thin_nodes;  // a level with at most this many nodes starts a group of blind levels
        uint32_t nbig = h->h_pinned[0], nmid = h->h_pinned[1], nwide = h->h_pinned[16], nmidc[2] = {h->h_pinned[22], h->h_pinned[24]}, nbigl = h->h_pinned[26];
        double tr_gap = 0; int tr_groups = 0;
        while (nbig > 0 || nmid > 0 || nwide > 0 || nmidc[0] > 0 || nmidc[1] > 0 || nbigl > 0) {
            const auto t_level = now();
            const bool thin = blind_levels > 1 && (uint64_t)nbig + nmid + nwide + nmidc[0] + nmidc[1] + nbigl <= thin_nodes;
            const int J = thin ? blind_levels : 1;
            uint64_t ub_big = nbig, ub_mid = nmid, ub_wide = nwide, ub_midc[2] = {nmidc[0], nmidc[1]}, ub_bigl = nbigl;
            for (int j = 0; j < J; ++j) {
                Task* qcur = cur ? h->qbig1.as<Task>() : h->qbig0.as<Task>();
                Task* qmcur = cur ? h->qmid1.as<Task>() : h->qmid0.as<Task>();
                g.qnext = cur ? h->qbig0.as<Task>() : h->qbig1.as<Task>();
                g.qnext_count = counters + 2 * (cur ^ 1);
                g.qmid_next = cur ? h->qmid0.as<Task>() : h->qmid1.as<Task>();
                g.qmid_next_count = counters + 2 * (cur ^ 1) + 1;
                Task* qwcur = cur ? h->qwide1.as<Task>() : h->qwide0.as<Task>();
                g.qwide_next = cur ? h->qwide0.as<Task>() : h->qwide1.as<Task>();
                g.qwide_next_count = counters + 16 + (cur ^ 1);
                CK(cudaMemsetAsync(g.qnext_count, 0, 8, st));  // both counters of the next level
                Task* qmccur[2]; uint32_t gmc[2]; const uint32_t* cmc[2];
                for (int c = 0; c < 2; ++c) {
                    qmccur[c] = h->qmidc[c][cur].as<Task>();
                    g.qmidc_next[c] = h->qmidc[c][cur ^ 1].as<Task>();
                    g.qmidc_next_count[c] = counters + 22 + 2 * c + (cur ^ 1);
                    if (g.midc_max[c]) CK(cudaMemsetAsync(g.qmidc_next_count[c], 0, 4, st));
                    gmc[c] = g.midc_max[c] ? (uint32_t)std::min<uint64_t>(ub_midc[c], g.qbig_cap) : 0u;
                    cmc[c] = j ? counters + 22 + 2 * c + cur : nullptr;
                }
                if (g.wide_max) CK(cudaMemsetAsync(g.qwide_next_count, 0, 4, st));
                Task* qblcur = h->qbigl[cur].as<Task>();
                g.qbigl_next = h->qbigl[cur ^ 1].as<Task>();
                g.qbigl_next_count = counters + 26 + (cur ^ 1);
                if (g.bigl) CK(cudaMemsetAsync(g.qbigl_next_count, 0, 4, st));
                const uint32_t gbl = g.bigl ? (uint32_t)std::min<uint64_t>(ub_bigl, g.qbig_cap) : 0u;
                const uint32_t* cbl = j ? counters + 26 + cur : nullptr;
                const uint32_t gw = g.wide_max ? (uint32_t)std::min<uint64_t>(ub_wide, g.qwide_cap) : 0u;
                const uint32_t* cw = j ? counters + 16 + cur : nullptr;
                const uint32_t gb = (uint32_t)std::min<uint64_t>(ub_big, g.qbig_cap);
                const uint32_t gm = g.mid_max > SMALL_MAX ? (uint32_t)std::min<uint64_t>(ub_mid, g.qbig_cap) : 0u;  // tier off: its histogram does not fit
                const uint32_t* cb = j ? counters + 2 * cur : nullptr;  // j == 0: counts are exact, no device read needed
                const uint32_t* cm = j ? counters + 2 * cur + 1 : nullptr;
                CK(cudaEventRecord(h->ev_lvl[j][0], st));
                if (gb) {
                    if (g.criterion == 1u) k_grow_big<BIG_THREADS, 1><<<gb, BIG_THREADS, big_smem, st>>>(g, qcur, gb, cb);
                    else if (reg) k_grow_big<BIG_THREADS, 2><<<gb, BIG_THREADS, big_smem, st>>>(g, qcur, gb, cb);
                    else k_grow_big<BIG_THREADS><<<gb, BIG_THREADS, big_smem, st>>>(g, qcur, gb, cb);
                    CKL();
                    h->stats.kernel_launches++; h->stats.launches_grow_big++;
                }
                if (gbl) {
                    k_grow_big<BIG_THREADS, 0, true><<<gbl, BIG_THREADS, bigl_smem, st>>>(g, qblcur, gbl, cbl);
                    CKL();
                    h->stats.kernel_launches++; h->stats.launches_grow_big++;
                }
                CK(cudaEventRecord(h->ev_lvl[j][1], st));
                if (gw) {
                    k_grow_wide<4><<<gw, 128, wide_smem, st>>>(g, qwcur, gw, cw);
                    CKL();
                    h->stats.kernel_launches++; h->stats.launches_grow_wide++;
                }
                CK(cudaEventRecord(h->ev_lvl[j][3], st));
                if (gm) {
                    if (reg) k_grow_mid4<MID_WARPS, MID_R, true><<<(gm + MID_WARPS - 1) / MID_WARPS, MID_WARPS * 32, mid_smem, st>>>(g, qmcur, gm, cm);
                    else if (mid_packed) k_grow_mid4<MID_WARPS, MID_R><<<(gm + MID_WARPS - 1) / MID_WARPS, MID_WARPS * 32, mid_smem, st>>>(g, qmcur, gm, cm);
                    else if (mid_warps == 8) k_grow_mid<8, MID_R><<<(gm + 7) / 8, 256, mid_smem, st>>>(g, qmcur, gm, cm);
                    else if (mid_warps == 4) k_grow_mid<4, MID_R><<<(gm + 3) / 4, 128, mid_smem, st>>>(g, qmcur, gm, cm);
                    else k_grow_mid<2, MID_R><<<(gm + 1) / 2, 64, mid_smem, st>>>(g, qmcur, gm, cm);
                    CKL();
                    h->stats.kernel_launches++; h->stats.launches_grow_mid++;
                }
                if (gmc[1]) {
                    const uint32_t nb = (gmc[1] + MID_WARPS - 1) / MID_WARPS;
                    if (reg) k_grow_mid4<MID_WARPS, 4, true, MIDC1_MINB_N><<<nb, MID_WARPS * 32, midc_smem[1], st>>>(g, qmccur[1], gmc[1], cmc[1]);
                    else k_grow_mid4<MID_WARPS, 4, false, MIDC1_MINB_N><<<nb, MID_WARPS * 32, midc_smem[1], st>>>(g, qmccur[1], gmc[1], cmc[1]);
                    CKL();
                    h->stats.kernel_launches++; h->stats.launches_grow_mid++;
                }
                if (gmc[0]) {
                    const uint32_t nb = (gmc[0] + MID_WARPS - 1) / MID_WARPS;
                    if (reg) k_grow_mid4<MID_WARPS, 2, true, MIDC0_MINB_N><<<nb, MID_WARPS * 32, midc_smem[0], st>>>(g, qmccur[0], gmc[0], cmc[0]);
                    else k_grow_mid4<MID_WARPS, 2, false, MIDC0_MINB_N><<<nb, MID_WARPS * 32, midc_smem[0], st>>>(g, qmccur[0], gmc[0], cmc[0]);
                    CKL();
                    h->stats.kernel_launches++; h->stats.launches_grow_mid++;
                }
                CK(cudaEventRecord(h->ev_lvl[j][2], st));
                h->stats.levels++;
                cur ^= 1;
                const uint64_t ub_cta = ub_big + ub_bigl;  // parents of the CTA tier, either queue
                const uint64_t nb2 = 2 * ub_big, nbl2 = 2 * ub_cta, nw2 = 2 * (ub_cta + ub_wide), nm2 = 2 * (ub_cta + ub_wide + ub_mid), nmc = 2 * (ub_cta + ub_wide + ub_mid + ub_midc[0] + ub_midc[1]);
                ub_big = nb2; ub_bigl = nbl2; ub_wide = nw2; ub_mid = nm2; ub_midc[0] = ub_midc[1] = nmc;  // (a child of any queued node can land in either class)
                if (++level_guard > n_cap + 8) return fail(TST_ERR_STATE, "tree growth did not terminate");
            }
            CK(cudaMemcpyAsync(h->h_pinned, counters, 128, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            {   // every level is timed, also the ones launched blind (the top levels of a small batch of trees are "thin" too)
                float ms_b = 0, ms_m = 0, ms_w = 0;
                for (int j = 0; j < J; ++j) {
                    float ms = 0;
                    cudaEventElapsedTime(&ms, h->ev_lvl[j][0], h->ev_lvl[j][1]); ms_b += ms;
                    cudaEventElapsedTime(&ms, h->ev_lvl[j][1], h->ev_lvl[j][3]); ms_w += ms;
                    cudaEventElapsedTime(&ms, h->ev_lvl[j][3], h->ev_lvl[j][2]); ms_m += ms;
                }
                h->stats.ms_grow_big += ms_b; h->stats.ms_grow_mid += ms_m; h->stats.ms_grow_wide += ms_w;
                if (trace) { tr_gap += ms_since(t_level) - ms_b - ms_m - ms_w; ++tr_groups; }
            }
            nbig = h->h_pinned[2 * cur];
            nmid = h->h_pinned[2 * cur + 1];
            nwide = g.wide_max ? h->h_pinned[16 + cur] : 0u;
            for (int c = 0; c < 2; ++c) nmidc[c] = g.midc_max[c] ? h->h_pinned[22 + 2 * c + cur] : 0u;
            nbigl = g.bigl ? h->h_pinned[26 + cur] : 0u;
            if (nbigl > g.qbig_cap || nbig > g.qbig_cap || nmid > g.qbig_cap || nwide > g.qwide_cap || nmidc[0] > g.qbig_cap || nmidc[1] > g.qbig_cap) return fail(TST_ERR_OVERFLOW, "node queue overflowed");
        }
        tr_levels = ms_since(t_batch);
        phase.next("tst subtree tiers + inference");
        // ---- subtrees of 33 .. sub_max rows: one CTA each, rows cached in shared memory, every level of the subtree in one launch
        const uint32_t nsub = g.sub_max ? h->h_pinned[20] : 0u;
        if (nsub > g.qbig_cap) return fail(TST_ERR_OVERFLOW, "subtree queue overflowed");
        if (nsub) {
            CK(cudaEventRecord(h->evk0, st));
            if (g.sub_max == 2048u) k_grow_sub<16, 2048, 1><<<nsub, 512, sub_smem, st>>>(g, g.qsub, nsub);
            else k_grow_sub<8, 1024, 2><<<nsub, 256, sub_smem, st>>>(g, g.qsub, nsub);
            CKL();
            CK(cudaEventRecord(h->evk1, st));
            sub_timed = true;
            h->stats.kernel_launches++; h->stats.launches_grow_sub++;
        }
        // ---- subtrees with 7..32 rows: one warp each
        const uint32_t nsmall = h->h_pinned[4];
        if (nsmall > g.qsmall_cap) return fail(TST_ERR_OVERFLOW, "small-node queue overflowed");
        if (nsmall) {
            uint32_t blocks = std::min<uint32_t>((nsmall + SMALL_WARPS - 1) / SMALL_WARPS, (uint32_t)h->sm_count * 64u);
            CK(cudaEventRecord(h->evs0, st));
            small_started = true;
            if (reg && small_multi) k_grow_small4<SMALL_WARPS, true, true><<<blocks, SMALL_WARPS * 32, small_smem, st>>>(g, g.qsmall, nsmall);
            else if (reg) k_grow_small4<SMALL_WARPS, true><<<blocks, SMALL_WARPS * 32, small_smem, st>>>(g, g.qsmall, nsmall);
            else if (small_packed && small_multi) k_grow_small4<SMALL_WARPS, false, true><<<blocks, SMALL_WARPS * 32, small_smem, st>>>(g, g.qsmall, nsmall);
            else if (small_packed) k_grow_small4<SMALL_WARPS><<<blocks, SMALL_WARPS * 32, small_smem, st>>>(g, g.qsmall, nsmall);
            else k_grow_small<SMALL_WARPS><<<blocks, SMALL_WARPS * 32, small_smem, st>>>(g, g.qsmall, nsmall);
            CKL();
            CK(cudaEventRecord(h->evs1, st));
            h->stats.kernel_launches++; h->stats.launches_grow_small++;
        }
        // ---- subtrees with <= TINY_MAX rows: one thread each (tasks come from every other tier)
        if (g.tiny_max) {
            CK(cudaMemcpyAsync(h->h_pinned, counters, 128, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            const uint32_t ntiny = h->h_pinned[5];
            if (ntiny > g.qtiny_cap) return fail(TST_ERR_OVERFLOW, "tiny-node queue overflowed");
            if (ntiny) {
                // bucket offsets, largest subtrees first: counters[6 + r] = number of tasks with r rows, counters[32 + r] = cursor
                CK(cudaEventRecord(h->evm0, st));
                k_tiny_hist<<<(ntiny + 255) / 256, 256, 0, st>>>((const uint4*)h->qtiny.p, ntiny, counters + 6);
                CKL();
                k_tiny_cursors<<<1, 32, 0, st>>>(counters + 6, TINY_MAX, counters + 32);
                CKL();
                k_tiny_sort<<<(ntiny + 255) / 256, 256, 0, st>>>((const uint4*)h->qtiny.p, ntiny, counters + 32, (uint4*)h->qtiny2.p);
                CKL();
                if (reg) k_grow_tiny4<TINY_MAX, true><<<(ntiny + 127) / 128, 128, 0, st>>>(g, (const uint4*)h->qtiny2.p, ntiny);
                else k_grow_tiny4<TINY_MAX><<<(ntiny + 127) / 128, 128, 0, st>>>(g, (const uint4*)h->qtiny2.p, ntiny);
                CKL();
                CK(cudaEventRecord(h->evm1, st));
                tiny_timed = true;
                h->stats.kernel_launches += 4; h->stats.launches_grow_tiny++;
            }
        }
        // ---- OOF inference of this batch's trees on their held-out rows
        std::vector<InferPart> ips;
        uint32_t max_fsz = 0;
        for (uint32_t i = 0; i < nb;) {
            uint32_t j = i;
            while (j < nb && all[b0 + j].part == all[b0 + i].part) ++j;
            if (all[b0 + i].fsz) {
                ips.push_back({all[b0 + i].fs, all[b0 + i].fsz, i, j});
                max_fsz = std::max(max_fsz, all[b0 + i].fsz);
            }
            i = j;
        }
        bool infer_timed = false;
        if (votes_dev && !ips.empty()) {
            CK(cudaMemcpyAsync(h->infer_parts.p, ips.data(), ips.size() * sizeof(InferPart), cudaMemcpyHostToDevice, st));
            CK(cudaEventRecord(h->evi0, st));
            infer_timed = true;
            // enough threads to fill the GPU: cut the trees of a partition into chunks when the held-out sets are small
            uint32_t max_trees = 0, rows_total = 0;
            for (auto& ip : ips) { max_trees = std::max(max_trees, ip.tree_hi - ip.tree_lo); rows_total += ip.fsz; }
            uint32_t chunks = 1;
            const uint32_t want_threads = (uint32_t)h->sm_count * 2048u * 2u;
            if (rows_total < want_threads) chunks = std::min<uint32_t>(max_trees, std::max<uint32_t>(1u, want_threads / std::max(1u, rows_total)));
            const uint32_t tpc = (max_trees + chunks - 1) / chunks;
            chunks = (max_trees + tpc - 1) / tpc;
            dim3 grid((max_fsz + 255) / 256, (unsigned)ips.size(), chunks);
            if (chunks > 1) k_oof_infer<true><<<grid, 256, 0, st>>>(g, h->infer_parts.as<InferPart>(), h->sorted.as<uint32_t>(), votes_dev, tpc);
            else k_oof_infer<false><<<grid, 256, 0, st>>>(g, h->infer_parts.as<InferPart>(), h->sorted.as<uint32_t>(), votes_dev, tpc);
            CKL();
            CK(cudaEventRecord(h->evi1, st));
            h->stats.kernel_launches++;
        }
        CK(cudaStreamSynchronize(st));
        if (small_started) { float ms = 0; cudaEventElapsedTime(&ms, h->evs0, h->evs1); h->stats.ms_grow_small += ms; }
        if (tiny_timed) { float ms = 0; cudaEventElapsedTime(&ms, h->evm0, h->evm1); h->stats.ms_grow_tiny += ms; }
        if (sub_timed) { float ms = 0; cudaEventElapsedTime(&ms, h->evk0, h->evk1); h->stats.ms_grow_sub += ms; }
        if (infer_timed) { float ms = 0; cudaEventElapsedTime(&ms, h->evi0, h->evi1); h->stats.ms_infer += ms; }
        if (keep) {
            std::vector<TreeDesc> back(nb);
            CK(cudaMemcpy(back.data(), h->trees.p, (size_t)nb * sizeof(TreeDesc), cudaMemcpyDeviceToHost));
            for (uint32_t i = 0; i < nb; ++i)
                h->kept.push_back({(int32_t)all_id[b0 + i].part, (int32_t)all_id[b0 + i].t, back[i].arena, back[i].n_inbag});
        }
        h->stats.trees += nb;
        h->imp_trees = (int64_t)(b0 + nb);
        if ((rc = pull_stats(h, st))) return rc;
        tr_tail = ms_since(t_batch);
        if (trace) fprintf(stderr, "[tst trace] batch of %u trees: boot+roots %.3f ms, levels %.3f ms (%d host round trips, %.3f ms outside the level kernels), small+tiny+infer+stats %.3f ms\n",
                           nb, tr_boot, tr_levels - tr_boot, tr_groups, tr_gap, tr_tail - tr_levels);
    }
    if (trace) fprintf(stderr, "[tst trace] fit_oof total %.3f ms\n", ms_since(t_entry));
    return TST_OK;
}

// --------------------------------------------------------------------------------------------- votes -> s(x)
extern "C" int tst_reduce_votes(tst_handle* h, const uint32_t* gathered_dev, int32_t world, int64_t nv, uint32_t* out_dev, void* stream) {
    if (!h || !gathered_dev || !out_dev || world < 1 || nv < 0) return fail(TST_ERR_INVALID, "tst_reduce_votes: bad argument");
    CK(cudaSetDevice(h->device));
    if (nv) k_reduce_votes<<<(unsigned)((nv + 255) / 256), 256, 0, (cudaStream_t)stream>>>(gathered_dev, world, nv, out_dev);
    CKL();
    h->stats.kernel_launches++;
    return TST_OK;
}

extern "C" int tst_allreduce_votes(tst_handle* h, uint32_t* votes_dev, int64_t nv, void* stream) {
    if (!h || !votes_dev || nv < 0) return fail(TST_ERR_INVALID, "tst_allreduce_votes: bad argument");
    if (h->world == 1 || nv == 0) return TST_OK;
    if (!h->comm) return fail(TST_ERR_STATE, "tst_allreduce_votes: the handle was created without an NCCL unique id");
    CK(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    ENSURE(h->gather, (size_t)h->world * nv * 4);
    int rc = g_nccl.AllGather(votes_dev, h->gather.p, (size_t)nv, NCCL_UINT32, h->comm, st);
    if (rc) return fail(TST_ERR_CUDA, std::string("ncclAllGather: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error"));
    if (g_nccl.CommGetAsyncError) {  // a peer that died or a link error surfaces here instead of as a hang in the caller's next sync
        int aerr = 0;
        if (g_nccl.CommGetAsyncError(h->comm, &aerr) == 0 && aerr != 0)
            return fail(TST_ERR_CUDA, std::string("ncclCommGetAsyncError: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(aerr) : "error"));
    }
    k_reduce_votes<<<(unsigned)((nv + 255) / 256), 256, 0, st>>>(h->gather.as<uint32_t>(), h->world, nv, votes_dev);
    CKL();
    h->stats.kernel_launches++;
    return TST_OK;
}

extern "C" int tst_finalize(tst_handle* h, const uint32_t* votes_dev, const uint8_t* y_dev, const uint8_t* fold_dev,
                            const int32_t* group_dev, int32_t n_groups, int32_t k, int32_t n_trees, int32_t C, float* s_dev,
                            uint32_t* correct_dev, uint32_t* total_dev, void* stream) {
    if (!h || !votes_dev || !y_dev || !fold_dev) return fail(TST_ERR_INVALID, "tst_finalize: null argument");
    if (!h->n) return fail(TST_ERR_STATE, "tst_finalize: no table");
    CK(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (correct_dev) CK(cudaMemsetAsync(correct_dev, 0, (size_t)n_groups * k * 4, st));
    if (total_dev) CK(cudaMemsetAsync(total_dev, 0, (size_t)n_groups * k * 4, st));
    double denom = (double)n_trees * 65536.0;
    k_finalize<<<(unsigned)((h->n + 255) / 256), 256, 0, st>>>(votes_dev, y_dev, fold_dev, group_dev, h->n, C, k, denom, s_dev, correct_dev, total_dev);
    CKL();
    h->stats.kernel_launches++;
    return TST_OK;
}

// One call for the whole diagnostic on device buffers: folds of the active rows, this rank's share of the k x n_trees forests,
// the vote exchange when the handle has a communicator, s(x).  The signature SURVEY 8(b) sketched for tst_fit_oof
// (active mask in, votes and s out); the separate entry points above stay for callers that keep their own folds.
extern "C" int tst_run_oof(tst_handle* h, const uint8_t* y_dev, const uint8_t* active_dev, const int32_t* group_dev, int32_t n_groups,
                           int32_t k, int32_t n_trees, int32_t C, const tst_params* prm, uint64_t seed, uint8_t* fold_out_dev,
                           uint32_t* votes_dev, float* s_dev, void* stream) {
    if (!h || !y_dev || !active_dev || !prm || !fold_out_dev || !votes_dev) return fail(TST_ERR_INVALID, "tst_run_oof: null argument");
    if (!h->n) return fail(TST_ERR_STATE, "tst_run_oof: bin the table first");
    if (h->world != 1 && !h->comm) return fail(TST_ERR_STATE, "tst_run_oof: world > 1 needs a handle created with an NCCL unique id");
    if (prm->criterion == TST_SQUARED_ERROR && s_dev) return fail(TST_ERR_UNSUPPORTED, "tst_run_oof: s(x) is defined on class votes; pass s_dev = NULL for the squared-error criterion");
    CK(cudaSetDevice(h->device));
    int rc = tst_assign_folds(h, y_dev, active_dev, group_dev, n_groups, k, C, fold_out_dev, stream);
    if (rc) return rc;
    CK(cudaMemsetAsync(votes_dev, 0, (size_t)h->n * C * 4, (cudaStream_t)stream));
    if ((rc = tst_fit_oof(h, y_dev, fold_out_dev, group_dev, n_groups, k, n_trees, C, prm, seed, 0u, 0, n_trees, votes_dev, stream))) return rc;
    if ((rc = tst_allreduce_votes(h, votes_dev, (int64_t)h->n * C, stream))) return rc;
    if (s_dev) rc = tst_finalize(h, votes_dev, y_dev, fold_out_dev, group_dev, n_groups, k, n_trees, C, s_dev, nullptr, nullptr, stream);
    return rc;
}

// --------------------------------------------------------------------------------------------- IBP
extern "C" int tst_ibp_select(tst_handle* h, const uint32_t* votes_dev, const uint8_t* y_dev, const uint8_t* fold_dev, int32_t k,
                              int32_t C, uint64_t vmin, int64_t batch, uint8_t* active_dev, int64_t* removed_host, void* stream) {
    if (!h || !votes_dev || !y_dev || !fold_dev || !active_dev || !removed_host) return fail(TST_ERR_INVALID, "tst_ibp_select: null argument");
    if (!h->n) return fail(TST_ERR_STATE, "tst_ibp_select: no table");
    CK(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    ENSURE(h->ibp_removed, 8);
    k_ibp_select<1024><<<1, 1024, 0, st>>>(votes_dev, y_dev, fold_dev, h->n, C, k, (unsigned long long)vmin, (long long)batch, active_dev,
                                          h->ibp_removed.as<long long>());
    CKL();
    h->stats.kernel_launches++;
    long long r = 0;
    CK(cudaMemcpyAsync(&r, h->ibp_removed.p, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    *removed_host = r;
    return TST_OK;
}

__global__ void k_fill_f32(float* p, int64_t n, float v) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

extern "C" int tst_ibp(tst_handle* h, const uint8_t* y_dev, uint8_t* active_dev, const int32_t* group_dev, int32_t n_groups,
                       int32_t k, int32_t n_trees, int32_t C, const tst_params* prm, uint64_t seed, int32_t rounds, double tau,
                       int64_t batch, int64_t* removed_host, float* s_per_round_dev, void* stream) {
    if (!h || !y_dev || !active_dev || !prm || !removed_host || rounds < 1) return fail(TST_ERR_INVALID, "tst_ibp: bad argument");
    if (!h->n) return fail(TST_ERR_STATE, "tst_ibp: bin the table first");
    if (prm->criterion == TST_SQUARED_ERROR) return fail(TST_ERR_UNSUPPORTED, "tst_ibp: pruning is defined on class votes; the squared-error criterion has none");
    if (h->world != 1 && !h->comm) return fail(TST_ERR_STATE, "tst_ibp: world > 1 needs a handle created with an NCCL unique id (or drive the rounds from the host, INTEGRATION.md)");
    CK(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n = h->n;
    ENSURE(h->ibp_fold, (size_t)n);
    ENSURE(h->ibp_votes, (size_t)n * C * 4);
    const uint64_t vmin = (uint64_t)std::ceil(tau * (double)n_trees * 65536.0);
    const uint64_t per_round = (uint64_t)n_groups * k * n_trees;
    if (per_round * rounds > 0xFFFFFFFFull) return fail(TST_ERR_INVALID, "tst_ibp: tree id overflows 32 bits");
    for (int r = 0; r < rounds; ++r) removed_host[r] = 0;
    if (s_per_round_dev) {
        int64_t tot = (int64_t)rounds * n;
        k_fill_f32<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(s_per_round_dev, tot, -1.0f);
        CKL();
    }
    for (int r = 0; r < rounds; ++r) {
        int rc = tst_assign_folds(h, y_dev, active_dev, group_dev, n_groups, k, C, h->ibp_fold.as<uint8_t>(), stream);
        if (rc) return rc;
        CK(cudaMemsetAsync(h->ibp_votes.p, 0, (size_t)n * C * 4, st));
        rc = tst_fit_oof(h, y_dev, h->ibp_fold.as<uint8_t>(), group_dev, n_groups, k, n_trees, C, prm, seed, (uint32_t)(r * per_round), 0,
                         n_trees, h->ibp_votes.as<uint32_t>(), stream);
        if (rc) return rc;
        if ((rc = tst_allreduce_votes(h, h->ibp_votes.as<uint32_t>(), (int64_t)n * C, stream))) return rc;
        if (s_per_round_dev) {
            rc = tst_finalize(h, h->ibp_votes.as<uint32_t>(), y_dev, h->ibp_fold.as<uint8_t>(), group_dev, n_groups, k, n_trees, C,
                              s_per_round_dev + (size_t)r * n, nullptr, nullptr, stream);
            if (rc) return rc;
        }
        int64_t took = 0;
        rc = tst_ibp_select(h, h->ibp_votes.as<uint32_t>(), y_dev, h->ibp_fold.as<uint8_t>(), k, C, vmin, batch, active_dev, &took, stream);
        if (rc) return rc;
        removed_host[r] = took;
        if (!took) break;
    }
    return TST_OK;
}

// --------------------------------------------------------------------------------------------- kept forest
extern "C" int tst_predict(tst_handle* h, const float* X_dev, int64_t n, uint32_t* votes_dev, void* stream) {
    if (!h || !X_dev || !votes_dev || n < 1) return fail(TST_ERR_INVALID, "tst_predict: bad argument");
    if (h->kept.empty()) return fail(TST_ERR_STATE, "tst_predict: no kept forest (fit with keep_forest = 1)");
    if (!h->have_edges) return fail(TST_ERR_STATE, "tst_predict: the table was not binned by tst_bin, no edges to bin new rows with");
    CK(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int d = h->d;
    if (int rcw = wait_bins(h, st)) return rcw;  // the edges come from the binning stream
    int64_t ld = (n + 15) / 16 * 16;
    ENSURE(h->pred_cm, (size_t)ld * d);
    ENSURE(h->pred_rm, (size_t)n * h->dstride);
    CK(cudaMemsetAsync(h->pred_rm.p, 0, (size_t)n * h->dstride, st));
    k_bin_map<<<(unsigned)((n + BIN_TILE - 1) / BIN_TILE), 256, (size_t)d * (BIN_TILE + 4), st>>>(
        X_dev, n, d, h->edges.as<double>(), h->n_edges.as<int>(), h->pred_cm.as<uint8_t>(), ld, h->pred_rm.as<uint8_t>(), h->dstride);
    CKL();
    std::vector<uint32_t> arenas;
    for (auto& kt : h->kept) arenas.push_back(kt.arena);
    ENSURE(h->keep_arenas, arenas.size() * 4);
    CK(cudaMemcpyAsync(h->keep_arenas.p, arenas.data(), arenas.size() * 4, cudaMemcpyHostToDevice, st));
    k_predict<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(h->pred_rm.as<uint8_t>(), h->dstride, n, h->kept_C, h->keep_nodes.as<uint2>(),
                                                         h->kept_node_stride, h->keep_leafc.as<uint32_t>(), h->kept_leafc_stride,
                                                         h->keep_arenas.as<uint32_t>(), (uint32_t)arenas.size(), votes_dev, h->kept_criterion == TST_SQUARED_ERROR);
    CKL();
    CK(cudaStreamSynchronize(st));  // `arenas` is a stack-owned staging buffer
    h->stats.kernel_launches += 2;
    return TST_OK;
}

extern "C" int tst_export_tree(tst_handle* h, int32_t part, int32_t t, tst_node* nodes_host, uint32_t* counts_host, int64_t cap,
                               int64_t* n_nodes) {
    if (!h || !n_nodes) return fail(TST_ERR_INVALID, "tst_export_tree: null argument");
    const KeptTree* kt = nullptr;
    for (auto& x : h->kept) if (x.part == part && x.t == t) { kt = &x; break; }
    if (!kt) return fail(TST_ERR_STATE, "tst_export_tree: tree not kept (keep_forest = 0, other rank, or out of range)");
    CK(cudaSetDevice(h->device));
    const int C = h->kept_C;
    const bool regx = h->kept_criterion == TST_SQUARED_ERROR;
    const uint64_t nslots = std::max<uint64_t>(1, 2ull * kt->n_inbag);
    const uint64_t nleafc = ((uint64_t)kt->n_inbag / 2 + 1) * C;
    std::vector<uint2> nodes(nslots);
    std::vector<uint32_t> leafc(nleafc);
    CK(cudaMemcpy(nodes.data(), h->keep_nodes.as<uint2>() + (uint64_t)kt->arena * h->kept_node_stride, nslots * sizeof(uint2), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(leafc.data(), h->keep_leafc.as<uint32_t>() + (uint64_t)kt->arena * h->kept_leafc_stride, nleafc * 4, cudaMemcpyDeviceToHost));
    struct Item { uint32_t slot, begin, end, depth; };
    std::vector<Item> q;
    q.push_back({0, 0, kt->n_inbag, 0});
    std::vector<tst_node> out;
    std::vector<uint32_t> cnt;
    for (size_t i = 0; i < q.size(); ++i) {
        Item it = q[i];
        if (it.slot >= nslots) return fail(TST_ERR_STATE, "tst_export_tree: corrupt node slot");
        uint2 rec = nodes[it.slot];
        uint32_t kind = rec.y >> 30;
        tst_node nd;
        nd.left = nd.right = -1; nd.feature = -1; nd.bin = -1;
        nd.begin = it.begin; nd.end = it.end; nd.depth = it.depth; nd.n_weighted = 0;
        cnt.resize(cnt.size() + C, 0u);
        uint32_t* c = cnt.data() + i * C;
        if (kind == KIND_INTERNAL) {
            uint32_t nl = (rec.x - it.slot) / 2;
            if (rec.x <= it.slot || nl == 0 || it.begin + nl >= it.end) return fail(TST_ERR_STATE, "tst_export_tree: corrupt internal node");
            nd.feature = (int32_t)((rec.y >> 8) & 0xFFFFu); nd.bin = (int32_t)(rec.y & 0xFFu);
            nd.left = (int32_t)q.size(); nd.right = (int32_t)q.size() + 1;
            q.push_back({it.slot + 1, it.begin, it.begin + nl, it.depth + 1});
            q.push_back({rec.x, it.begin + nl, it.end, it.depth + 1});
        } else if (kind == KIND_PURE && regx) {  // criterion 2: counts = (S low, S high), S = sum of w * yq
            const uint64_t S = (uint64_t)rec.x * (rec.y & 0xFFFFu);
            c[0] = (uint32_t)S; c[1] = (uint32_t)(S >> 32); nd.n_weighted = rec.x;
        } else if (kind == KIND_MIXED && regx) {
            c[0] = leafc[(uint64_t)rec.x * 2]; c[1] = leafc[(uint64_t)rec.x * 2 + 1]; nd.n_weighted = rec.y & 0x3FFFFFFFu;
        } else if (kind == KIND_PURE) {
            c[rec.y & 0xFFu] = rec.x; nd.n_weighted = rec.x;
        } else if (kind == KIND_MIXED) {
            for (int k2 = 0; k2 < C; ++k2) { c[k2] = leafc[(uint64_t)rec.x * C + k2]; nd.n_weighted += c[k2]; }
        } else {
            return fail(TST_ERR_STATE, "tst_export_tree: unwritten node slot " + std::to_string(it.slot));
        }
        out.push_back(nd);
    }
    for (size_t i = out.size(); i-- > 0;) {
        if (out[i].left >= 0 && regx) {  // 64-bit sums in two words
            const size_t l = (size_t)out[i].left * 2, r2 = (size_t)out[i].right * 2;
            const uint64_t S = ((uint64_t)cnt[l] | ((uint64_t)cnt[l + 1] << 32)) + ((uint64_t)cnt[r2] | ((uint64_t)cnt[r2 + 1] << 32));
            cnt[i * 2] = (uint32_t)S; cnt[i * 2 + 1] = (uint32_t)(S >> 32);
            out[i].n_weighted = out[out[i].left].n_weighted + out[out[i].right].n_weighted;
        } else if (out[i].left >= 0) {
            for (int k2 = 0; k2 < C; ++k2) cnt[i * C + k2] = cnt[(size_t)out[i].left * C + k2] + cnt[(size_t)out[i].right * C + k2];
            out[i].n_weighted = out[out[i].left].n_weighted + out[out[i].right].n_weighted;
        }
    }
    *n_nodes = (int64_t)out.size();
    if (!nodes_host || !counts_host || cap < (int64_t)out.size()) return fail(TST_ERR_INVALID, "tst_export_tree: capacity too small");
    memcpy(nodes_host, out.data(), out.size() * sizeof(tst_node));
    memcpy(counts_host, cnt.data(), out.size() * C * 4);
    return TST_OK;
}

// --------------------------------------------------------------------------------------------- interpretability
extern "C" int tst_feature_importances(tst_handle* h, uint64_t* sums_host, int64_t* n_counted, uint64_t* raw_host, int64_t raw_cap_trees) {
    if (!h || !sums_host || !n_counted) return fail(TST_ERR_INVALID, "tst_feature_importances: null argument");
    if (h->imp_d <= 0) return fail(TST_ERR_STATE, "tst_feature_importances: no fit yet");
    CK(cudaSetDevice(h->device));
    const int d = h->imp_d;
    const int64_t T = h->imp_trees;
    for (int f = 0; f < d; ++f) sums_host[f] = 0;
    *n_counted = 0;
    if (T == 0) return TST_OK;
    ENSURE(h->imp_sums, (size_t)(d + 1) * 8);
    CK(cudaMemsetAsync(h->imp_sums.p, 0, (size_t)(d + 1) * 8, 0));
    k_imp_reduce<<<(unsigned)((T + 127) / 128), 128>>>(h->imp.as<unsigned long long>(), T, d, h->imp_sums.as<unsigned long long>());
    CKL();
    h->stats.kernel_launches++;
    std::vector<uint64_t> sums((size_t)d + 1);
    CK(cudaMemcpy(sums.data(), h->imp_sums.p, (size_t)(d + 1) * 8, cudaMemcpyDeviceToHost));
    memcpy(sums_host, sums.data(), (size_t)d * 8);
    *n_counted = (int64_t)sums[d];
    if (raw_host) {
        if (raw_cap_trees < T) return fail(TST_ERR_INVALID, "tst_feature_importances: raw buffer holds fewer trees than the last fit built");
        CK(cudaMemcpy(raw_host, h->imp.p, (size_t)T * d * 8, cudaMemcpyDeviceToHost));
    }
    return TST_OK;
}

extern "C" int tst_get_stats(tst_handle* h, tst_stats* out) {
    if (!h || !out) return fail(TST_ERR_INVALID, "tst_get_stats: null argument");
    *out = h->stats;
    return TST_OK;
}

extern "C" int tst_reset_stats(tst_handle* h) {
    if (!h) return fail(TST_ERR_INVALID, "tst_reset_stats: null argument");
    memset(&h->stats, 0, sizeof(h->stats));
    return TST_OK;
}
