"""Host side of the TsT Random-Forest diagnostic: a sklearn-like surface over the libtst.so C ABI.

    TsTForest(...).fit(X, y).predict_proba(X)      names/defaults: SK/ensemble/_forest.py:1515-1537
    run_tst(X, y, k, n_trees, seed)  -> TsTResult   k-fold OOF bias score s(x)   (REF/README.md:35)
    ibp(X, y, rounds, k, n_trees, tau, batch)       Iterative Bias Pruning       (REF/README.md:35)

torch is used only to own device memory, streams and the one NCCL all-gather; all compute is in
the hand-written CUDA kernels behind include/tst.h.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _lib
from . import binning
from ._lib import FOLD_INACTIVE, FOLD_TRAIN_ONLY, Params, Stats, TstError, check


def _torch():
    import torch

    return torch


def resolve_max_features(max_features, d: int) -> int:
    """SK/tree/_classes.py:330-345."""
    if max_features is None:
        return d
    if isinstance(max_features, str):
        if max_features == "sqrt":
            return max(1, int(np.sqrt(d)))
        if max_features == "log2":
            return max(1, int(np.log2(d)))
        raise ValueError(f"max_features={max_features!r}")
    if isinstance(max_features, float):
        return max(1, min(d, int(max_features * d)))
    return max(1, min(d, int(max_features)))


CRITERIA = {"gini": 0, "squared_error_levels": 1, "squared_error": 2}  # include/tst.h: TST_GINI, TST_SQUARED_ERROR_LEVELS, TST_SQUARED_ERROR


def make_params(d, max_depth=None, min_samples_split=2, min_samples_leaf=1, max_features="sqrt",
                trees_in_flight=0, keep_forest=False, criterion="gini") -> Params:
    if min_samples_split < 2 or min_samples_leaf < 1:
        raise ValueError("min_samples_split >= 2 and min_samples_leaf >= 1 required")
    if criterion not in CRITERIA:
        raise ValueError(f"criterion must be one of {sorted(CRITERIA)}")
    return Params(0 if max_depth is None else int(max_depth), int(min_samples_split), int(min_samples_leaf),
                  resolve_max_features(max_features, d), int(trees_in_flight), int(bool(keep_forest)), CRITERIA[criterion])


def encode_labels(y):
    """Class order = np.unique(y) (SK/ensemble/_forest.py:810-827)."""
    y = np.asarray(y)
    if y.dtype.kind in "ui" and y.size and int(y.min()) >= 0 and int(y.max()) < 4096:
        # small non-negative integers (the common case): a bincount is 20x cheaper than a sort-based unique at 1M rows
        hi = int(y.max())
        if hi < 16:  # a handful of classes: one vectorised compare per class beats the int64 cast inside bincount (0.7 vs 2.9 ms at 1M rows)
            counts = np.array([np.count_nonzero(y == c) for c in range(hi + 1)])
        else:
            counts = np.bincount(y.ravel())
        classes = np.flatnonzero(counts).astype(y.dtype)
        if classes.shape[0] <= _lib.MAX_CLASSES:
            if y.dtype == np.uint8 and classes.shape[0] == counts.shape[0]:
                return classes, np.ascontiguousarray(y.ravel())  # labels are 0 .. C-1 already: no copy (6 ms of LUT gathers at 1M rows)
            lut = np.zeros(counts.shape[0], dtype=np.uint8)
            lut[classes] = np.arange(classes.shape[0], dtype=np.uint8)
            return classes, lut[y.ravel()]
        idx = np.zeros(0, dtype=np.int64)  # too many classes: the error below
    else:
        classes, idx = np.unique(y, return_inverse=True)
    if classes.shape[0] > _lib.MAX_CLASSES:
        raise ValueError(f"at most {_lib.MAX_CLASSES} classes are supported")
    return classes, idx.astype(np.uint8)


def dist_info():
    try:
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except Exception:  # pragma: no cover
        pass
    return 0, 1


def shard_trees(n_parts: int, n_trees: int, rank: int, world: int, t0: int = 0, t1: int | None = None):
    """The (part, t) pairs rank `rank` of `world` builds: local id part*n_trees + t == rank (mod world)."""
    t1 = n_trees if t1 is None else t1
    return [(p, t) for p in range(n_parts) for t in range(t0, t1) if (p * n_trees + t) % world == rank]


class Engine:
    """One tst_handle = one (process, device).  All tensors handed to it live on that device."""

    def __init__(self, device: int | None = None, rank: int = 0, world: int = 1, lib_nccl: bool | None = None):
        """lib_nccl: let libtst own the NCCL communicator (unique id broadcast through torch.distributed) so the vote
        exchange and the multi-rank IBP loop run inside the library; default: on when torch.distributed is initialised
        with `world` ranks, off for emulated ranks (tests)."""
        torch = _torch()
        if not torch.cuda.is_available():
            raise TstError(-2, "no CUDA device: the TsT engine has no CPU fallback")
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.dev = torch.device("cuda", self.device)
        self.rank, self.world = rank, world
        self._h = C.c_void_p()
        uid = None
        if lib_nccl is None:
            lib_nccl = world > 1 and dist_info() == (rank, world)
        if lib_nccl and world > 1:
            import torch.distributed as dist

            buf = (C.c_char * 128)()
            if rank == 0:
                check(_lib.lib().tst_nccl_unique_id(buf))
            t = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).to(self.dev)
            dist.broadcast(t, src=0)
            uid = (C.c_char * 128).from_buffer_copy(bytes(t.cpu().numpy().tobytes()))
        self.lib_nccl = uid is not None
        with torch.cuda.device(self.dev):
            check(_lib.lib().tst_create(C.byref(self._h), self.device, rank, world, uid))
        self.n = self.d = 0
        self._edges = self._n_edges = None
        self._edges_lazy = False
        self._side = None  # side stream: H2D copy + binning of a host table overlap whatever the caller queues next

    def close(self):
        if self._h:
            _lib.lib().tst_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- bin edges: fetched from the device on first use (the fit does not need them on the host)
    def _fetch_edges(self):
        if self._edges_lazy:
            edges = np.zeros((self.d, binning.MAX_EDGES), np.float64)
            n_edges = np.zeros(self.d, np.int32)
            check(_lib.lib().tst_get_edges(self._h, edges.ctypes.data_as(C.c_void_p), n_edges.ctypes.data_as(C.c_void_p)))
            self._edges, self._n_edges, self._edges_lazy = edges, n_edges, False

    @property
    def edges(self):
        self._fetch_edges()
        return self._edges

    @property
    def n_edges(self):
        self._fetch_edges()
        return self._n_edges

    def set_option(self, name: str, value: int):
        """Tuning / diagnostic knob of this handle (include/tst.h: tst_set_option)."""
        check(_lib.lib().tst_set_option(self._h, name.encode(), C.c_int64(int(value))))
        return self

    # ---- helpers
    def _stream(self):
        return C.c_void_p(_torch().cuda.current_stream(self.dev).cuda_stream)

    def to_dev(self, a, dtype):
        torch = _torch()
        if isinstance(a, torch.Tensor):
            return a.to(device=self.dev, dtype=dtype, non_blocking=True).contiguous()
        t = torch.from_numpy(np.ascontiguousarray(a))
        return t.to(device=self.dev, dtype=dtype, non_blocking=True).contiguous()

    @staticmethod
    def _ptr(t):
        return C.c_void_p(0 if t is None else t.data_ptr())

    # ---- table
    def bin(self, X, n_bins: int = 256, edges=None, n_edges=None, host_edges: bool = False):
        """X: host array or device tensor [n, d].  By default the bin edges are selected on the GPU
        (tst_bin_auto: per-column radix sort + integer-rank quantiles); `edges=` / `host_edges=True` use the
        host rule of binning.py (bit-identical, kept as the cross-check)."""
        torch = _torch()
        if edges is None and not host_edges:
            # Host tables: the H2D copy and the binning kernels go to a side stream and the call returns at once; whatever the
            # caller queues next that does not need the table (label upload, fold assignment, the bootstrap inside tst_fit_oof)
            # overlaps them, and tst_fit_oof waits for the binned table on the device (include/tst.h: tst_get_edges).
            on_host = not (isinstance(X, torch.Tensor) and X.is_cuda) and int(np.prod(X.shape)) >= (1 << 20)  # small tables: nothing to hide
            cur = torch.cuda.current_stream(self.dev)
            if on_host:
                if self._side is None:
                    self._side = torch.cuda.Stream(device=self.dev)
                self._side.wait_stream(cur)
            st = self._side if on_host else cur
            with torch.cuda.stream(st):
                Xd = self.to_dev(X, torch.float32)
                n, d = Xd.shape
                check(_lib.lib().tst_bin_auto(self._h, self._ptr(Xd), C.c_int64(n), C.c_int32(d), C.c_int32(n_bins), None, None,
                                              C.c_void_p(st.cuda_stream)))
            Xd.record_stream(st)
            self.n, self.d, self._edges, self._n_edges, self._edges_lazy = n, d, None, None, True
            return self
        Xd = self.to_dev(X, torch.float32)
        n, d = Xd.shape
        if edges is None:
            Xh = X.detach().cpu().numpy() if isinstance(X, torch.Tensor) else np.asarray(X)
            edges, n_edges = binning.table_edges(Xh, n_bins)
        edges = np.ascontiguousarray(edges, np.float64)
        n_edges = np.ascontiguousarray(n_edges, np.int32)
        check(_lib.lib().tst_bin(self._h, self._ptr(Xd), C.c_int64(n), C.c_int32(d), edges.ctypes.data_as(C.c_void_p),
                                 n_edges.ctypes.data_as(C.c_void_p), self._stream()))
        torch.cuda.current_stream(self.dev).synchronize()  # edges/X staging buffers may be released
        self.n, self.d, self._edges, self._n_edges, self._edges_lazy = n, d, edges, n_edges, False
        return self

    def set_target(self, yq):
        """Continuous target of criterion "squared_error": uint16 fixed point, one value per row of the binned table."""
        torch = _torch()
        t = self.to_dev(np.ascontiguousarray(yq, dtype=np.uint16).view(np.int16), torch.int16)
        check(_lib.lib().tst_set_target(self._h, self._ptr(t), C.c_int64(t.numel()), self._stream()))
        torch.cuda.current_stream(self.dev).synchronize()
        return self

    def set_bins(self, bins_cm):
        torch = _torch()
        b = self.to_dev(bins_cm, torch.uint8)
        d, n = b.shape
        check(_lib.lib().tst_set_bins(self._h, self._ptr(b), C.c_int64(n), C.c_int32(d), self._stream()))
        torch.cuda.current_stream(self.dev).synchronize()
        self.n, self.d, self._edges, self._n_edges, self._edges_lazy = n, d, None, None, False
        return self

    def get_bins(self) -> np.ndarray:
        torch = _torch()
        out = torch.empty((self.d, self.n), dtype=torch.uint8, device=self.dev)
        check(_lib.lib().tst_get_bins(self._h, self._ptr(out), self._stream()))
        return out.cpu().numpy()

    # ---- folds / fit / votes
    def assign_folds(self, y_dev, active_dev, k, n_classes, group_dev=None, n_groups=1):
        torch = _torch()
        fold = torch.empty(self.n, dtype=torch.uint8, device=self.dev)
        check(_lib.lib().tst_assign_folds(self._h, self._ptr(y_dev), self._ptr(active_dev), self._ptr(group_dev),
                                          C.c_int32(n_groups), C.c_int32(k), C.c_int32(n_classes), self._ptr(fold),
                                          self._stream()))
        return fold

    def new_votes(self, n_classes):
        torch = _torch()
        return torch.zeros((self.n, n_classes), dtype=torch.int32, device=self.dev)  # bit pattern of uint32

    def fit_oof(self, y_dev, fold_dev, k, n_trees, n_classes, params: Params, seed=0, gid_base=0, t0=0, t1=None,
                group_dev=None, n_groups=1, votes=None):
        t1 = n_trees if t1 is None else t1
        check(_lib.lib().tst_fit_oof(self._h, self._ptr(y_dev), self._ptr(fold_dev), self._ptr(group_dev),
                                     C.c_int32(n_groups), C.c_int32(k), C.c_int32(n_trees), C.c_int32(n_classes),
                                     C.byref(params), C.c_uint64(seed), C.c_uint32(gid_base), C.c_int32(t0), C.c_int32(t1),
                                     self._ptr(votes), self._stream()))
        return votes

    def run_oof(self, y_dev, active_dev, k, n_trees, n_classes, params: Params, seed=0, group_dev=None, n_groups=1, want_s=True):
        """One C call for folds + forests + vote exchange + s(x) on device buffers (include/tst.h: tst_run_oof)
        -> (fold uint8 [n], votes int32 [n, C] (uint32 bit pattern), s float32 [n] | None)."""
        torch = _torch()
        fold = torch.empty(self.n, dtype=torch.uint8, device=self.dev)
        votes = torch.empty((self.n, n_classes), dtype=torch.int32, device=self.dev)
        s = torch.empty(self.n, dtype=torch.float32, device=self.dev) if want_s else None
        check(_lib.lib().tst_run_oof(self._h, self._ptr(y_dev), self._ptr(active_dev), self._ptr(group_dev), C.c_int32(n_groups),
                                     C.c_int32(k), C.c_int32(n_trees), C.c_int32(n_classes), C.byref(params), C.c_uint64(seed),
                                     self._ptr(fold), self._ptr(votes), self._ptr(s), self._stream()))
        return fold, votes, s

    def allreduce_votes(self, votes):
        """The path's one exchange step: NCCL all-gather of the per-rank votes, then a rank-ordered integer sum."""
        if self.world == 1:
            return votes
        if self.lib_nccl:  # in place, inside libtst: ncclAllGather + k_reduce_votes
            check(_lib.lib().tst_allreduce_votes(self._h, self._ptr(votes), C.c_int64(votes.numel()), self._stream()))
            return votes
        torch = _torch()
        import torch.distributed as dist

        gathered = torch.empty((self.world,) + tuple(votes.shape), dtype=votes.dtype, device=self.dev)
        dist.all_gather_into_tensor(gathered, votes)
        out = torch.empty_like(votes)
        check(_lib.lib().tst_reduce_votes(self._h, self._ptr(gathered), C.c_int32(self.world), C.c_int64(votes.numel()),
                                          self._ptr(out), self._stream()))
        return out

    def finalize(self, votes, y_dev, fold_dev, k, n_trees, n_classes, group_dev=None, n_groups=1):
        torch = _torch()
        s = torch.empty(self.n, dtype=torch.float32, device=self.dev)
        correct = torch.empty(n_groups * k, dtype=torch.int32, device=self.dev)
        total = torch.empty(n_groups * k, dtype=torch.int32, device=self.dev)
        check(_lib.lib().tst_finalize(self._h, self._ptr(votes), self._ptr(y_dev), self._ptr(fold_dev), self._ptr(group_dev),
                                      C.c_int32(n_groups), C.c_int32(k), C.c_int32(n_trees), C.c_int32(n_classes),
                                      self._ptr(s), self._ptr(correct), self._ptr(total), self._stream()))
        return s, correct, total

    def ibp_select(self, votes, y_dev, fold_dev, k, n_classes, vmin, batch, active_dev) -> int:
        removed = C.c_int64(0)
        check(_lib.lib().tst_ibp_select(self._h, self._ptr(votes), self._ptr(y_dev), self._ptr(fold_dev), C.c_int32(k),
                                        C.c_int32(n_classes), C.c_uint64(vmin), C.c_int64(batch), self._ptr(active_dev),
                                        C.byref(removed), self._stream()))
        return int(removed.value)

    def ibp(self, y_dev, active_dev, k, n_trees, n_classes, params, seed, rounds, tau, batch, group_dev=None, n_groups=1,
            want_s=True):
        torch = _torch()
        removed = (C.c_int64 * rounds)()
        s = torch.empty((rounds, self.n), dtype=torch.float32, device=self.dev) if want_s else None
        check(_lib.lib().tst_ibp(self._h, self._ptr(y_dev), self._ptr(active_dev), self._ptr(group_dev), C.c_int32(n_groups),
                                 C.c_int32(k), C.c_int32(n_trees), C.c_int32(n_classes), C.byref(params), C.c_uint64(seed),
                                 C.c_int32(rounds), C.c_double(tau), C.c_int64(batch), removed, self._ptr(s), self._stream()))
        return np.array(list(removed), dtype=np.int64), s

    # ---- kept forest
    def predict_votes(self, X, n_classes):
        torch = _torch()
        Xd = self.to_dev(X, torch.float32)
        votes = torch.zeros((Xd.shape[0], n_classes), dtype=torch.int32, device=self.dev)
        check(_lib.lib().tst_predict(self._h, self._ptr(Xd), C.c_int64(Xd.shape[0]), self._ptr(votes), self._stream()))
        return votes

    def export_tree(self, part, t, n_classes):
        n_nodes = C.c_int64(0)
        _lib.lib().tst_export_tree(self._h, C.c_int32(part), C.c_int32(t), None, None, C.c_int64(0), C.byref(n_nodes))
        if n_nodes.value <= 0:
            raise TstError(-4, _lib.lib().tst_last_error().decode())
        nodes = np.zeros(n_nodes.value, dtype=NODE_DTYPE)
        counts = np.zeros((n_nodes.value, n_classes), dtype=np.uint32)
        check(_lib.lib().tst_export_tree(self._h, C.c_int32(part), C.c_int32(t), nodes.ctypes.data_as(C.c_void_p),
                                         counts.ctypes.data_as(C.c_void_p), C.c_int64(n_nodes.value), C.byref(n_nodes)))
        return nodes, counts

    def importance_sums(self, raw: bool = False):
        """Integer feature-importance accumulator of the trees THIS rank built in the last fit (include/tst.h:
        tst_feature_importances): -> (sums uint64 [d], trees counted[, raw Q24 uint64 [trees, d]])."""
        sums = np.zeros(self.d, dtype=np.uint64)
        n_counted = C.c_int64(0)
        raw_arr, cap = None, 0
        if raw:
            cap = int(self.stats()["trees"])
            raw_arr = np.zeros((max(cap, 1), self.d), dtype=np.uint64)
        check(_lib.lib().tst_feature_importances(self._h, sums.ctypes.data_as(C.c_void_p), C.byref(n_counted),
                                                 None if raw_arr is None else raw_arr.ctypes.data_as(C.c_void_p), C.c_int64(cap)))
        return (sums, int(n_counted.value), raw_arr[:cap]) if raw else (sums, int(n_counted.value))

    def feature_importances(self) -> np.ndarray:
        """Mean decrease in impurity per feature over all trees of the last fit (all ranks), normalised to sum 1
        (SK/ensemble/_forest.py feature_importances_): accumulated on the device while the trees grew."""
        sums, _ = self.importance_sums()
        if self.world > 1:
            torch = _torch()
            import torch.distributed as dist

            t = torch.from_numpy(sums.view(np.int64).copy()).to(self.dev)
            dist.all_reduce(t)  # integer sums: any order
            sums = t.cpu().numpy().view(np.uint64)
        s = sums.astype(np.float64)
        tot = s.sum()
        return s / tot if tot > 0 else s

    def stats(self, reset=False) -> dict:
        st = Stats()
        check(_lib.lib().tst_get_stats(self._h, C.byref(st)))
        if reset:
            check(_lib.lib().tst_reset_stats(self._h))
        return st.as_dict()


NODE_DTYPE = np.dtype([("left", "<i4"), ("right", "<i4"), ("feature", "<i4"), ("bin", "<i4"),
                       ("begin", "<u4"), ("end", "<u4"), ("depth", "<u4"), ("n_weighted", "<u4")])

_engines: dict = {}


def get_engine(device=None) -> Engine:
    torch = _torch()
    if not torch.cuda.is_available():
        raise TstError(-2, "no CUDA device: the TsT engine has no CPU fallback")
    rank, world = dist_info()
    device = torch.cuda.current_device() if device is None else device
    key = (device, rank, world)
    if key not in _engines:
        _engines[key] = Engine(device, rank, world)
    return _engines[key]


# ------------------------------------------------------------------------------------------------- run_tst
@dataclass
class TsTResult:
    s: np.ndarray               # [n] float32 out-of-fold probability of the true class (bias score); -1 if never held out
    oof_votes: np.ndarray       # [n, C] uint32 Q0.16 vote sums over the fold's n_trees
    fold_id: np.ndarray         # [n] uint8
    fold_accuracy: np.ndarray   # [n_groups * k] blind accuracy per (group, fold)
    classes: np.ndarray
    stats: dict = field(default_factory=dict)
    feature_importances: np.ndarray | None = None  # [d] mean decrease in Gini impurity over the k x n_trees out-of-fold forests


def run_tst(X, y, k: int = 5, n_trees: int = 100, seed: int = 0, *, max_depth=None, min_samples_split=2,
            min_samples_leaf=1, max_features="sqrt", n_bins=256, group=None, fold=None, engine: Engine | None = None,
            trees_in_flight=0, tree_range=None, prebinned=False) -> TsTResult:
    """k-fold x n_trees RF diagnostic.  X, y are host arrays (or device tensors); every copy is inside this call."""
    torch = _torch()
    eng = engine or get_engine()
    if prebinned:
        eng.set_bins(X)
    else:
        eng.bin(X, n_bins=n_bins)  # large host tables: H2D + binning run on a side stream, under the label encoding below
    classes, yi = encode_labels(y)
    n_classes = int(classes.shape[0])
    y_dev = eng.to_dev(yi, torch.uint8)
    g_dev, n_groups = None, 1
    if group is not None:
        gi = np.asarray(group, dtype=np.int32)
        n_groups = int(gi.max()) + 1
        g_dev = eng.to_dev(gi, torch.int32)
    if fold is None:
        active = torch.ones(eng.n, dtype=torch.uint8, device=eng.dev)
        fold_dev = eng.assign_folds(y_dev, active, k, n_classes, g_dev, n_groups)
    else:
        fold_dev = eng.to_dev(np.asarray(fold, dtype=np.uint8), torch.uint8)
    prm = make_params(eng.d, max_depth, min_samples_split, min_samples_leaf, max_features, trees_in_flight)
    votes = eng.new_votes(n_classes)
    t0, t1 = (0, n_trees) if tree_range is None else tree_range
    eng.stats(reset=True)
    eng.fit_oof(y_dev, fold_dev, k, n_trees, n_classes, prm, seed=seed, t0=t0, t1=t1, group_dev=g_dev, n_groups=n_groups,
                votes=votes)
    votes = eng.allreduce_votes(votes)
    s, correct, total = eng.finalize(votes, y_dev, fold_dev, k, n_trees, n_classes, g_dev, n_groups)
    s_h = s.cpu().numpy()
    votes_h = votes.cpu().numpy().view(np.uint32)
    tot = total.cpu().numpy().astype(np.float64)
    acc = np.divide(correct.cpu().numpy().astype(np.float64), tot, out=np.full_like(tot, np.nan), where=tot > 0)
    return TsTResult(s_h, votes_h, fold_dev.cpu().numpy(), acc, classes, eng.stats(), eng.feature_importances())


# ------------------------------------------------------------------------------------------------- numeric answers
# the squared-error criterion (SK/tree/_criterion.pyx:707-1000) for question types whose answer is a
# number.  criterion="squared_error": the target is mapped to 16-bit fixed point on the host; the kernels build (weight, sum y)
# histograms with exact integer sums, leaves hold (N, sum y), every tree votes its leaf mean in Q16.16.
# criterion="squared_error_levels" (round 1) histograms <= 32 ordinal levels like classes; kept for comparison.
def quantise_target16(y):
    """-> (yq uint16 [n], lo, span): yq = round((y - lo) / span * 65535); y ~ lo + yq / 65535 * span."""
    y = np.asarray(y, dtype=np.float64)
    if y.size == 0 or not np.all(np.isfinite(y)):
        raise ValueError("numeric target must be non-empty and finite")
    lo, hi = float(y.min()), float(y.max())
    span = hi - lo if hi > lo else 1.0
    return np.round((y - lo) / span * 65535.0).astype(np.uint16), lo, span


def votes16_to_value(votes, n_trees: int, lo: float, span: float) -> np.ndarray:
    """[n, 2] vote words (low / high 16-bit halves of the Q16.16 leaf means, summed over trees) -> mean prediction on the
    target's scale; NaN for rows no tree voted on."""
    v = np.asarray(votes).astype(np.float64)
    tot = v[:, 1] * 65536.0 + v[:, 0]
    out = lo + tot / (float(n_trees) * 65536.0) / 65535.0 * span
    return np.where((v[:, 0] == 0) & (v[:, 1] == 0), np.nan, out)


def quantise_target(y, n_levels: int = 32):
    """(criterion "squared_error_levels") -> (levels uint8 [n], lo, step): level = floor((y - lo) / step), step = (max - min) / n_levels."""
    if not 2 <= n_levels <= 32:
        raise ValueError("n_levels must be in [2, 32]")
    y = np.asarray(y, dtype=np.float64)
    if y.size == 0 or not np.all(np.isfinite(y)):
        raise ValueError("numeric target must be non-empty and finite")
    lo, hi = float(y.min()), float(y.max())
    step = (hi - lo) / n_levels if hi > lo else 1.0
    lev = np.minimum(np.floor((y - lo) / step), n_levels - 1).astype(np.uint8)
    return lev, lo, step


def votes_to_value(votes, lo: float, step: float) -> np.ndarray:
    """(criterion "squared_error_levels") vote-weighted mean level -> target scale; NaN without votes."""
    v = np.asarray(votes).astype(np.float64)
    tot = v.sum(axis=1)
    mean_level = np.divide(v @ np.arange(v.shape[1], dtype=np.float64), tot, out=np.full(v.shape[0], np.nan), where=tot > 0)
    return lo + (mean_level + 0.5) * step


@dataclass
class TsTRegressionResult:
    y_pred: np.ndarray          # [n] float64 out-of-fold blind prediction of the numeric answer (NaN if never held out)
    s: np.ndarray               # [n] float64 bias score 1 - |y_pred - y| / (max(y) - min(y)) in [0, 1]
    oof_votes: np.ndarray       # [n, 2] uint32 vote words (criterion "squared_error") or [n, n_levels] Q0.16 level votes
    fold_id: np.ndarray
    levels: np.ndarray          # [n] uint8 stratification levels of the fold assignment
    lo: float = 0.0
    step: float = 1.0           # span of the target ("squared_error") or width of one level ("squared_error_levels")
    stats: dict = field(default_factory=dict)
    feature_importances: np.ndarray | None = None


def run_tst_regression(X, y, k: int = 5, n_trees: int = 100, seed: int = 0, *, criterion: str = "squared_error", n_levels: int = 32,
                       max_depth=None, min_samples_split=2, min_samples_leaf=1, max_features=1.0, n_bins=256, group=None,
                       engine: Engine | None = None, trees_in_flight=0) -> TsTRegressionResult:
    """k-fold x n_trees diagnostic for numeric answers (max_features default 1.0 as in RandomForestRegressor).
    Folds are stratified by `n_levels` coarse levels of the target."""
    torch = _torch()
    eng = engine or get_engine()
    if criterion not in ("squared_error", "squared_error_levels"):
        raise ValueError("criterion must be 'squared_error' or 'squared_error_levels'")
    eng.bin(X, n_bins=n_bins)
    g_dev, n_groups = None, 1
    if group is not None:
        gi = np.asarray(group, dtype=np.int32)
        n_groups = int(gi.max()) + 1
        g_dev = eng.to_dev(gi, torch.int32)
    active = torch.ones(eng.n, dtype=torch.uint8, device=eng.dev)
    y64 = np.asarray(y, dtype=np.float64)
    if criterion == "squared_error":
        yq, lo, span = quantise_target16(y64)
        lev = np.minimum((yq.astype(np.uint32) * n_levels) >> 16, n_levels - 1).astype(np.uint8)
        eng.set_target(yq)
        n_classes, step = 2, span
    else:
        lev, lo, step = quantise_target(y64, n_levels)
        n_classes, span = n_levels, n_levels * step
    y_dev = eng.to_dev(lev, torch.uint8)
    fold_dev = eng.assign_folds(y_dev, active, k, n_levels, g_dev, n_groups)  # folds stratified by level
    prm = make_params(eng.d, max_depth, min_samples_split, min_samples_leaf, max_features, trees_in_flight, criterion=criterion)
    votes = eng.new_votes(n_classes)
    eng.stats(reset=True)
    eng.fit_oof(y_dev, fold_dev, k, n_trees, n_classes, prm, seed=seed, group_dev=g_dev, n_groups=n_groups, votes=votes)
    votes_h = eng.allreduce_votes(votes).cpu().numpy().view(np.uint32)
    y_pred = votes16_to_value(votes_h, n_trees, lo, span) if criterion == "squared_error" else votes_to_value(votes_h, lo, step)
    s = 1.0 - np.minimum(1.0, np.abs(y_pred - y64) / span)
    return TsTRegressionResult(y_pred, s, votes_h, fold_dev.cpu().numpy(), lev, lo, step, eng.stats(), eng.feature_importances())


# ------------------------------------------------------------------------------------------------- IBP
@dataclass
class IBPResult:
    keep_mask: np.ndarray          # [n] bool, rows surviving all rounds
    removed_per_round: np.ndarray  # [rounds] int64
    s_per_round: np.ndarray | None  # [rounds, n] float32 (-1 where the row was not scored in that round)
    stats: dict = field(default_factory=dict)


def ibp_vmin(tau: float, n_trees: int) -> int:
    """Integer form of s >= tau: votes[y] >= ceil(tau * n_trees * 65536)."""
    return int(np.ceil(np.float64(tau) * np.float64(n_trees) * 65536.0))


def ibp(X, y, rounds: int = 10, k: int = 5, n_trees: int = 200, tau: float = 0.5, batch: int | None = None, seed: int = 0, *,
        max_depth=None, min_samples_split=2, min_samples_leaf=1, max_features="sqrt", n_bins=256, group=None,
        engine: Engine | None = None, want_s=True, prebinned=False) -> IBPResult:
    """Iterative Bias Pruning; the table, labels, active mask, votes and s(x) never leave the device between rounds."""
    torch = _torch()
    eng = engine or get_engine()
    classes, yi = encode_labels(y)
    n_classes = int(classes.shape[0])
    if prebinned:
        eng.set_bins(X)
    else:
        eng.bin(X, n_bins=n_bins)
    n = eng.n
    batch = n if batch is None else int(batch)
    y_dev = eng.to_dev(yi, torch.uint8)
    g_dev, n_groups = None, 1
    if group is not None:
        gi = np.asarray(group, dtype=np.int32)
        n_groups = int(gi.max()) + 1
        g_dev = eng.to_dev(gi, torch.int32)
    active = torch.ones(n, dtype=torch.uint8, device=eng.dev)
    prm = make_params(eng.d, max_depth, min_samples_split, min_samples_leaf, max_features)
    eng.stats(reset=True)
    if eng.world == 1 or eng.lib_nccl:  # the whole loop runs inside libtst (every rank calls it)
        removed, s = eng.ibp(y_dev, active, k, n_trees, n_classes, prm, seed, rounds, tau, batch, g_dev, n_groups, want_s)
    else:  # the per-round vote all-gather is issued from the host; data stays on the device
        removed = np.zeros(rounds, np.int64)
        s = torch.full((rounds, n), -1.0, dtype=torch.float32, device=eng.dev) if want_s else None
        vmin = ibp_vmin(tau, n_trees)
        per_round = n_groups * k * n_trees
        for r in range(rounds):
            fold_dev = eng.assign_folds(y_dev, active, k, n_classes, g_dev, n_groups)
            votes = eng.new_votes(n_classes)
            eng.fit_oof(y_dev, fold_dev, k, n_trees, n_classes, prm, seed=seed, gid_base=r * per_round, group_dev=g_dev,
                        n_groups=n_groups, votes=votes)
            votes = eng.allreduce_votes(votes)
            if want_s:
                s[r] = eng.finalize(votes, y_dev, fold_dev, k, n_trees, n_classes, g_dev, n_groups)[0]
            removed[r] = eng.ibp_select(votes, y_dev, fold_dev, k, n_classes, vmin, batch, active)
            if removed[r] == 0:
                break
    return IBPResult(active.cpu().numpy().astype(bool), removed, None if s is None else s.cpu().numpy(), eng.stats())


# ------------------------------------------------------------------------------------------------- TsTForest
class TsTForest:
    """sklearn-like classifier over the same kernels (argument names: SK/ensemble/_forest.py:1515-1537)."""

    def __init__(self, n_estimators=100, max_depth=None, min_samples_split=2, min_samples_leaf=1, max_features="sqrt",
                 n_bins=256, random_state=0, engine: Engine | None = None, criterion="gini", n_levels=32):
        self.n_estimators, self.max_depth = n_estimators, max_depth
        self.min_samples_split, self.min_samples_leaf = min_samples_split, min_samples_leaf
        self.max_features, self.n_bins, self.random_state = max_features, n_bins, random_state
        self.criterion, self.n_levels = criterion, n_levels  # "squared_error" / "squared_error_levels": y is numeric, predict() returns numbers
        self._engine = engine

    def fit(self, X, y):
        torch = _torch()
        # a kept forest lives in its engine: use a private one unless the caller shares one
        self._eng = self._engine or Engine(torch.cuda.current_device() if torch.cuda.is_available() else 0)
        yq = None
        if self.criterion == "squared_error":
            yq, self._lo, self._step = quantise_target16(y)
            yi = np.zeros(len(yq), dtype=np.uint8)  # labels only stratify folds; a plain fit has none
            self.classes_ = np.arange(2)
        elif self.criterion == "squared_error_levels":
            yi, self._lo, self._step = quantise_target(y, self.n_levels)
            self.classes_ = np.arange(self.n_levels)
        else:
            self.classes_, yi = encode_labels(y)
        self.n_classes_ = int(self.classes_.shape[0])
        self._eng.bin(X, n_bins=self.n_bins)
        if yq is not None:
            self._eng.set_target(yq)
        self.n_features_in_ = self._eng.d
        y_dev = self._eng.to_dev(yi, torch.uint8)
        fold = torch.full((self._eng.n,), FOLD_TRAIN_ONLY, dtype=torch.uint8, device=self._eng.dev)
        prm = make_params(self._eng.d, self.max_depth, self.min_samples_split, self.min_samples_leaf, self.max_features,
                          keep_forest=True, criterion=self.criterion)
        self._eng.fit_oof(y_dev, fold, 1, self.n_estimators, self.n_classes_, prm, seed=int(self.random_state or 0))
        self._importances = self._eng.feature_importances()
        return self

    def predict_votes(self, X) -> np.ndarray:
        return self._eng.predict_votes(X, self.n_classes_).cpu().numpy().view(np.uint32)

    def predict_proba(self, X) -> np.ndarray:
        v = self.predict_votes(X).astype(np.float64)
        return v / v.sum(axis=1, keepdims=True)

    def predict(self, X):
        if self.criterion == "squared_error":
            return votes16_to_value(self.predict_votes(X), self.n_estimators, self._lo, self._step)
        if self.criterion == "squared_error_levels":
            return votes_to_value(self.predict_votes(X), self._lo, self._step)
        return self.classes_[np.argmax(self.predict_votes(X), axis=1)]

    def export_tree(self, t: int):
        return self._eng.export_tree(0, t, self.n_classes_)

    @property
    def feature_importances_(self) -> np.ndarray:
        """Gini importances (mean over trees of the per-tree normalised decrease in impurity, normalised) -- the
        "interpretable auditing" output of REF/README.md:35 (SK/ensemble/_forest.py feature_importances_).  Accumulated on
        the device while the trees grew (integer fixed point: independent of batching and rank count)."""
        return self._importances

    def feature_importances_from_trees(self) -> np.ndarray:
        """The same quantity recomputed on the host from the exported trees (cross-check of the device accumulation)."""
        from .features import gini_importances

        imp = np.zeros(self.n_features_in_, dtype=np.float64)
        for t in range(self.n_estimators):
            nodes, counts = self.export_tree(t)
            imp += gini_importances(nodes, counts, self.n_features_in_)
        imp /= self.n_estimators
        s = imp.sum()
        return imp / s if s > 0 else imp

    def explain(self, X, y) -> tuple:
        """Mean per-feature attribution of P(true class) over the forest's trees for the rows of X
        (features.path_contributions on every exported tree): -> (prior [n], contributions [n, d]) with
        prior + contributions.sum(1) == mean over trees of the leaf probability of the true class."""
        from . import binning
        from .features import path_contributions

        bins_rows = binning.bin_rows(np.asarray(X, dtype=np.float32), self._eng.edges, self._eng.n_edges)
        idx = {c: i for i, c in enumerate(self.classes_.tolist())}
        cls = np.array([idx[v] for v in np.asarray(y).tolist()], dtype=np.int64)
        prior = np.zeros(bins_rows.shape[0])
        contrib = np.zeros((bins_rows.shape[0], self.n_features_in_))
        for t in range(self.n_estimators):
            nodes, counts = self.export_tree(t)
            p, c = path_contributions(nodes, counts, bins_rows, cls, self.n_features_in_)
            prior += p
            contrib += c
        return prior / self.n_estimators, contrib / self.n_estimators

    def sklearn_tree(self, t: int) -> dict:
        """Tree `t` as sklearn-style node arrays with raw-space thresholds (SK/tree/_tree.pxd:15-25)."""
        from .features import to_sklearn_arrays

        nodes, counts = self.export_tree(t)
        return to_sklearn_arrays(nodes, counts, self._eng.edges, self._eng.n_edges)

    def export(self):
        """The whole kept forest on the host (interop.ExportedForest): .save(path), .to_sklearn() -> a genuine
        sklearn RandomForestClassifier, .predict_votes(X)."""
        from .interop import ExportedForest

        trees = [self.export_tree(t) for t in range(self.n_estimators)]
        if self.criterion == "squared_error_levels":
            raise ValueError("export: use criterion='squared_error' (the level-based criterion keeps level counts, not target sums)")
        reg = self.criterion == "squared_error"
        return ExportedForest([n for n, _ in trees], [c for _, c in trees], self._eng.edges.copy(), self._eng.n_edges.copy(),
                              np.asarray(self.classes_), int(self.n_features_in_), self.criterion,
                              float(self._lo) if reg else 0.0, float(self._step) if reg else 1.0)

    def save(self, path: str) -> None:
        self.export().save(path)

    def to_sklearn(self):
        return self.export().to_sklearn()
