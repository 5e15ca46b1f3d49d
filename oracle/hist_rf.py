"""ctypes wrapper around oracle/hist_rf.c plus the numpy restatement of the bin-edge rule.

TEST INFRASTRUCTURE ONLY (see hist_rf.c header): imported by tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs, never by test-set-training_b200/.
"parity unpinned": the reference ships no golden vectors; pins are listed in hist_rf.c.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

FOLD_TRAIN_ONLY = 254
FOLD_INACTIVE = 255
MAX_EDGES = 255


class OrcParams(C.Structure):
    _fields_ = [("max_depth", C.c_int32), ("min_samples_split", C.c_int32),
                ("min_samples_leaf", C.c_int32), ("max_features", C.c_int32), ("criterion", C.c_int32)]


class OrcStats(C.Structure):
    _fields_ = [("row_visits", C.c_uint64), ("nodes", C.c_uint64), ("leaves", C.c_uint64),
                ("searched", C.c_uint64), ("fallback", C.c_uint64), ("inbag_rows", C.c_uint64),
                ("trees", C.c_uint64), ("oof_steps", C.c_uint64), ("max_depth", C.c_uint32),
                ("pad", C.c_uint32)]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_ if k != "pad"}


NODE_DTYPE = np.dtype([("left", "<i4"), ("right", "<i4"), ("feature", "<i4"), ("bin", "<i4"),
                       ("begin", "<u4"), ("end", "<u4"), ("depth", "<u4"), ("n_weighted", "<u4")])


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liborc.so")
    src = os.path.join(_HERE, "hist_rf.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liborc.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_fit_oof.restype = C.c_int
        L.orc_export_tree.restype = C.c_int
        L.orc_tree_size.restype = C.c_int64
        L.orc_ibp_select.restype = C.c_int64
        _LIB = L
    return _LIB


def _p(a, t=C.c_void_p):
    return None if a is None else a.ctypes.data_as(t)


def philox(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    lib().orc_philox(_p(c), _p(k), _p(out))
    return out


# ------------------------------------------------------------------ bin edges
def find_edges_column(col: np.ndarray, n_bins: int = 256) -> np.ndarray:
    """Restates SK/ensemble/_hist_gradient_boosting/binning.py:31-111 with exact integer ranks.

    max_bins = n_bins - 1 non-missing bins.  NaNs ignored.  <= max_bins distinct values ->
    midpoints of consecutive distinct values; otherwise the max_bins-1 interior quantiles by the
    'averaged_inverted_cdf' rule written in integers: t = i*n, j = t // max_bins; the edge is a[j]
    if t % max_bins else (a[j-1] + a[j]) / 2; duplicates removed.  All in float64.
    """
    max_bins = n_bins - 1
    a = np.asarray(col, dtype=np.float64)
    a = np.sort(a[~np.isnan(a)])
    n = a.shape[0]
    if n == 0:
        return np.zeros(0, dtype=np.float64)
    distinct = a[np.concatenate(([True], a[1:] != a[:-1]))]
    if distinct.shape[0] == 1:
        return np.zeros(0, dtype=np.float64)
    if distinct.shape[0] <= max_bins:
        e = (distinct[:-1] + distinct[1:]) * 0.5
    else:
        e = np.empty(max_bins - 1, dtype=np.float64)
        for i in range(1, max_bins):
            t = i * n
            j, rem = divmod(t, max_bins)
            e[i - 1] = a[j] if rem else (a[j - 1] + a[j]) * 0.5
        keep = np.concatenate(([True], e[1:] != e[:-1]))
        e = e[keep]
    return np.minimum(e, 1e300)


def find_edges(X: np.ndarray, n_bins: int = 256):
    X = np.asarray(X, dtype=np.float32)
    n, d = X.shape
    edges = np.zeros((d, MAX_EDGES), dtype=np.float64)
    n_edges = np.zeros(d, dtype=np.int32)
    for f in range(d):
        e = find_edges_column(X[:, f], n_bins)
        n_edges[f] = e.shape[0]
        edges[f, : e.shape[0]] = e
    return edges, n_edges


def bin_table(X: np.ndarray, edges: np.ndarray, n_edges: np.ndarray) -> np.ndarray:
    """-> uint8 [d, n] column-major bins."""
    X = np.ascontiguousarray(X, dtype=np.float32)
    n, d = X.shape
    edges = np.ascontiguousarray(edges, dtype=np.float64)
    n_edges = np.ascontiguousarray(n_edges, dtype=np.int32)
    out = np.zeros((d, n), dtype=np.uint8)
    rc = lib().orc_bin(_p(X), C.c_int64(n), C.c_int32(d), _p(edges), _p(n_edges), _p(out))
    if rc:
        raise RuntimeError(f"orc_bin rc={rc}")
    return out


def assign_folds(y, active, k, n_classes, group=None, n_groups=1) -> np.ndarray:
    y = np.ascontiguousarray(y, dtype=np.uint8)
    active = np.ascontiguousarray(active, dtype=np.uint8)
    g = None if group is None else np.ascontiguousarray(group, dtype=np.int32)
    fold = np.zeros(y.shape[0], dtype=np.uint8)
    rc = lib().orc_assign_folds(_p(y), _p(active), _p(g), C.c_int32(n_groups), C.c_int64(y.shape[0]),
                                C.c_int32(k), C.c_int32(n_classes), _p(fold))
    if rc:
        raise RuntimeError(f"orc_assign_folds rc={rc}")
    return fold


def resolve_max_features(max_features, d: int) -> int:
    """SK/tree/_classes.py:330-345 ("sqrt" -> max(1, int(sqrt(d))))."""
    if max_features in (None, "all"):
        return d
    if max_features == "sqrt":
        return max(1, int(np.sqrt(d)))
    if max_features == "log2":
        return max(1, int(np.log2(d)))
    if isinstance(max_features, float):
        return max(1, int(max_features * d))
    return int(max_features)


CRITERIA = {"gini": 0, "squared_error_levels": 1, "squared_error": 2}


def make_params(d, max_depth=None, min_samples_split=2, min_samples_leaf=1, max_features="sqrt", criterion="gini"):
    return OrcParams(0 if max_depth is None else int(max_depth), int(min_samples_split),
                     int(min_samples_leaf), resolve_max_features(max_features, d), CRITERIA[criterion])


@dataclass
class Tree:
    nodes: np.ndarray   # NODE_DTYPE, canonical (depth, begin) order
    counts: np.ndarray  # [n_nodes, C] uint32


class Forest:
    def __init__(self, handle, n_parts, n_trees, n_classes):
        self._h = handle
        self.n_parts, self.n_trees, self.n_classes = n_parts, n_trees, n_classes

    def tree(self, part: int, t: int) -> Tree:
        L = lib()
        m = L.orc_tree_size(self._h, C.c_int32(part), C.c_int32(self.n_trees), C.c_int32(t))
        if m <= 0:
            raise KeyError((part, t))
        nodes = np.zeros(m, dtype=NODE_DTYPE)
        counts = np.zeros((m, self.n_classes), dtype=np.uint32)
        rc = L.orc_export_tree(self._h, C.c_int32(part), C.c_int32(self.n_trees), C.c_int32(t),
                               _p(nodes), _p(counts))
        if rc:
            raise RuntimeError(f"orc_export_tree rc={rc}")
        return Tree(nodes, counts)

    def predict_votes(self, bins_cm: np.ndarray) -> np.ndarray:
        bins_cm = np.ascontiguousarray(bins_cm, dtype=np.uint8)
        n = bins_cm.shape[1]
        votes = np.zeros((n, self.n_classes), dtype=np.uint32)
        lib().orc_predict(self._h, _p(bins_cm), C.c_int64(n), _p(votes))
        return votes

    def __del__(self):
        if self._h:
            lib().orc_free_forest(self._h)
            self._h = None


def fit_oof(bins_cm, y, fold, k, n_trees, n_classes, params: OrcParams, seed=0, gid_base=0, t0=0, t1=None,
            rank=0, world=1, group=None, n_groups=1, votes=None, keep=False, nthreads=0, importances=None):
    """-> (votes [n, C] uint32, stats dict, Forest | None).  importances: optional uint64 [d + 1] accumulator (zeroed by the
    caller): sum over the trees built of floor(imp_t[f] * 2^32 / sum_f imp_t[f]); [d] counts the trees that split at all."""
    bins_cm = np.ascontiguousarray(bins_cm, dtype=np.uint8)
    d, n = bins_cm.shape
    y = np.ascontiguousarray(y, dtype=np.uint8)
    fold = np.ascontiguousarray(fold, dtype=np.uint8)
    g = None if group is None else np.ascontiguousarray(group, dtype=np.int32)
    if votes is None:
        votes = np.zeros((n, n_classes), dtype=np.uint32)
    t1 = n_trees if t1 is None else t1
    st = OrcStats()
    fh = C.c_void_p()
    rc = lib().orc_fit_oof(_p(bins_cm), C.c_int64(n), C.c_int32(d), _p(y), _p(fold), _p(g), C.c_int32(n_groups),
                           C.c_int32(k), C.c_int32(n_trees), C.c_int32(n_classes), C.byref(params),
                           C.c_uint64(seed), C.c_uint32(gid_base), C.c_int32(t0), C.c_int32(t1),
                           C.c_int32(rank), C.c_int32(world), _p(votes), C.byref(st),
                           C.byref(fh) if keep else None, C.c_int32(nthreads), _p(importances))
    if rc:
        raise RuntimeError(f"orc_fit_oof rc={rc}")
    forest = Forest(fh, n_groups * k, n_trees, n_classes) if keep else None
    return votes, st.as_dict(), forest


def importances_from_sums(sums) -> np.ndarray:
    """Forest-level feature importances from the integer accumulator: mean over trees of the per-tree normalised importances,
    normalised again (SK/ensemble/_forest.py feature_importances_)."""
    s = np.asarray(sums[:-1], dtype=np.float64)
    tot = s.sum()
    return s / tot if tot > 0 else s


def set_target(yq):
    """Continuous target of criterion 'squared_error' (uint16 fixed point); keep the array alive while fitting."""
    global _TARGET
    _TARGET = None if yq is None else np.ascontiguousarray(yq, dtype=np.uint16)
    lib().orc_set_target(_p(_TARGET))


_TARGET = None


def finalize(votes, y, fold, k, n_trees, group=None, n_groups=1):
    """-> (s float32 [n], correct [parts], total [parts])."""
    votes = np.ascontiguousarray(votes, dtype=np.uint32)
    n, n_classes = votes.shape
    y = np.ascontiguousarray(y, dtype=np.uint8)
    fold = np.ascontiguousarray(fold, dtype=np.uint8)
    g = None if group is None else np.ascontiguousarray(group, dtype=np.int32)
    s = np.zeros(n, dtype=np.float32)
    correct = np.zeros(n_groups * k, dtype=np.uint32)
    total = np.zeros(n_groups * k, dtype=np.uint32)
    lib().orc_finalize(_p(votes), _p(y), _p(fold), _p(g), C.c_int64(n), C.c_int32(n_classes), C.c_int32(k),
                       C.c_int32(n_trees), _p(s), _p(correct), _p(total))
    return s, correct, total


def ibp_select(votes, y, fold, k, vmin, batch, active):
    votes = np.ascontiguousarray(votes, dtype=np.uint32)
    n, n_classes = votes.shape
    y = np.ascontiguousarray(y, dtype=np.uint8)
    fold = np.ascontiguousarray(fold, dtype=np.uint8)
    assert active.dtype == np.uint8 and active.flags.c_contiguous
    return int(lib().orc_ibp_select(_p(votes), _p(y), _p(fold), C.c_int64(n), C.c_int32(n_classes), C.c_int32(k),
                                    C.c_uint64(vmin), C.c_int64(batch), _p(active)))


def ibp(bins_cm, y, k, n_trees, n_classes, params, seed=0, rounds=10, tau=0.5, batch=None, active=None,
        group=None, n_groups=1, nthreads=0, want_s=True):
    """-> (keep_mask uint8 [n], removed_per_round int64 [rounds], s_per_round float32 [rounds, n] | None)."""
    bins_cm = np.ascontiguousarray(bins_cm, dtype=np.uint8)
    d, n = bins_cm.shape
    y = np.ascontiguousarray(y, dtype=np.uint8)
    act = np.ones(n, dtype=np.uint8) if active is None else np.array(active, dtype=np.uint8, copy=True)
    g = None if group is None else np.ascontiguousarray(group, dtype=np.int32)
    batch = n if batch is None else int(batch)
    removed = np.zeros(rounds, dtype=np.int64)
    s = np.zeros((rounds, n), dtype=np.float32) if want_s else None
    rc = lib().orc_ibp(_p(bins_cm), C.c_int64(n), C.c_int32(d), _p(y), _p(act), _p(g), C.c_int32(n_groups),
                       C.c_int32(k), C.c_int32(n_trees), C.c_int32(n_classes), C.byref(params), C.c_uint64(seed),
                       C.c_int32(rounds), C.c_double(tau), C.c_int64(batch), _p(removed), _p(s), C.c_int32(nthreads))
    if rc:
        raise RuntimeError(f"orc_ibp rc={rc}")
    return act, removed, s


def num_threads() -> int:
    return int(lib().orc_num_threads())
