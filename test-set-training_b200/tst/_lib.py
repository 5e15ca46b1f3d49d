"""ctypes binding of libtst.so (include/tst.h).  There is no CPU fallback: a missing or unloadable
library, or a machine without a CUDA device, raises instead of degrading."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
LIB_PATH = os.environ.get("TST_LIB") or os.path.join(_ROOT, "libtst.so")  # TST_LIB: A/B builds while tuning
CSRC = os.path.join(_ROOT, "csrc")

FOLD_TRAIN_ONLY = 254
FOLD_INACTIVE = 255
MAX_EDGES = 255
MAX_CLASSES = 32
MAX_FEATURES = 1024
MAX_ROWS = 0x7FFFFFF0

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared", "-ldl"]


class TstError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libtst error {code}: {msg}")
        self.code = code


class Params(C.Structure):
    _fields_ = [("max_depth", C.c_int32), ("min_samples_split", C.c_int32), ("min_samples_leaf", C.c_int32),
                ("max_features", C.c_int32), ("trees_in_flight", C.c_int32), ("keep_forest", C.c_int32),
                ("criterion", C.c_int32), ("reserved", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("row_visits_big", C.c_uint64), ("row_visits_small", C.c_uint64), ("nodes", C.c_uint64),
                ("leaves", C.c_uint64), ("searched", C.c_uint64), ("fallback", C.c_uint64),
                ("inbag_rows", C.c_uint64), ("trees", C.c_uint64), ("oof_steps", C.c_uint64),
                ("kernel_launches", C.c_uint64), ("levels", C.c_uint64), ("max_depth", C.c_uint32),
                ("pad", C.c_uint32), ("ms_grow_big", C.c_double), ("ms_grow_small", C.c_double),
                ("ms_prep", C.c_double), ("ms_infer", C.c_double), ("launches_grow_big", C.c_uint64),
                ("launches_grow_small", C.c_uint64), ("ms_grow_mid", C.c_double), ("launches_grow_mid", C.c_uint64), ("row_visits_mid", C.c_uint64), ("ms_grow_tiny", C.c_double), ("launches_grow_tiny", C.c_uint64), ("launches_root", C.c_uint64),
                ("row_visits_sub", C.c_uint64), ("ms_grow_sub", C.c_double), ("launches_grow_sub", C.c_uint64),
                ("row_visits_wide", C.c_uint64), ("ms_grow_wide", C.c_double), ("launches_grow_wide", C.c_uint64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "pad"}


class Node(C.Structure):
    _fields_ = [("left", C.c_int32), ("right", C.c_int32), ("feature", C.c_int32), ("bin", C.c_int32),
                ("begin", C.c_uint32), ("end", C.c_uint32), ("depth", C.c_uint32), ("n_weighted", C.c_uint32)]


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))]


def build(force: bool = False, verbose: bool = False) -> str:
    """nvcc cross-compiles for sm_100a without a GPU; the .so stays in-tree so it travels with the repo."""
    srcs = sources() + [os.path.join(os.path.dirname(_ROOT), "include", "tst.h")]
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in srcs):
        return LIB_PATH
    cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH, os.path.join(CSRC, "tst_api.cu")]
    subprocess.check_call(cmd)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise TstError(-2, f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(libtst has no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.tst_last_error.restype = C.c_char_p
        for name in ("tst_create", "tst_destroy", "tst_bin", "tst_bin_auto", "tst_set_bins", "tst_get_bins", "tst_assign_folds",
                     "tst_fit_oof", "tst_reduce_votes", "tst_allreduce_votes", "tst_nccl_unique_id", "tst_finalize", "tst_ibp_select", "tst_ibp", "tst_predict",
                     "tst_export_tree", "tst_get_stats", "tst_reset_stats", "tst_abi_version", "tst_set_option", "tst_feature_importances", "tst_set_target", "tst_get_edges", "tst_run_oof"):
            getattr(L, name).restype = C.c_int
        _lib = L
    return _lib


EXPORTS = ["tst_abi_version", "tst_last_error", "tst_create", "tst_destroy", "tst_bin", "tst_bin_auto", "tst_set_bins", "tst_get_bins",
           "tst_assign_folds", "tst_fit_oof", "tst_reduce_votes", "tst_allreduce_votes", "tst_nccl_unique_id", "tst_finalize", "tst_ibp_select", "tst_ibp",
           "tst_predict", "tst_export_tree", "tst_get_stats", "tst_reset_stats", "tst_set_option", "tst_feature_importances", "tst_set_target", "tst_get_edges", "tst_run_oof"]


def check(rc: int):
    if rc != 0:
        raise TstError(rc, lib().tst_last_error().decode("utf-8", "replace"))
