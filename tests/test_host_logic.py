"""CPU tests of the host side: C-ABI surface, loud failure without CUDA, host helpers, rank sharding."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "tst.h")).read()
    return sorted(set(re.findall(r"^(?:int|const char\*)\s+(tst_\w+)\s*\(", src, flags=re.M)))


def test_library_exports_every_declared_symbol():
    from tst import _lib

    _lib.build()
    L = ctypes.CDLL(_lib.LIB_PATH)
    names = header_functions()
    assert len(names) >= 15 and set(names) == set(_lib.EXPORTS)
    for n in names:
        assert hasattr(L, n), n
    L.tst_abi_version.restype = ctypes.c_int
    assert L.tst_abi_version() == 3


def test_ctypes_structs_match_the_header(tmp_path):
    """The ctypes mirrors in tst/_lib.py must have the size and field offsets gcc gives include/tst.h."""
    import subprocess

    from tst import _lib

    prog = tmp_path / "layout.c"
    fields_p = [n for n, _ in _lib.Params._fields_]
    fields_s = [n for n, _ in _lib.Stats._fields_]
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "tst.h"', 'int main(void) {',
             'printf("%zu %zu %zu\\n", sizeof(tst_params), sizeof(tst_stats), sizeof(tst_node));']
    lines += [f'printf("p {f} %zu\\n", offsetof(tst_params, {f}));' for f in fields_p]
    lines += [f'printf("s {f} %zu\\n", offsetof(tst_stats, {f}));' for f in fields_s]
    lines += ['return 0; }']
    prog.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(prog)])
    out = subprocess.check_output([str(exe)], text=True).splitlines()
    sp, ss, sn = map(int, out[0].split())
    assert sp == ctypes.sizeof(_lib.Params) and ss == ctypes.sizeof(_lib.Stats)
    for l in out[1:]:
        kind, name, off = l.split()
        cls = _lib.Params if kind == "p" else _lib.Stats
        assert getattr(cls, name).offset == int(off), (kind, name)
    from oracle import hist_rf as H

    assert sn == H.NODE_DTYPE.itemsize  # tst_export_tree fills the oracle's node layout


def test_no_cpu_fallback():
    """Without a CUDA device every entry point must fail loudly -- never route through the oracle."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import tst
    from tst import _lib

    h = ctypes.c_void_p()
    rc = _lib.lib().tst_create(ctypes.byref(h), 0, 0, 1, None)
    assert rc == -2 and b"no CUDA device" in _lib.lib().tst_last_error()
    X = np.zeros((10, 3), np.float32)
    y = np.arange(10) % 2
    for call in (lambda: tst.run_tst(X, y), lambda: tst.ibp(X, y), lambda: tst.TsTForest().fit(X, y)):
        with pytest.raises(tst.TstError):
            call()


def test_product_path_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "test-set-training_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in text.replace("the CPU oracle", "").replace("CPU oracle", "").replace("oracle/hist_rf.c", ""), f


def test_resolve_max_features_and_params():
    import tst

    assert tst.resolve_max_features("sqrt", 32) == 5 and tst.resolve_max_features("sqrt", 64) == 8
    assert tst.resolve_max_features(None, 7) == 7 and tst.resolve_max_features(0.5, 10) == 5
    assert tst.resolve_max_features("log2", 64) == 6 and tst.resolve_max_features(100, 9) == 9
    p = tst.make_params(64, max_depth=None)
    assert (p.max_depth, p.min_samples_split, p.min_samples_leaf, p.max_features) == (0, 2, 1, 8)
    with pytest.raises(ValueError):
        tst.make_params(4, min_samples_split=1)


def test_encode_labels_class_order():
    from tst.api import encode_labels

    classes, idx = encode_labels(np.array(["b", "a", "c", "a"]))
    assert classes.tolist() == ["a", "b", "c"] and idx.tolist() == [1, 0, 2, 0] and idx.dtype == np.uint8


@pytest.mark.parametrize("y", [np.array([2, 0, 2, 1, 3, 3], dtype=np.uint8),            # 0 .. C-1 already: returned without a copy
                               np.array([5, 5, 7, 200], dtype=np.uint8),                 # gaps
                               np.array([0, 2, 2, 300, 4000], dtype=np.int64),           # sparse non-negative integers (bincount path)
                               np.array([-3, 7, 7, -3, 0], dtype=np.int32),              # negative values: the sort-based path
                               np.array([1.5, -2.0, 1.5, 9.25]),                         # floats
                               (np.arange(1000) * 7 % 31).astype(np.uint8)])             # 31 classes
def test_encode_labels_equals_np_unique(y):
    """Every fast path of encode_labels is np.unique(y, return_inverse=True) (SK/ensemble/_forest.py:810-827)."""
    from tst.api import encode_labels

    classes, idx = encode_labels(y)
    rc, ri = np.unique(y, return_inverse=True)
    assert np.array_equal(classes, rc) and classes.dtype == rc.dtype
    assert idx.dtype == np.uint8 and np.array_equal(idx, ri.astype(np.uint8))
    assert np.array_equal(classes[idx], y)


def test_encode_labels_rejects_too_many_classes():
    from tst import _lib
    from tst.api import encode_labels

    for y in (np.arange(_lib.MAX_CLASSES + 1, dtype=np.uint8), np.arange(_lib.MAX_CLASSES + 1, dtype=np.int64) * 1000):
        with pytest.raises(ValueError):
            encode_labels(y)


def test_ibp_vmin_matches_oracle_rule():
    import tst

    for tau, T in ((0.5, 200), (0.45, 10), (1.0, 7), (0.3333, 99)):
        assert tst.ibp_vmin(tau, T) == int(np.ceil(tau * T * 65536.0))


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_shard_trees_partitions_the_job(world):
    import tst

    n_parts, T = 10, 24
    seen = []
    for r in range(world):
        mine = tst.shard_trees(n_parts, T, r, world, 4, 20)
        assert all((p * T + t) % world == r for p, t in mine)
        seen += mine
    assert sorted(seen) == [(p, t) for p in range(n_parts) for t in range(4, 20)]
    counts = [len(tst.shard_trees(n_parts, T, r, world)) for r in range(world)]
    assert max(counts) - min(counts) <= 1 + n_parts  # interleaving balances folds across ranks


def test_synthetic_generators_are_deterministic():
    from tst import synth

    X1, y1 = synth.vsi_table(300, 16, 4, seed=3)
    X2, y2 = synth.vsi_table(300, 16, 4, seed=3)
    assert np.array_equal(X1, X2) and np.array_equal(y1, y2) and X1.dtype == np.float32 and y1.dtype == np.uint8
    X, y, g, ng = synth.grouped_table("mmmu")
    assert X.shape == (900, 16) and ng == 30 and np.bincount(g).tolist() == [30] * 30
    X, y = synth.scaleup_table(2000, 8)
    assert X.shape == (2000, 8) and set(np.unique(y)) <= {0, 1, 2, 3}
