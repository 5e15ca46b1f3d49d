"""Oracle (and host edge rule) against the committed golden vectors in tests/golden/ (made by make_golden.py)."""
import os

import numpy as np
import pytest

from oracle import hist_rf as H

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def sk():
    return np.load(os.path.join(GOLD, "sklearn_binning.npz"))


@pytest.fixture(scope="module")
def spec():
    return np.load(os.path.join(GOLD, "spec_small.npz"))


@pytest.mark.parametrize("name", ["normal", "poisson", "lowcard", "heavy_ties"])
def test_edges_equal_sklearn_thresholds(sk, name):
    """Vectors produced by sklearn's _find_binning_thresholds (SK/.../binning.py:31-111)."""
    from tst import binning

    col, thr = sk[f"col_{name}"], sk[f"thr_{name}"]
    a = np.sort(col.astype(np.float64))
    n, B = len(a), 255
    for fn in (H.find_edges_column, binning.column_edges):  # oracle restatement and the product's host rule
        e = fn(col)
        assert e.shape == thr.shape, name
        diff = np.flatnonzero(e != thr)
        if len(np.unique(a)) <= B:
            assert diff.size == 0, name
            continue
        # Quantile rule: sklearn evaluates n*q in floating point, the spec in exact integers.  They can
        # only disagree where i*n/B is an exact integer (numpy may land one ulp off and take a single
        # order statistic instead of the average); everywhere else the edges are bit-identical.
        assert e.shape[0] == B - 1
        for i0 in diff:
            i = i0 + 1
            assert (i * n) % B == 0, (name, i0)
            j = i * n // B
            assert e[i0] == (a[j - 1] + a[j]) * 0.5 and thr[i0] in (a[j - 1], a[j], e[i0]), (name, i0)
        assert diff.size <= 8


def test_bins_equal_sklearn_binmapper(sk):
    """Vectors produced by sklearn's _BinMapper.transform (SK/.../_binning.pyx:50-85)."""
    X = sk["binmapper_X"]
    edges, ne = H.find_edges(X)
    assert np.array_equal(H.bin_table(X, edges, ne).T, sk["binmapper_bins"])


def test_spec_frozen(spec):
    """The normative spec may not drift: same inputs -> same folds, votes, s(x), tree, counters, IBP."""
    X, y = spec["X"], spec["y"]
    edges, ne = H.find_edges(X)
    assert np.array_equal(edges, spec["edges"]) and np.array_equal(ne, spec["n_edges"])
    bins = H.bin_table(X, edges, ne)
    assert np.array_equal(bins, spec["bins"])
    fold = H.assign_folds(y, np.ones(len(y), np.uint8), 4, 3)
    assert np.array_equal(fold, spec["fold"])
    prm = H.make_params(12)
    votes, st, forest = H.fit_oof(bins, y, fold, 4, 6, 3, prm, seed=11, keep=True)
    assert np.array_equal(votes, spec["votes"])
    s, cor, tot = H.finalize(votes, y, fold, 4, 6)
    assert np.array_equal(s.view(np.uint32), spec["s"].view(np.uint32))
    assert np.array_equal(cor, spec["correct"]) and np.array_equal(tot, spec["total"])
    t = forest.tree(2, 3)
    assert np.array_equal(t.nodes, spec["tree_nodes"]) and np.array_equal(t.counts, spec["tree_counts"])
    keys = ("row_visits", "nodes", "leaves", "searched", "fallback", "inbag_rows", "max_depth")
    assert [st[k] for k in keys] == spec["stats"].tolist()
    act, removed, s_rounds = H.ibp(bins, y, 4, 6, 3, prm, seed=11, rounds=3, tau=0.4, batch=25)
    assert np.array_equal(act, spec["ibp_active"]) and np.array_equal(removed, spec["ibp_removed"])
    assert np.array_equal(s_rounds.view(np.uint32), spec["ibp_s"].view(np.uint32))


@pytest.mark.gpu
def test_gpu_matches_golden(spec):
    """The CUDA path against the committed fixture (no oracle in the loop)."""
    import tst

    X, y = spec["X"], spec["y"]
    res = tst.run_tst(X, y, k=4, n_trees=6, seed=11)
    assert np.array_equal(res.fold_id, spec["fold"])
    assert np.array_equal(res.oof_votes, spec["votes"])
    assert np.array_equal(res.s.view(np.uint32), spec["s"].view(np.uint32))
    r = tst.ibp(X, y, rounds=3, k=4, n_trees=6, tau=0.4, batch=25, seed=11)
    assert np.array_equal(r.keep_mask, spec["ibp_active"].astype(bool)) and r.removed_per_round.tolist() == spec["ibp_removed"].tolist()
    assert np.array_equal(r.s_per_round.view(np.uint32), spec["ibp_s"].view(np.uint32))
    clf = tst.TsTForest(n_estimators=6, random_state=11)
    # tree (part 2, t 3) of the k-fold run is not reproduced by a plain fit; check the binned table instead
    clf.fit(X, y)
    assert np.array_equal(clf._eng.get_bins(), spec["bins"])
