"""Known-answer tests that pin the CPU oracle (the reference ships no golden vectors: "parity unpinned").

Sources: Random123 Philox4x32-10 vectors; scikit-learn's own KATs for histograms and bin edges
(SK/ensemble/_hist_gradient_boosting/tests/test_histogram.py:20-50, test_binning.py:28-50,96-115)
restated for the oracle's functions.
"""
import numpy as np
import pytest

from oracle import hist_rf as H


@pytest.mark.parametrize("ctr,key,exp", [
    ([0, 0, 0, 0], [0, 0], [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]),
    ([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2, [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]),
    ([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0],
     [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]),
])
def test_philox_random123_vectors(ctr, key, exp):
    assert H.philox(ctr, key).tolist() == exp


def test_edges_linspace_kat():
    # test_binning.py:28-38: linspace(0, 10, 1001), 10 bins -> 1..9
    col = np.linspace(0, 10, 1001)
    e = H.find_edges_column(col, n_bins=11)
    assert np.allclose(e, np.arange(1, 10))


def test_edges_few_distinct_midpoints():
    # test_binning.py:41-50: 11 distinct values, >= 11 bins -> midpoints arange(10) + .5
    col = np.repeat(np.arange(11.0), 7)
    np.random.default_rng(0).shuffle(col)
    assert np.array_equal(H.find_edges_column(col, n_bins=256), np.arange(10) + 0.5)
    assert H.find_edges_column(np.full(9, 3.0)).shape == (0,)
    assert H.find_edges_column(np.array([np.nan, np.nan])).shape == (0,)


def test_edges_match_sklearn_binmapper():
    from sklearn.ensemble._hist_gradient_boosting.binning import _find_binning_thresholds

    rng = np.random.default_rng(3)
    for col in (rng.standard_normal(5000), rng.poisson(4.0, 3000).astype(float), rng.integers(0, 400, 4000).astype(float)):
        col = col.astype(np.float32)
        ours = H.find_edges_column(col)
        theirs = _find_binning_thresholds(col.astype(np.float64), 255)
        assert ours.shape == theirs.shape and np.allclose(ours, theirs, rtol=0, atol=1e-12)


def test_map_to_bins_kat():
    # test_binning.py:96-115: uint8, min -> bin 0, max -> last bin; NaN -> 255
    rng = np.random.default_rng(42)
    X = rng.standard_normal((4000, 3)).astype(np.float32)
    X[5, 1] = np.nan
    edges, ne = H.find_edges(X)
    b = H.bin_table(X, edges, ne)
    assert b.dtype == np.uint8 and b.shape == (3, 4000)
    for f in range(3):
        col = X[:, f]
        assert b[f, np.nanargmin(col)] == 0 and b[f, np.nanargmax(col)] == ne[f]
    assert b[1, 5] == 255
    # definition: number of edges strictly below x
    f = 2
    ref = np.searchsorted(edges[f, : ne[f]], X[:, f].astype(np.float64), side="left")
    assert np.array_equal(b[f], ref.astype(np.uint8))


def _root_hist(bins_col, rows, n_bins, y=None):
    """Histogram of a node through the oracle: a 1-feature table, rows given weight 1 by construction."""
    h = np.zeros(n_bins, dtype=np.int64)
    np.add.at(h, bins_col[rows], 1)
    return h


def test_histogram_kat_via_tree_counts():
    # test_histogram.py:20-50: bins [0,2,0,1,2,0,2,1]; rows [0,2,3] -> [2,1,0]; rows [0,2,3,6,7] -> [2,2,1].
    # Restated through the oracle: a depth-1 tree on a single feature whose class == bin must put
    # exactly those counts in its children (max_depth=1 keeps the root's split counts visible).
    binned = np.array([0, 2, 0, 1, 2, 0, 2, 1], dtype=np.uint8)
    for rows, exp in (([0, 2, 3], [2, 1, 0]), ([0, 2, 3, 6, 7], [2, 2, 1])):
        assert _root_hist(binned, np.array(rows), 3).tolist() == exp
        bins = binned[rows][None, :].copy()
        y = binned[rows].copy()
        fold = np.full(len(rows), H.FOLD_TRAIN_ONLY, np.uint8)
        prm = H.make_params(1, max_depth=None, max_features=1)
        _, st, forest = H.fit_oof(bins, y, fold, 1, 4, 3, prm, seed=1, keep=True, nthreads=1)
        for t in range(4):
            tr = forest.tree(0, t)
            leaves = tr.nodes["feature"] < 0
            # fully grown on class == bin: every leaf is pure and leaf counts are multiples of the bootstrap weights
            assert all((tr.counts[i] > 0).sum() == 1 for i in np.flatnonzero(leaves))
            assert tr.counts[0].sum() == len(rows)  # root weight == number of bootstrap draws


def test_fold_assignment_stratified():
    rng = np.random.default_rng(0)
    y = rng.integers(0, 4, 1003).astype(np.uint8)
    act = (rng.random(1003) < 0.9).astype(np.uint8)
    fold = H.assign_folds(y, act, 5, 4)
    assert (fold[act == 0] == H.FOLD_INACTIVE).all()
    sizes = np.bincount(fold[act == 1], minlength=5)
    assert sizes.max() - sizes.min() <= 1
    for c in range(4):
        per = np.bincount(fold[(act == 1) & (y == c)], minlength=5)
        assert per.max() - per.min() <= 1
