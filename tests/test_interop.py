"""On-disk format and scikit-learn interop.  The CPU tests build the forest with the oracle (same
canonical node arrays as tst_export_tree); the GPU test goes through TsTForest."""
import numpy as np
import pytest

from oracle import hist_rf as H


def _forest_from_oracle(X, y, T, C, seed=3, **pk):
    from tst.interop import ExportedForest

    edges, ne = H.find_edges(X)
    bins = H.bin_table(X, edges, ne)
    fold = np.full(len(y), H.FOLD_TRAIN_ONLY, np.uint8)
    _, _, forest = H.fit_oof(bins, y, fold, 1, T, C, H.make_params(X.shape[1], **pk), seed=seed, keep=True)
    trees = [forest.tree(0, t) for t in range(T)]
    ex = ExportedForest([t.nodes for t in trees], [t.counts for t in trees], edges, ne, np.arange(C), X.shape[1])
    return ex, forest, bins


def _table(n=900, d=10, C=3, seed=5):
    from tst import synth

    X, y = synth.vsi_table(n, d, C, seed=seed)
    X[::7, 2] = np.nan  # missing values: bin 255
    X[:, 4] = np.where(np.isnan(X[:, 4]), 0, X[:, 4])
    return X, y


def test_sklearn_tree_roundtrip_and_tools(tmp_path):
    import sklearn.tree as sktree

    X, y = _table()
    ex, forest, bins = _forest_from_oracle(X, y, T=6, C=3)
    rf = ex.to_sklearn()
    # genuine Tree objects: sklearn's own tooling works on them
    t0 = rf.estimators_[0]
    assert type(t0.tree_).__name__ == "Tree" and t0.tree_.node_count == len(ex.nodes[0])
    text = sktree.export_text(t0, max_depth=3)
    assert "feature_" in text and "class" in text
    leaves = t0.apply(X.astype(np.float32))
    assert np.array_equal(leaves, ex.apply_tree(ex.nodes[0], H.bin_table(X, ex.edges, ex.n_edges).T.copy()))
    # predict_proba of the sklearn forest == our integer votes to 2^-16 per tree
    votes = forest.predict_votes(bins)
    assert np.array_equal(ex.predict_votes(X), votes)
    ours = votes.astype(np.float64) / (6 * 65536.0)
    assert np.abs(rf.predict_proba(X.astype(np.float32)) - ours).max() <= 2.0 ** -16 + 1e-12
    # sklearn's own importances on the converted trees == ours from the same trees
    from tst.features import gini_importances

    mine = np.mean([gini_importances(n, c, X.shape[1]) for n, c in zip(ex.nodes, ex.counts)], axis=0)
    assert np.allclose(rf.feature_importances_, mine / mine.sum(), atol=1e-12)
    # disk round trip
    from tst.interop import load_forest

    path = str(tmp_path / "forest.npz")
    ex.save(path)
    back = load_forest(path)
    assert len(back.nodes) == 6 and all(np.array_equal(a, b) for a, b in zip(back.nodes, ex.nodes))
    assert all(np.array_equal(a, b) for a, b in zip(back.counts, ex.counts))
    assert np.array_equal(back.predict_votes(X), votes)
    assert np.allclose(back.to_sklearn().predict_proba(X.astype(np.float32)), rf.predict_proba(X.astype(np.float32)))


def test_nan_only_split_exports_infinite_threshold():
    """A split that separates every finite value from the NaN bin sits at bin == n_edges[f]: raw threshold +inf."""
    from tst.features import to_sklearn_arrays
    from tst.interop import ExportedForest

    rng = np.random.default_rng(0)
    x = np.where(rng.random(400) < 0.5, np.nan, 1.0)  # one finite value: no edges, bins are 0 and 255
    X = np.stack([x, rng.standard_normal(400)], axis=1).astype(np.float32)
    y = np.isnan(x).astype(np.uint8)
    ex, forest, bins = _forest_from_oracle(X, y, T=3, C=2, max_features=None)
    assert ex.n_edges[0] == 0
    nd = ex.nodes[0]
    assert nd["feature"][0] == 0 and nd["bin"][0] == 0  # root splits finite | NaN
    assert ex.raw_thresholds(0)[0] == np.inf
    assert to_sklearn_arrays(nd, ex.counts[0], ex.edges, ex.n_edges)["threshold"][0] == np.inf
    rf = ex.to_sklearn()
    assert (rf.predict(X) == y).all()


@pytest.mark.gpu
def test_forest_export_save_load_sklearn_gpu(tmp_path):
    import tst

    X, y = _table(2000, 16, 4, seed=9)
    clf = tst.TsTForest(n_estimators=8, random_state=2).fit(X, y)
    ex = clf.export()
    votes = clf.predict_votes(X[:500])
    assert np.array_equal(ex.predict_votes(X[:500]), votes)
    rf = clf.to_sklearn()
    assert np.abs(rf.predict_proba(X[:500]) - votes / (8 * 65536.0)).max() <= 2.0 ** -16 + 1e-12
    assert np.allclose(rf.feature_importances_, clf.feature_importances_, atol=1e-6)
    path = str(tmp_path / "f.npz")
    clf.save(path)
    assert np.array_equal(tst.load_forest(path).predict_votes(X[:500]), votes)


@pytest.mark.gpu
def test_regression_forest_to_sklearn_gpu(tmp_path):
    """criterion="squared_error": the exported forest becomes a genuine RandomForestRegressor whose predictions equal the
    device's Q16.16 leaf-mean votes to 2^-16 of the target's span."""
    import tst
    from tst import synth

    X, y = synth.numeric_table(1500, 10, seed=6)
    reg = tst.TsTForest(n_estimators=6, criterion="squared_error", max_features=1.0, random_state=4).fit(X, y)
    ours = reg.predict(X[:400])
    rf = reg.to_sklearn()
    assert type(rf).__name__ == "RandomForestRegressor" and type(rf.estimators_[0].tree_).__name__ == "Tree"
    span = float(y.max() - y.min())
    assert np.abs(rf.predict(X[:400]) - ours).max() <= span * 2.0 ** -15
    path = str(tmp_path / "r.npz")
    reg.save(path)
    back = tst.load_forest(path)
    assert back.criterion == "squared_error" and np.allclose(back.to_sklearn().predict(X[:400]), rf.predict(X[:400]))
