"""Generates tests/golden/*.npz.  Run from the repo root:  python tests/golden/make_golden.py

Two kinds of vectors (the reference repo ships none -- "parity unpinned"):
  sklearn_*   outputs of the third-party reference the oracle restates (scikit-learn, imported here):
              bin thresholds of _find_binning_thresholds and _BinMapper bins on fixed seeded inputs.
  spec_*      outputs of the CPU oracle (oracle/hist_rf.c) on small seeded tables: they freeze the
              normative spec so that neither side can drift between rounds.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "test-set-training_b200"))
HERE = os.path.dirname(os.path.abspath(__file__))


def sklearn_vectors():
    import sklearn
    from sklearn.ensemble._hist_gradient_boosting.binning import _BinMapper, _find_binning_thresholds

    rng = np.random.default_rng(2024)
    cols = {
        "normal": rng.standard_normal(3000).astype(np.float32),
        "poisson": rng.poisson(6.0, 2000).astype(np.float32),
        "lowcard": rng.integers(0, 7, 500).astype(np.float32),
        "heavy_ties": np.round(rng.standard_normal(4000) * 40).astype(np.float32),
    }
    out = {"sklearn_version": np.array(sklearn.__version__)}
    for k, v in cols.items():
        out[f"col_{k}"] = v
        out[f"thr_{k}"] = _find_binning_thresholds(v.astype(np.float64), 255)
    X = np.stack([cols["normal"][:500], cols["poisson"][:500], cols["lowcard"]], axis=1).astype(np.float64)
    bm = _BinMapper(n_bins=256, subsample=None, random_state=0).fit(X)
    out["binmapper_X"] = X.astype(np.float32)
    out["binmapper_bins"] = bm.transform(X)
    np.savez_compressed(os.path.join(HERE, "sklearn_binning.npz"), **out)


def spec_vectors():
    from oracle import hist_rf as H
    from tst import synth

    X, y = synth.vsi_table(400, 12, 3, seed=5)
    edges, ne = H.find_edges(X)
    bins = H.bin_table(X, edges, ne)
    fold = H.assign_folds(y, np.ones(len(y), np.uint8), 4, 3)
    prm = H.make_params(12)
    votes, st, forest = H.fit_oof(bins, y, fold, 4, 6, 3, prm, seed=11, keep=True, nthreads=1)
    s, cor, tot = H.finalize(votes, y, fold, 4, 6)
    t = forest.tree(2, 3)
    act, removed, s_rounds = H.ibp(bins, y, 4, 6, 3, prm, seed=11, rounds=3, tau=0.4, batch=25)
    np.savez_compressed(os.path.join(HERE, "spec_small.npz"), X=X, y=y, edges=edges, n_edges=ne, bins=bins, fold=fold,
                        votes=votes, s=s, correct=cor, total=tot, tree_nodes=t.nodes, tree_counts=t.counts,
                        stats=np.array([st[k] for k in ("row_visits", "nodes", "leaves", "searched", "fallback", "inbag_rows", "max_depth")]),
                        ibp_active=act, ibp_removed=removed, ibp_s=s_rounds)


def spec_regression_vectors():
    """Squared-error criterion on ordinal levels: frozen oracle output."""
    from oracle import hist_rf as H
    from tst import quantise_target, synth

    X, y = synth.numeric_table(500, 10, seed=3)
    lev, lo, step = quantise_target(y, 16)
    edges, ne = H.find_edges(X)
    bins = H.bin_table(X, edges, ne)
    fold = H.assign_folds(lev, np.ones(len(lev), np.uint8), 3, 16)
    prm = H.make_params(10, max_features=1.0, criterion="squared_error_levels")
    votes, st, forest = H.fit_oof(bins, lev, fold, 3, 5, 16, prm, seed=21, keep=True, nthreads=1)
    t = forest.tree(1, 2)
    np.savez_compressed(os.path.join(HERE, "spec_regression.npz"), X=X, y=y, levels=lev, lo=lo, step=step, bins=bins, fold=fold,
                        votes=votes, tree_nodes=t.nodes, tree_counts=t.counts,
                        stats=np.array([st[k] for k in ("row_visits", "nodes", "leaves", "searched", "fallback", "inbag_rows", "max_depth")]))


def sklearn_seed_averaged(n_seeds=8, n_seeds_binned=4, T=500):
    """The statistical oracle, seed-averaged (python tests/golden/make_golden.py sklearn_rf; ~6 min on 8 cores).
    sklearn's RandomForestClassifier on the VSI-shaped table with the oracle's fold assignment: per-seed per-fold accuracy
    and s(x) at T trees per fold -- on the raw floats, and on the uint8-binned table (same candidate splits as the
    histogram RF: differs only in RNG and tie-breaks; SURVEY 8(c) third tier) where the impurity-based feature
    importances are recorded as well.  Seed-averaging is what makes the +-1 point per-fold tolerance testable: two
    sklearn seeds alone differ by 0.8 .. 1.3 points per fold at 500 .. 2000 trees."""
    import sklearn
    from sklearn.ensemble import RandomForestClassifier

    from oracle import hist_rf as H
    from oracle import sklearn_tst as S
    from tst import synth

    X, y = synth.vsi_table()
    edges, ne = H.find_edges(X)
    bins = H.bin_table(X, edges, ne)
    fold = H.assign_folds(y, np.ones(len(y), np.uint8), 5, 4)
    raw = [S.run_tst(X, y, 5, T, seed=sd, fold=fold) for sd in range(n_seeds)]
    Xb = bins.T.astype(np.float32)
    binned = [S.run_tst(Xb, y, 5, T, seed=100 + sd, fold=fold) for sd in range(n_seeds_binned)]
    imp = np.zeros(X.shape[1])
    for f in range(5):  # importances of the same 5 training sets on the binned table
        rf = RandomForestClassifier(n_estimators=T, random_state=7 + f, n_jobs=-1).fit(Xb[fold != f], y[fold != f])
        imp += rf.feature_importances_ / 5
    np.savez_compressed(os.path.join(HERE, "sklearn_rf_vsi.npz"), sklearn_version=np.array(sklearn.__version__), n_trees=T, fold=fold,
                        acc_raw=np.stack([r["fold_accuracy"] for r in raw]), s_raw=np.stack([r["s"] for r in raw]).astype(np.float32),
                        acc_binned=np.stack([r["fold_accuracy"] for r in binned]), s_binned=np.stack([r["s"] for r in binned]).astype(np.float32),
                        importances_binned=imp)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "sklearn_rf":
        sklearn_seed_averaged()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "regression":
        spec_regression_vectors()
        sys.exit(0)
    sklearn_vectors()
    spec_vectors()
    spec_regression_vectors()
    print("wrote", sorted(os.listdir(HERE)))
