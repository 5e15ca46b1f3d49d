"""Statistical gate of the histogram RF against sklearn's exact-split RandomForestClassifier
(BASELINE.json north_star: per-fold accuracy within +-1 point, Spearman rho(s) >= 0.95), evaluated at
500 trees per fold and printed next to the sklearn-vs-sklearn seed-noise floor (SURVEY.md section 7,
hard part 5: at 100 trees two sklearn seeds only agree to rho = 0.952 / 1.7 points)."""
import numpy as np
import pytest
from scipy.stats import spearmanr

from oracle import hist_rf as H
from oracle import sklearn_tst as S

pytestmark = pytest.mark.slow


def test_hist_rf_within_tolerance_of_sklearn():
    from tst import synth

    X, y = synth.vsi_table()
    edges, ne = H.find_edges(X)
    bins = H.bin_table(X, edges, ne)
    fold = H.assign_folds(y, np.ones(len(y), np.uint8), 5, 4)
    T = 500
    votes, st, _ = H.fit_oof(bins, y, fold, 5, T, 4, H.make_params(32), seed=0)
    s, cor, tot = H.finalize(votes, y, fold, 5, T)
    acc = cor / tot
    a = S.run_tst(X, y, 5, T, seed=0, fold=fold)
    b = S.run_tst(X, y, 5, T, seed=1, fold=fold)
    rho = spearmanr(s, a["s"]).statistic
    rho_floor = spearmanr(a["s"], b["s"]).statistic
    dacc = np.abs(acc - a["fold_accuracy"]) * 100
    dacc_floor = np.abs(a["fold_accuracy"] - b["fold_accuracy"]) * 100
    print(f"\nrho(hist, sklearn) = {rho:.4f}   sklearn-vs-sklearn floor = {rho_floor:.4f}")
    print(f"per-fold |d acc| points: hist-vs-sklearn {np.round(dacc, 2)}   floor {np.round(dacc_floor, 2)}")
    assert rho >= 0.95
    assert abs(acc.mean() - a["fold_accuracy"].mean()) * 100 <= 1.0  # tolerance: +-1 point on the mean ...
    assert dacc.max() <= max(1.0, dacc_floor.max() + 0.75)           # ... per fold no worse than sklearn's own seed noise
