"""scikit-learn restatement of the paper's RF diagnostic (REF/README.md:35) -- the statistical oracle.

TEST INFRASTRUCTURE / CPU BASELINE ONLY.  The reference ships no code (REF/README.md:38-39), so this
script *is* the committed description of "k-fold cross-validated Random Forest on hand-crafted
non-visual features -> out-of-fold bias score s(x) -> Iterative Bias Pruning", written with the
canonical RandomForestClassifier (defaults: gini, max_features="sqrt", full depth, bootstrap;
SK/ensemble/_forest.py:1515-1537).  It records sklearn.__version__ with every result.

Tolerance vs the histogram RF (BASELINE.json north_star): per-fold accuracy within +-1 point and
Spearman rho(s) >= 0.95, evaluated at >= 500 trees per fold and reported next to the
sklearn-vs-sklearn seed-noise floor (hard part 5).
"""
from __future__ import annotations

import os
import time

import numpy as np


def usable_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:  # pragma: no cover
        return os.cpu_count() or 1


def stratified_folds(y, k, seed=0):
    """sklearn's own StratifiedKFold (SK/model_selection/_split.py:774-840), shuffled."""
    from sklearn.model_selection import StratifiedKFold

    fold = np.zeros(len(y), dtype=np.uint8)
    for f, (_, te) in enumerate(StratifiedKFold(k, shuffle=True, random_state=seed).split(np.zeros(len(y)), y)):
        fold[te] = f
    return fold


def run_tst(X, y, k=5, n_trees=100, seed=0, n_jobs=None, fold=None, rf_kwargs=None):
    """-> dict(s, proba, fold, fold_accuracy, seconds, sklearn_version, cores)."""
    import sklearn
    from sklearn.ensemble import RandomForestClassifier

    X = np.asarray(X, dtype=np.float32)
    y = np.asarray(y)
    n = X.shape[0]
    classes = np.unique(y)
    n_jobs = usable_cores() if n_jobs is None else n_jobs
    fold = stratified_folds(y, k, seed) if fold is None else np.asarray(fold)
    proba = np.zeros((n, classes.shape[0]), dtype=np.float64)
    acc = np.zeros(k)
    t0 = time.perf_counter()
    for f in range(k):
        tr, te = fold != f, fold == f
        if not te.any():
            continue
        rf = RandomForestClassifier(n_estimators=n_trees, random_state=seed * 1000 + f, n_jobs=n_jobs,
                                    **(rf_kwargs or {}))
        rf.fit(X[tr & (fold != 255)], y[tr & (fold != 255)])
        p = rf.predict_proba(X[te])
        proba[np.ix_(te, np.searchsorted(classes, rf.classes_))] = p
        acc[f] = (classes[np.argmax(proba[te], axis=1)] == y[te]).mean()
    dt = time.perf_counter() - t0
    s = proba[np.arange(n), np.searchsorted(classes, y)]
    return dict(s=s, proba=proba, fold=fold, fold_accuracy=acc, seconds=dt, trees_per_s=k * n_trees / dt,
                sklearn_version=sklearn.__version__, cores=n_jobs)


def ibp(X, y, rounds=10, k=5, n_trees=200, tau=0.5, batch=None, seed=0, n_jobs=None):
    """IBP with the same selection rule as DESIGN.md 2.8 (s >= tau, top-`batch`, ties -> low row id)."""
    n = len(y)
    batch = n if batch is None else batch
    active = np.ones(n, dtype=bool)
    removed = []
    t0 = time.perf_counter()
    for r in range(rounds):
        idx = np.flatnonzero(active)
        res = run_tst(X[idx], y[idx], k=k, n_trees=n_trees, seed=seed + r, n_jobs=n_jobs)
        s = res["s"]
        cand = np.flatnonzero(s >= tau)
        order = cand[np.lexsort((cand, -s[cand]))][:batch]
        active[idx[order]] = False
        removed.append(len(order))
        if len(order) == 0:
            break
    removed += [0] * (rounds - len(removed))
    return dict(keep_mask=active, removed_per_round=np.array(removed), seconds=time.perf_counter() - t0)


def time_fit_sample(X, y, train_mask, n_trees, seed=0, n_jobs=None, X_test=None):
    """Time `n_trees` trees on one fold's training rows (+ OOF predict): the bounded CPU-baseline sample.
    train_mask=None: X, y already are the training rows (the slicing copy then stays outside the timed region)."""
    from sklearn.ensemble import RandomForestClassifier

    n_jobs = usable_cores() if n_jobs is None else n_jobs
    t0 = time.perf_counter()
    rf = RandomForestClassifier(n_estimators=n_trees, random_state=seed, n_jobs=n_jobs)
    if train_mask is None:
        rf.fit(X, y)
    else:
        rf.fit(X[train_mask], y[train_mask])
    if X_test is not None:
        rf.predict_proba(X_test)
    dt = time.perf_counter() - t0
    return dict(seconds=dt, trees_per_s=n_trees / dt, cores=n_jobs)


if __name__ == "__main__":  # python oracle/sklearn_tst.py  -> VSI-shaped config 0 on CPU
    import json
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "test-set-training_b200"))
    from tst import synth

    X, y = synth.vsi_table()
    r = run_tst(X, y, k=5, n_trees=100)
    print(json.dumps(dict(config="vsi 5k x 32, 5-fold x 100", seconds=r["seconds"], trees_per_s=r["trees_per_s"],
                          cores=r["cores"], fold_accuracy=r["fold_accuracy"].tolist(),
                          sklearn=r["sklearn_version"])))
