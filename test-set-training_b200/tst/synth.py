"""Synthetic benchmark-shaped feature tables; seeds are part of the contract.

The reference ships no data and the container has no network, so every config in BASELINE.json
runs on these generators.  Shapes follow the benchmarks the paper names (REF/README.md:35); group
counts/sizes for CV-Bench / MMMU / VideoMME are from memory of those benchmarks and are parameters.
"""
from __future__ import annotations

import numpy as np


def vsi_table(n: int = 5000, d: int = 32, n_classes: int = 4, seed: int = 0):
    """VSI-Bench-shaped table: d/2 continuous (half informative), d/4 low-cardinality, d/4 counts."""
    rng = np.random.default_rng(seed)
    n_cont, n_cat = d // 2, d // 4
    n_cnt = d - n_cont - n_cat
    cont = rng.standard_normal((n, n_cont))
    card = rng.integers(2, 13, size=n_cat)
    cat = np.stack([rng.integers(0, c, size=n) for c in card], axis=1).astype(np.float64)
    lam = rng.uniform(1.0, 20.0, size=n_cnt)
    cnt = rng.poisson(lam, size=(n, n_cnt)).astype(np.float64)
    X = np.concatenate([cont, cat, cnt], axis=1)
    # sparse linear score over standardised columns + Gumbel noise -> ~50 % blind accuracy
    Z = (X - X.mean(0)) / (X.std(0) + 1e-9)
    W = np.zeros((d, n_classes))
    informative = np.concatenate([np.arange(n_cont // 2), n_cont + np.arange(n_cat // 2),
                                  n_cont + n_cat + np.arange(n_cnt // 2)])
    W[informative] = rng.standard_normal((informative.shape[0], n_classes))
    score = Z @ W / np.sqrt(informative.shape[0]) * 2.0 + rng.gumbel(size=(n, n_classes)) * 0.9
    y = np.argmax(score, axis=1).astype(np.uint8)
    return X.astype(np.float32), y


def numeric_table(n: int = 4000, d: int = 16, seed: int = 0):
    """Numeric-answer question type (object counts, distances ...): the same column mix as vsi_table with a positive
    real-valued answer that a blind model can partly predict from a few columns (non-linear, heteroscedastic)."""
    X, _ = vsi_table(n, d, 2, seed)
    rng = np.random.default_rng(seed + 7919)
    Z = (X - X.mean(0)) / (X.std(0) + 1e-9)
    y = 5.0 + 2.0 * Z[:, 0] + 1.5 * (Z[:, 1] > 0.3) + Z[:, 2] * Z[:, d // 2] + 0.8 * rng.standard_normal(n)
    return X, np.exp(0.35 * y).astype(np.float64)


def scaleup_table(n: int = 1_000_000, d: int = 64, n_classes: int = 4, seed: int = 0):
    """Scale-up table: standard-normal float32, labels = argmax(X W + 1.5 noise), W 25 % dense."""
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, d), dtype=np.float32)
    W = rng.standard_normal((d, n_classes)).astype(np.float32)
    W *= (rng.random((d, n_classes)) < 0.25)
    score = X @ W + 1.5 * rng.standard_normal((n, n_classes), dtype=np.float32)
    y = np.argmax(score, axis=1).astype(np.uint8)
    return X, y


_GROUPED = {
    # name: (groups, rows per group, features, seed)
    "cvbench": (4, 660, 16, 1),
    "mmmu": (30, 30, 16, 2),
    "videomme": (12, 225, 24, 3),
}


def grouped_table(kind: str, n_classes: int = 4, groups=None, rows=None, d=None, seed=None):
    """Per-question-type tables: one forest set per group (BASELINE.json configs[2])."""
    g0, r0, d0, s0 = _GROUPED[kind]
    groups, rows, d, seed = groups or g0, rows or r0, d or d0, s0 if seed is None else seed
    Xs, ys, gs = [], [], []
    for g in range(groups):
        X, y = vsi_table(rows, d, n_classes, seed=seed * 1000 + g)
        Xs.append(X); ys.append(y); gs.append(np.full(rows, g, dtype=np.int32))
    return np.concatenate(Xs), np.concatenate(ys), np.concatenate(gs), groups
