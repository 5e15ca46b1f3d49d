"""Host-side bin-edge selection.  The *mapping* of the table to uint8 bins is the CUDA
kernel k_bin_map; the edges are a one-off O(n log n) per table and are shared with the CPU oracle.

Rule (restating SK/ensemble/_hist_gradient_boosting/binning.py:31-111 with exact integer ranks): NaNs
are ignored; with v = sorted non-missing values and B = n_bins - 1 non-missing bins,
  * <= B distinct values: edges are the midpoints of consecutive distinct values;
  * otherwise edge_i (i = 1..B-1) is v[j] if i*len(v) % B else (v[j-1] + v[j]) / 2, j = i*len(v) // B
    ('averaged_inverted_cdf'), duplicates removed.
All arithmetic in float64; a midpoint of two float32 values is exact in float64.
"""
from __future__ import annotations

import numpy as np

MAX_EDGES = 255


def column_edges(col: np.ndarray, n_bins: int = 256) -> np.ndarray:
    if not 3 <= n_bins <= 256:
        raise ValueError("n_bins must be in [3, 256]")
    B = n_bins - 1
    v = np.asarray(col, dtype=np.float64).ravel()
    v = v[v == v]
    if v.size == 0:
        return np.empty(0, np.float64)
    v.sort()
    uniq = np.unique(v)
    if uniq.size <= 1:
        return np.empty(0, np.float64)
    if uniq.size <= B:
        e = 0.5 * (uniq[1:] + uniq[:-1])
    else:
        t = np.arange(1, B, dtype=np.int64) * v.size
        j, rem = np.divmod(t, B)
        e = np.where(rem != 0, v[j], 0.5 * (v[j - 1] + v[j]))
        e = np.unique(e)
    return np.minimum(e, 1e300)


def table_edges(X: np.ndarray, n_bins: int = 256):
    """-> (edges [d, 255] float64 zero-padded, n_edges [d] int32)."""
    X = np.asarray(X)
    d = X.shape[1]
    edges = np.zeros((d, MAX_EDGES), np.float64)
    n_edges = np.zeros(d, np.int32)
    for f in range(d):
        e = column_edges(X[:, f], n_bins)
        n_edges[f] = e.size
        edges[f, : e.size] = e
    return edges, n_edges


def bin_rows(X: np.ndarray, edges: np.ndarray, n_edges: np.ndarray) -> np.ndarray:
    """Host bin map for a few rows (interpretability helpers; the product path bins on the GPU with k_bin_map):
    bin = number of edges strictly below x, NaN -> 255.  -> [n, d] uint8."""
    X = np.asarray(X, dtype=np.float32)
    out = np.empty(X.shape, dtype=np.uint8)
    for f in range(X.shape[1]):
        col = X[:, f].astype(np.float64)
        b = np.searchsorted(edges[f, : int(n_edges[f])], col, side="left")
        out[:, f] = np.where(np.isnan(col), 255, b).astype(np.uint8)
    return out
