"""On-disk format and scikit-learn interop of a kept forest.

    forest = TsTForest(...).fit(X, y)
    ex = forest.export()                       # ExportedForest: canonical node arrays of every tree + bin edges, on the host
    ex.save("forest.npz"); ex = load_forest("forest.npz")
    rf = ex.to_sklearn()                       # a genuine sklearn RandomForestClassifier (Tree objects via __setstate__)
    rf.predict_proba(X); sklearn.tree.export_text(rf.estimators_[0]); rf.estimators_[0].apply(X)

The field set is sklearn's Tree node record (SK/tree/_tree.pxd:15-25: left_child, right_child, feature, threshold,
impurity, n_node_samples, weighted_n_node_samples, missing_go_to_left) and its `values` array.  A histogram split
"bin <= b" is the raw-space split "x <= edges[feature][b]" exactly (bin(x) = number of edges strictly below x), so the
converted trees route every row -- including NaN (bin 255, always right) -- like the device trees do.

Host-side, numpy only: this is the interchange format, not the hot path.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import binning

FORMAT_VERSION = 1


@dataclass
class ExportedForest:
    nodes: list          # per tree: structured array (api.NODE_DTYPE), canonical breadth-first order
    counts: list         # per tree: [n_nodes, C] uint32 bootstrap-weighted class counts
    edges: np.ndarray    # [d, 255] float64 bin edges
    n_edges: np.ndarray  # [d] int32
    classes: np.ndarray
    n_features: int
    criterion: str = "gini"      # "squared_error": counts[t] is [n_nodes, 2] = (low, high) 32 bits of S = sum of w * yq, see node_values
    target_lo: float = 0.0       # squared_error: y = target_lo + yq / 65535 * target_span
    target_span: float = 1.0
    meta: dict | None = None

    # ---- disk
    def save(self, path: str) -> None:
        off = np.zeros(len(self.nodes) + 1, dtype=np.int64)
        for i, n in enumerate(self.nodes):
            off[i + 1] = off[i] + len(n)
        np.savez_compressed(path, format_version=FORMAT_VERSION, tree_offsets=off, nodes=np.concatenate(self.nodes),
                            counts=np.concatenate(self.counts), edges=self.edges, n_edges=self.n_edges, classes=self.classes,
                            n_features=self.n_features, criterion=self.criterion, target_lo=self.target_lo, target_span=self.target_span)

    # ---- sklearn
    def raw_thresholds(self, t: int) -> np.ndarray:
        nd = self.nodes[t]
        f, b = np.maximum(nd["feature"], 0), np.maximum(nd["bin"], 0)
        thr = np.where(b < self.n_edges[f], self.edges[f, np.minimum(b, self.edges.shape[1] - 1)], np.inf)
        return np.where(nd["feature"] >= 0, thr, -2.0)

    def node_values(self, t: int) -> np.ndarray:
        """squared_error: the mean target of every node of tree `t` on the target's scale (S / N mapped back)."""
        c = self.counts[t]
        S = c[:, 0].astype(np.float64) + c[:, 1].astype(np.float64) * 4294967296.0
        N = self.nodes[t]["n_weighted"].astype(np.float64)
        return self.target_lo + np.divide(S, N, out=np.zeros_like(S), where=N > 0) / 65535.0 * self.target_span

    def _sklearn_regression_tree(self, t: int):
        from sklearn.tree._tree import Tree

        nd = self.nodes[t]
        tree = Tree(self.n_features, np.array([1], dtype=np.intp), 1)
        dt = tree.__getstate__()["nodes"].dtype
        rec = np.zeros(len(nd), dtype=dt)
        internal = nd["feature"] >= 0
        rec["left_child"], rec["right_child"] = nd["left"], nd["right"]
        rec["feature"] = np.where(internal, nd["feature"], -2)
        rec["threshold"] = self.raw_thresholds(t)
        rec["n_node_samples"] = nd["end"].astype(np.int64) - nd["begin"].astype(np.int64)
        rec["weighted_n_node_samples"] = nd["n_weighted"]
        # (node impurities would need the sum of squares, which the trees do not carry: left at 0)
        tree.__setstate__(dict(max_depth=int(nd["depth"].max()), node_count=len(nd), nodes=rec,
                               values=np.ascontiguousarray(self.node_values(t)[:, None, None])))
        return tree

    def sklearn_tree(self, t: int):
        """Tree `t` as a genuine sklearn.tree._tree.Tree (state dict -> __setstate__)."""
        from sklearn.tree._tree import Tree

        if self.criterion == "squared_error":
            return self._sklearn_regression_tree(t)
        nd, cnt = self.nodes[t], self.counts[t].astype(np.float64)
        C = cnt.shape[1]
        tree = Tree(self.n_features, np.array([C], dtype=np.intp), 1)
        dt = tree.__getstate__()["nodes"].dtype
        N = cnt.sum(1)
        rec = np.zeros(len(nd), dtype=dt)
        internal = nd["feature"] >= 0
        rec["left_child"] = nd["left"]
        rec["right_child"] = nd["right"]
        rec["feature"] = np.where(internal, nd["feature"], -2)
        rec["threshold"] = self.raw_thresholds(t)
        rec["impurity"] = 1.0 - np.divide((cnt * cnt).sum(1), N * N, out=np.ones_like(N), where=N > 0)
        rec["n_node_samples"] = (nd["end"].astype(np.int64) - nd["begin"].astype(np.int64))
        rec["weighted_n_node_samples"] = N
        if "missing_go_to_left" in dt.names:
            rec["missing_go_to_left"] = 0  # NaN sits in bin 255, above every split bin: always right
        values = np.divide(cnt, N[:, None], out=np.zeros_like(cnt), where=N[:, None] > 0)[:, None, :]
        tree.__setstate__(dict(max_depth=int(nd["depth"].max()), node_count=len(nd), nodes=rec, values=np.ascontiguousarray(values)))
        return tree

    def to_sklearn(self):
        """-> sklearn.ensemble.RandomForestClassifier (or RandomForestRegressor for criterion "squared_error") whose
        estimators_ wrap the converted trees."""
        from sklearn.ensemble import RandomForestClassifier
        from sklearn.tree import DecisionTreeClassifier

        if self.criterion == "squared_error":
            from sklearn.ensemble import RandomForestRegressor
            from sklearn.tree import DecisionTreeRegressor

            ests = []
            for t in range(len(self.nodes)):
                e = DecisionTreeRegressor()
                e.tree_ = self.sklearn_tree(t)
                e.n_features_in_, e.n_outputs_, e.max_features_ = self.n_features, 1, self.n_features
                ests.append(e)
            rf = RandomForestRegressor(n_estimators=len(ests))
            rf.estimators_ = ests
            rf.n_features_in_, rf.n_outputs_ = self.n_features, 1
            return rf

        C = len(self.classes)
        ests = []
        for t in range(len(self.nodes)):
            e = DecisionTreeClassifier()
            e.tree_ = self.sklearn_tree(t)
            e.n_features_in_, e.n_outputs_, e.n_classes_, e.classes_, e.max_features_ = self.n_features, 1, C, self.classes, self.n_features
            ests.append(e)
        rf = RandomForestClassifier(n_estimators=len(ests))
        rf.estimators_ = ests
        rf.n_features_in_, rf.n_outputs_, rf.n_classes_, rf.classes_ = self.n_features, 1, C, self.classes
        return rf

    # ---- host-side apply (exact integer votes, the arithmetic of DESIGN.md 2.7)
    def predict_votes(self, X) -> np.ndarray:
        rows = binning.bin_rows(np.asarray(X, dtype=np.float32), self.edges, self.n_edges)
        votes = np.zeros((rows.shape[0], len(self.classes)), dtype=np.uint64)
        for nd, cnt in zip(self.nodes, self.counts):
            leaf = self.apply_tree(nd, rows)
            c = cnt[leaf].astype(np.uint64)
            N = c.sum(1, keepdims=True)
            votes += np.where(N > 0, (c << np.uint64(16)) // np.maximum(N, 1), 0)
        return votes.astype(np.uint32)

    @staticmethod
    def apply_tree(nd, rows) -> np.ndarray:
        cur = np.zeros(rows.shape[0], dtype=np.int64)
        active = nd["feature"][cur] >= 0
        idx = np.arange(rows.shape[0])
        while active.any():
            i = idx[active]
            n = cur[i]
            cur[i] = np.where(rows[i, nd["feature"][n]] <= nd["bin"][n], nd["left"][n], nd["right"][n])
            active = nd["feature"][cur] >= 0
        return cur


def load_forest(path: str) -> ExportedForest:
    z = np.load(path, allow_pickle=False)
    if int(z["format_version"]) != FORMAT_VERSION:
        raise ValueError(f"{path}: format version {int(z['format_version'])}, expected {FORMAT_VERSION}")
    off = z["tree_offsets"]
    nodes = [z["nodes"][off[i]:off[i + 1]] for i in range(len(off) - 1)]
    counts = [z["counts"][off[i]:off[i + 1]] for i in range(len(off) - 1)]
    return ExportedForest(nodes, counts, z["edges"], z["n_edges"], z["classes"], int(z["n_features"]), str(z["criterion"]),
                          float(z["target_lo"]) if "target_lo" in z else 0.0, float(z["target_span"]) if "target_span" in z else 1.0)
