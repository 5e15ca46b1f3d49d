"""Small end-to-end run for compute-sanitizer (one tool per gpurun call):
    compute-sanitizer --tool memcheck  python tools/sanitize_run.py
    compute-sanitizer --tool racecheck python tools/sanitize_run.py
Covers every kernel: device edge selection, binning, folds, partitions, bootstrap, the three growth tiers
(generic and packed), OOF inference, finalize, IBP selection, kept-forest predict/export."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "test-set-training_b200"))
import numpy as np
import tst
from oracle import hist_rf as H
from tst import synth

X, y = synth.vsi_table(1500, 16, 4, seed=1)          # roots > 256 rows: CTA kernel + packed warp kernels
r = tst.run_tst(X, y, k=3, n_trees=4, seed=0)
edges, ne = H.find_edges(X); bins = H.bin_table(X, edges, ne)
fold = H.assign_folds(y, np.ones(len(y), np.uint8), 3, 4)
votes, _, _ = H.fit_oof(bins, y, fold, 3, 4, 4, H.make_params(16), seed=0)
assert np.array_equal(r.oof_votes, votes)
X7, y7 = synth.vsi_table(700, 12, 7, seed=2)         # 7 classes: generic (unpacked) warp kernels
r7 = tst.run_tst(X7, y7, k=3, n_trees=3, seed=1)
e7, n7 = H.find_edges(X7); b7 = H.bin_table(X7, e7, n7)
f7 = H.assign_folds(y7, np.ones(len(y7), np.uint8), 3, 7)
v7, _, _ = H.fit_oof(b7, y7, f7, 3, 3, 7, H.make_params(12), seed=1)
assert np.array_equal(r7.oof_votes, v7)
ri = tst.ibp(X, y, rounds=2, k=3, n_trees=3, tau=0.4, batch=50, seed=3)
oa, orm, _ = H.ibp(bins, y, 3, 3, 4, H.make_params(16), seed=3, rounds=2, tau=0.4, batch=50)
assert ri.removed_per_round.tolist() == orm.tolist() and np.array_equal(ri.keep_mask, oa.astype(bool))
Xg, yg, g, ng = synth.grouped_table("mmmu", groups=6)
rg = tst.run_tst(Xg, yg, k=3, n_trees=3, seed=2, group=g)
clf = tst.TsTForest(n_estimators=3, random_state=1).fit(X, y)
p = clf.predict_proba(X[:100]); clf.export_tree(1); clf.feature_importances_
print("sanitize_run ok")
