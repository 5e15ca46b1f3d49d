"""Latency probe for the benchmark-sized configs (VSI-shaped TsT, IBP, grouped forests)."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "test-set-training_b200"))
import numpy as np, torch, tst
from tst import synth

X, y = synth.vsi_table()
for i in range(4):
    torch.cuda.synchronize(); t = time.perf_counter()
    r = tst.run_tst(X, y, k=5, n_trees=100, seed=i)
    dt = time.perf_counter() - t
    print("run_tst ms", round(dt * 1e3, 2), {k: (round(v, 2) if isinstance(v, float) else v) for k, v in r.stats.items() if k.startswith("ms_") or k in ("levels", "kernel_launches")})
for i in range(3):
    torch.cuda.synchronize(); t = time.perf_counter()
    r = tst.ibp(X, y, rounds=10, k=5, n_trees=200, tau=0.5, batch=250, seed=0, want_s=False)
    dt = time.perf_counter() - t
    print("ibp ms", round(dt * 1e3, 2), r.removed_per_round.tolist(), {k: (round(v, 2) if isinstance(v, float) else v) for k, v in r.stats.items() if k.startswith("ms_") or k in ("levels", "kernel_launches", "trees")})
for kind in ("cvbench", "mmmu", "videomme"):
    Xg, yg, g, ng = synth.grouped_table(kind)
    for i in range(2):
        torch.cuda.synchronize(); t = time.perf_counter()
        r = tst.run_tst(Xg, yg, k=5, n_trees=100, seed=i, group=g)
        dt = time.perf_counter() - t
    print(kind, "grouped run_tst ms", round(dt * 1e3, 2), "trees", r.stats["trees"], "acc", np.nanmean(r.fold_accuracy).round(3))
