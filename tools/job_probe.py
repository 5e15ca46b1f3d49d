"""Whole-job wall-clock (10-fold x 1000 trees at 1M x 64, host arrays in -> s(x) out) for several numbers of trees in flight."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "test-set-training_b200"))
import numpy as np, torch
import tst
from tst import synth

X, y = synth.scaleup_table(1_000_000, 64, 4, seed=0)
Xp = torch.from_numpy(X).pin_memory()
eng = tst.Engine(0)
tst.run_tst(Xp, y, k=10, n_trees=1000, seed=0, tree_range=(0, 8), engine=eng)
for tif in [int(a) for a in sys.argv[1:]] or [0, 640, 1024]:
    torch.cuda.synchronize(); t = time.perf_counter()
    r = tst.run_tst(Xp, y, k=10, n_trees=1000, seed=0, engine=eng, trees_in_flight=tif)
    dt = time.perf_counter() - t
    print(f"trees_in_flight={tif}: {dt:.2f} s, {10000 / dt:.0f} trees/s, levels {r.stats['levels']}, launches {r.stats['kernel_launches']}", flush=True)
