"""Where do the size-class queues of the warp-per-node tier start to pay?  Fit time with MID_CLASSES 0 / 3 over job sizes.
python tools/class_crossover.py   (on the GPU box)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "test-set-training_b200"))
import numpy as np, torch
import tst
from tst import synth

eng = tst.Engine(0)
for n, d, T in ((5000, 32, 100), (20000, 32, 20), (20000, 32, 64), (50000, 32, 64), (100000, 32, 64), (200000, 64, 32)):
    X, y = synth.scaleup_table(n, d)
    row = []
    for c in (0, 3, 0, 3):
        eng.set_option("MID_CLASSES", c)
        tst.run_tst(X, y, k=5, n_trees=T, seed=1, engine=eng)
        best = 1e9
        for i in range(3):
            torch.cuda.synchronize(); t = time.perf_counter()
            r = tst.run_tst(X, y, k=5, n_trees=T, seed=1, engine=eng)
            best = min(best, time.perf_counter() - t)
        row.append(f"c{c} {best * 1e3:.2f} ms (mid {r.stats['ms_grow_mid']:.2f})")
    print(f"n {n} d {d} 5 x {T} trees, in-bag rows per batch ~{5 * T * n * 0.8 * 0.632 / 1e6:.1f}M:", " | ".join(row), flush=True)
