import os, sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/test-set-training_b200')
import numpy as np, torch, tst
from tst import synth
X, y = synth.vsi_table()
eng = tst.Engine(0)
tst.run_tst(X, y, k=5, n_trees=100, seed=0, engine=eng)
best = 1e9
for i in range(5):
    torch.cuda.synchronize(); t = time.perf_counter()
    r = tst.run_tst(X, y, k=5, n_trees=100, seed=1 + i, engine=eng)
    best = min(best, time.perf_counter() - t)
tst.ibp(X, y, rounds=2, k=5, n_trees=200, tau=0.5, batch=250, seed=0, want_s=False, engine=eng)
bi = 1e9
for i in range(3):
    torch.cuda.synchronize(); t = time.perf_counter()
    rb = tst.ibp(X, y, rounds=10, k=5, n_trees=200, tau=0.5, batch=250, seed=0, want_s=False, engine=eng)
    bi = min(bi, time.perf_counter() - t)
print(os.environ.get("TST_THIN_NODES"), os.environ.get("TST_BLIND_LEVELS"), "vsi ms %.2f ibp ms %.2f levels %d launches %d" % (best * 1e3, bi * 1e3, r.stats["levels"], r.stats["kernel_launches"]))
