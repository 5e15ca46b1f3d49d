import os, sys, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/test-set-training_b200")
import numpy as np, torch, tst
from tst import synth
eng = tst.Engine(0)
X, y = synth.vsi_table()
f = lambda: tst.ibp(X, y, rounds=10, k=5, n_trees=200, tau=0.5, batch=250, seed=0, want_s=False, engine=eng)
f(); f()
torch.cuda.synchronize(); t = time.perf_counter(); r = f(); print("ibp ms", 1e3 * (time.perf_counter() - t))
eng.set_option("TRACE", 1)
f()
