"""Where does an end-to-end run_tst step on the 1M x 64 table spend its time?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "test-set-training_b200"))
import numpy as np, torch, tst
from tst import synth

X, y = synth.scaleup_table()
Xp = torch.from_numpy(X).pin_memory()
eng = tst.Engine(0)
classes, yi = tst.api.encode_labels(y)
def T(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t = time.perf_counter(); fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t)
    return round(best * 1e3, 2)
print("H2D X pinned ms", T(lambda: Xp.to("cuda", non_blocking=True)))
Xd = Xp.to("cuda")
print("bin (device edges) ms", T(lambda: eng.bin(Xd)))
print("bin (given edges) ms", T(lambda: eng.bin(Xd, edges=eng.edges, n_edges=eng.n_edges)))
print("encode_labels (host) ms", T(lambda: tst.api.encode_labels(y)))
y_dev = eng.to_dev(yi, torch.uint8)
act = torch.ones(len(y), dtype=torch.uint8, device="cuda")
print("assign_folds ms", T(lambda: eng.assign_folds(y_dev, act, 10, 4)))
fold = eng.assign_folds(y_dev, act, 10, 4)
prm = tst.make_params(64)
votes = eng.new_votes(4)
print("fit 320 trees ms", T(lambda: eng.fit_oof(y_dev, fold, 10, 1000, 4, prm, seed=0, t0=0, t1=32, votes=votes), 2))
print("fit 10 trees ms (fixed costs)", T(lambda: eng.fit_oof(y_dev, fold, 10, 1000, 4, prm, seed=0, t0=0, t1=1, votes=votes), 2))
print("finalize + D2H ms", T(lambda: [t.cpu() for t in eng.finalize(votes, y_dev, fold, 10, 1000, 4)] + [votes.cpu(), fold.cpu()]))
