"""Where the end-to-end step of the 1M x 64 job spends its time outside the fit: python tools/e2e_probe.py (on the GPU box)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "test-set-training_b200"))
import numpy as np, torch
import tst
from tst import synth

eng = tst.Engine(0)
X, y = synth.scaleup_table()
Xp = torch.from_numpy(X).pin_memory(); yp = torch.from_numpy(y).pin_memory()
def T(f, n=3):
    f(); b = 1e9
    for _ in range(n):
        torch.cuda.synchronize(); t = time.perf_counter(); r = f(); torch.cuda.synchronize(); b = min(b, time.perf_counter() - t)
    return b * 1e3, r
ms, _ = T(lambda: Xp.cuda(non_blocking=True)); print(f"H2D X 256 MB from pinned: {ms:.2f} ms ({0.256 / ms * 1e3:.1f} GB/s)")
ms, _ = T(lambda: eng.bin(Xp.numpy())); print(f"eng.bin (H2D + edges + map): {ms:.2f} ms")
ms, r = T(lambda: tst.run_tst(Xp.numpy(), yp.numpy(), k=10, n_trees=64, seed=1, engine=eng), 2)
st = r.stats
print(f"run_tst 10 x 64: {ms:.1f} ms wall; device phases: prep {st['ms_prep']:.1f} big {st['ms_grow_big']:.1f} wide {st['ms_grow_wide']:.1f} mid {st['ms_grow_mid']:.1f} small {st['ms_grow_small']:.1f} tiny {st['ms_grow_tiny']:.1f} infer {st['ms_infer']:.1f}"
      f" = {sum(st[k] for k in ('ms_prep','ms_grow_big','ms_grow_wide','ms_grow_mid','ms_grow_small','ms_grow_tiny','ms_infer')):.1f}")
eng.set_option("TRACE", 1)
tst.run_tst(Xp.numpy(), yp.numpy(), k=10, n_trees=64, seed=1, engine=eng)
