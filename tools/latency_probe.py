"""Latency of the benchmark-sized configurations (BASELINE.json configs 1-3) through the public API; knobs from the environment
(TST_SUB_ROWS, TST_THIN_NODES, TST_BLIND_LEVELS, TST_TRACE ...).  python tools/latency_probe.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "test-set-training_b200"))
import numpy as np, torch
import tst
from tst import synth


def best_of(f, n=5):
    f()
    b = 1e9
    for _ in range(n):
        torch.cuda.synchronize()
        t = time.perf_counter()
        r = f()
        b = min(b, time.perf_counter() - t)
    return b * 1e3, r


eng = tst.Engine(0)
X, y = synth.vsi_table()
ms, r = best_of(lambda: tst.run_tst(X, y, k=5, n_trees=100, seed=1, engine=eng))
print(f"vsi 5x100: {ms:.2f} ms, launches {r.stats['kernel_launches']}, levels {r.stats['levels']}, "
      f"ms big {r.stats['ms_grow_big']:.2f} mid {r.stats['ms_grow_mid']:.2f} small {r.stats['ms_grow_small']:.2f} sub {r.stats['ms_grow_sub']:.2f} tiny {r.stats['ms_grow_tiny']:.2f} prep {r.stats['ms_prep']:.2f} infer {r.stats['ms_infer']:.2f}")
# where the host time goes outside the fit
torch.cuda.synchronize(); t = time.perf_counter(); eng.bin(X); torch.cuda.synchronize(); print(f"  bin (H2D + edges + map): {1e3 * (time.perf_counter() - t):.2f} ms")
ms, r = best_of(lambda: tst.ibp(X, y, rounds=10, k=5, n_trees=200, tau=0.5, batch=250, seed=0, want_s=False, engine=eng), 3)
print(f"  ibp 10 x 5 x 200: {ms:.2f} ms")
for kind in ("cvbench", "mmmu", "videomme"):
    Xg, yg, gg, ng = synth.grouped_table(kind)
    ms, r = best_of(lambda: tst.run_tst(Xg, yg, k=5, n_trees=100, seed=1, group=gg, engine=eng))
    print(f"  {kind}: {ms:.2f} ms, launches {r.stats['kernel_launches']}")
Xn, yn = synth.numeric_table(5000, 32, seed=0)
ms, r = best_of(lambda: tst.run_tst_regression(Xn, yn, k=5, n_trees=100, seed=1, engine=eng))
print(f"  regression 5k x 32, 5 x 100, all features per split: {ms:.2f} ms, launches {r.stats['kernel_launches']}, levels {r.stats['levels']}, "
      f"ms big {r.stats['ms_grow_big']:.2f} wide {r.stats['ms_grow_wide']:.2f} mid {r.stats['ms_grow_mid']:.2f} small {r.stats['ms_grow_small']:.2f} tiny {r.stats['ms_grow_tiny']:.2f}")
if os.environ.get("TST_TRACE_ONE"):
    eng.set_option("TRACE", 1)
    tst.run_tst(X, y, k=5, n_trees=100, seed=1, engine=eng)
