"""More seeds of the randomised GPU parity tests than the suite runs (on the GPU box): python tools/fuzz_parity.py [n_small] [n_large]"""
import os, sys, time, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "test-set-training_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import tst
import test_gpu_parity as T

n_small = int(sys.argv[1]) if len(sys.argv) > 1 else 150
n_large = int(sys.argv[2]) if len(sys.argv) > 2 else 40
eng = tst.Engine(0)
bad = 0
t0 = time.time()
for mode in ("auto", "deep"):
    knobs = dict(SMALL_MULTI=1, BIG_LITE=1, MID_CLASSES=3) if mode == "deep" else {}
    for kk, v in knobs.items():
        eng.set_option(kk, v)
    for seed in range(24, 24 + n_small):
        try:
            T.test_random_configurations_bit_exact(eng, mode, seed)
        except Exception:
            bad += 1
            print(f"FAIL small mode={mode} seed={seed}\n{traceback.format_exc()[-1500:]}", flush=True)
    for kk in knobs:
        eng.set_option(kk, -1)
for seed in range(16, 16 + n_large):
    try:
        T.test_random_large_configurations_bit_exact(eng, seed)
    except Exception:
        bad += 1
        print(f"FAIL large seed={seed}\n{traceback.format_exc()[-1500:]}", flush=True)
print(f"fuzz: {2 * n_small} small + {n_large} large configurations, {bad} failures, {time.time() - t0:.0f} s")
sys.exit(1 if bad else 0)
