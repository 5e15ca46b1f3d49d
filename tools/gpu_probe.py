"""Ad-hoc timing probe (not the bench contract): python tools/gpu_probe.py [n] [d] [k] [trees_per_fold]."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "test-set-training_b200"))

import numpy as np
import torch

import tst
from tst import synth


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    d = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    k = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    T = int(sys.argv[4]) if len(sys.argv) > 4 else 4
    reps = int(sys.argv[5]) if len(sys.argv) > 5 else 2
    X, y = (synth.scaleup_table(n, d) if n > 20000 else synth.vsi_table(n, d))
    eng = tst.Engine(0)
    t = time.perf_counter()
    edges, ne = tst.binning.table_edges(X)
    t_edges = time.perf_counter() - t
    t = time.perf_counter()
    eng.bin(X, edges=edges, n_edges=ne)
    torch.cuda.synchronize()
    t_bin = time.perf_counter() - t
    classes, yi = tst.api.encode_labels(y)
    C = len(classes)
    y_dev = eng.to_dev(yi, torch.uint8)
    active = torch.ones(n, dtype=torch.uint8, device=eng.dev)
    fold = eng.assign_folds(y_dev, active, k, C)
    prm = tst.make_params(d)
    out = dict(n=n, d=d, k=k, T=T, t_edges=t_edges, t_bin=t_bin)
    for rep in range(reps):
        votes = eng.new_votes(C)
        eng.stats(reset=True)
        torch.cuda.synchronize()
        t = time.perf_counter()
        eng.fit_oof(y_dev, fold, k, T, C, prm, seed=rep, votes=votes)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t
        st = eng.stats()
        out[f"rep{rep}"] = dict(seconds=dt, trees_per_s=k * T / dt, **st)
        print(json.dumps(out[f"rep{rep}"]), flush=True)
    s, cor, tot = eng.finalize(votes, y_dev, fold, k, T, C)
    out["acc"] = (cor.float() / tot.float()).cpu().numpy().tolist()
    print(json.dumps(out))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"probe_{n}x{d}.json"), "w") as f:
        json.dump(out, f)


if __name__ == "__main__":
    main()
