"""Multi-GPU correctness check, run under torchrun on N GPUs of one box:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/dist_check.py
Every rank builds its shard; one NCCL all-gather; the result must equal the single-rank CPU oracle bit for bit."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "test-set-training_b200"))
import numpy as np, torch, torch.distributed as dist
import tst
from oracle import hist_rf as H
from tst import synth

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
X, y = synth.vsi_table(3000, 24, 4, seed=4)
res = tst.run_tst(X, y, k=5, n_trees=23, seed=6)
edges, ne = H.find_edges(X)
bins = H.bin_table(X, edges, ne)
fold = H.assign_folds(y, np.ones(len(y), np.uint8), 5, 4)
votes, st, _ = H.fit_oof(bins, y, fold, 5, 23, 4, H.make_params(24), seed=6)
s, _, _ = H.finalize(votes, y, fold, 5, 23)
assert np.array_equal(res.oof_votes, votes), f"rank {rank}: votes differ"
assert np.array_equal(res.s.view(np.uint32), s.view(np.uint32))
assert res.stats["trees"] == len(tst.shard_trees(5, 23, rank, world))
# same through the host-side exchange (torch.distributed all_gather + tst_reduce_votes)
eng_host = tst.Engine(local, rank, world, lib_nccl=False)
res2 = tst.run_tst(X, y, k=5, n_trees=23, seed=6, engine=eng_host)
assert np.array_equal(res2.oof_votes, votes) and not eng_host.lib_nccl and tst.get_engine().lib_nccl
r2 = tst.ibp(X, y, rounds=3, k=5, n_trees=10, tau=0.45, batch=100, seed=2, engine=eng_host)
r = tst.ibp(X, y, rounds=3, k=5, n_trees=10, tau=0.45, batch=100, seed=2)
assert np.array_equal(r.keep_mask, r2.keep_mask) and r.removed_per_round.tolist() == r2.removed_per_round.tolist()
oact, orem, os_ = H.ibp(bins, y, 5, 10, 4, H.make_params(24), seed=2, rounds=3, tau=0.45, batch=100)
assert r.removed_per_round.tolist() == orem.tolist() and np.array_equal(r.keep_mask, oact.astype(bool))
assert np.array_equal(r.s_per_round.view(np.uint32), os_.view(np.uint32))
dist.barrier()
if rank == 0:
    print(f"dist_check ok: world={world}, votes / s(x) / IBP bit-exact vs the single-rank CPU oracle (in-library NCCL and host-side exchange)")
dist.destroy_process_group()
