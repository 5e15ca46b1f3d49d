"""world_size = 2 on CPU (gloo): the N > 1 host logic -- tree sharding by (part * n_trees + t) % world, one
all-gather of the per-rank vote tensors, rank-ordered integer sum -- gives exactly the single-rank votes.
The per-rank votes come from the CPU oracle here (no GPU in this container); on the GPU box the same
host code runs over NCCL (bench.py --gpus N, tests/test_gpu_parity.py::test_rank_sharding_sums_to_single_rank)."""
import os
import socket
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "test-set-training_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist

    import tst
    from oracle import hist_rf as H
    from tst import synth

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    assert tst.api.dist_info() == (rank, world)
    X, y = synth.vsi_table(600, 12, 3, seed=2)
    edges, ne = H.find_edges(X)
    bins = H.bin_table(X, edges, ne)
    k, T, C = 4, 6, 3
    fold = H.assign_folds(y, np.ones(len(y), np.uint8), k, C)
    imp = np.zeros(13, np.uint64)
    votes, st, _ = H.fit_oof(bins, y, fold, k, T, C, H.make_params(12), seed=9, rank=rank, world=world, nthreads=1, importances=imp)
    assert st["trees"] == len(tst.shard_trees(k, T, rank, world))  # the host's shard list is the engine's
    # feature importances: every rank holds integer sums over ITS trees; one all-reduce of d + 1 integers gives the forest's
    ti = torch.from_numpy(imp.view(np.int64).copy())
    dist.all_reduce(ti)
    np.save(os.path.join(out_dir, f"imp_{rank}.npy"), ti.numpy().view(np.uint64))
    mine = torch.from_numpy(votes.view(np.int32).copy())
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)  # the path's one exchange step
    total = torch.stack(gathered).sum(0, dtype=torch.int64).numpy().astype(np.uint32)  # rank order, integers
    np.save(os.path.join(out_dir, f"votes_{rank}.npy"), total)
    dist.barrier()
    dist.destroy_process_group()


def test_world2_votes_equal_single_rank(tmp_path):
    import torch.multiprocessing as mp

    from oracle import hist_rf as H
    from tst import synth

    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    X, y = synth.vsi_table(600, 12, 3, seed=2)
    edges, ne = H.find_edges(X)
    bins = H.bin_table(X, edges, ne)
    fold = H.assign_folds(y, np.ones(len(y), np.uint8), 4, 3)
    imp = np.zeros(13, np.uint64)
    ref = H.fit_oof(bins, y, fold, 4, 6, 3, H.make_params(12), seed=9, importances=imp)[0]
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"votes_{r}.npy"), ref), r
        assert np.array_equal(np.load(tmp_path / f"imp_{r}.npy"), imp), r  # integer sums: the same on any number of ranks
