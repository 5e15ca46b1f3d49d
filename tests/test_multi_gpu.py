"""Multi-GPU exactness on hardware: runs whenever >= 2 CUDA devices are visible (gpurun --gpus N).
tools/dist_check.py under torchrun: every rank builds its shard of the (fold, tree) units, ONE NCCL all-gather of the
votes, and votes / s(x) / IBP masks must equal the single-rank CPU oracle bit for bit -- through the in-library NCCL path
and through the host-side exchange.  The log is kept under gpurun_out/ so the run leaves evidence."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch

    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("world", [2, 4, 8])
def test_dist_check_bit_exact(world):
    n = _gpus()
    if n < world:
        pytest.skip(f"{n} GPU(s) visible, {world} needed")
    port = 29500 + world * 7 + os.getpid() % 50
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "dist_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"dist_check_world{world}.log"), "w") as f:
        f.write(r.stdout + "\n--- stderr ---\n" + r.stderr)
    assert r.returncode == 0, r.stderr[-2000:]
    assert f"dist_check ok: world={world}" in r.stdout
