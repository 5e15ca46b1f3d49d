import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "test-set-training_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")
    config.addinivalue_line("markers", "slow: statistical gate vs sklearn (tens of seconds on CPU)")


@pytest.fixture(scope="session")
def vsi_small():
    """1,200 x 16 VSI-shaped table binned with the oracle's edges."""
    import numpy as np
    from oracle import hist_rf as H
    from tst import synth

    X, y = synth.vsi_table(1200, 16, 4, seed=7)
    edges, ne = H.find_edges(X)
    bins = H.bin_table(X, edges, ne)
    fold = H.assign_folds(y, np.ones(len(y), np.uint8), 5, 4)
    return dict(X=X, y=y, edges=edges, n_edges=ne, bins=bins, fold=fold, C=4, k=5)
