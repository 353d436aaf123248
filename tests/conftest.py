import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # The CPU-side checkers (oracle, reference goldens) run tiny tensors: with one intra-op thread per core the suite is
    # 2x slower on an idle 8-core box (oracle pipeline tests 27 s vs 12 s) and many times slower on a loaded one.
    import torch
    torch.set_num_threads(min(4, os.cpu_count() or 1))


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    class G:
        def __getitem__(self, name):
            return np.load(os.path.join(GOLDEN, name + ".npz"))
    return G()
