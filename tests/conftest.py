import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    """tests/golden/klang_ref_*.npz — outputs of the compiled reference (tests/gen_golden.py)."""
    import numpy as np
    g = {}
    for fs in (44100, 48000):
        path = os.path.join(ROOT, "tests", "golden", f"klang_ref_fs{fs}.npz")
        with np.load(path) as z:
            g[fs] = {k: z[k] for k in z.files}
    return g
