import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")   # the reference's iteration counts depend on it


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden_ba():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "ba_golden.npz"))


@pytest.fixture(scope="session")
def golden_bd():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "blockdiag_golden.npz"))


@pytest.fixture(scope="session")
def golden_dlt():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "dlt_golden.npz"))
