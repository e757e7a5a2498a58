import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden"), ROOT,
          os.path.join(ROOT, "pytorch-detect-to-track_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import cpu
    cpu.build()
    return cpu


@pytest.fixture(scope="session")
def golden_rpn():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "rpn_reference.npz"))


@pytest.fixture(scope="session")
def golden_cuda():
    import numpy as np
    path = os.path.join(ROOT, "tests", "golden", "ref_cuda.npz")
    if not os.path.exists(path):
        pytest.skip("tests/golden/ref_cuda.npz not generated yet (tests/golden/make_golden_gpu.py on a B200)")
    return np.load(path)
