import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "r-super_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA (sm_100a) device; run with -m gpu on the B200 box")
    config.addinivalue_line("markers", "staged: GPU test written after the round's GPU budget was spent — never run on "
                                       "hardware yet; reported as XPASS / XFAIL (non-strict) instead of stopping the suite")


def pytest_collection_modifyitems(config, items):
    """`staged` tests exercise kernels that compile for sm_100a but have not had their first run on a B200 (no GPU in the
    build container, gpurun budget exhausted).  They run with the suite, last, as non-strict xfail: a pass shows up as
    XPASS, a failure as XFAIL with its traceback under -rx, and neither masks the verified tests.  Remove the marker from
    a test once it has been seen green on hardware."""
    for item in items:
        if item.get_closest_marker("staged") is not None:
            item.add_marker(pytest.mark.xfail(strict=False, reason="staged: first run on a B200 pending"))


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    path = os.path.join(ROOT, "tests", "golden", "reference_outputs.npz")
    return np.load(path, allow_pickle=False)


@pytest.fixture(scope="session")
def cuda_dev():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return torch.device("cuda:0")
