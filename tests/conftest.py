import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "python-qinfer_b200"), os.path.join(ROOT, "oracle"),
          os.path.join(ROOT, "tests", "golden"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "refcheck: needs /root/reference (build container only)")


def pytest_collection_modifyitems(config, items):
    import torch
    have_gpu = torch.cuda.is_available()
    have_ref = os.path.isdir("/root/reference/src/qinfer")
    for item in items:
        if "gpu" in item.keywords and not have_gpu:
            item.add_marker(pytest.mark.skip(reason="no CUDA device"))
        if "refcheck" in item.keywords and not have_ref:
            item.add_marker(pytest.mark.skip(reason="reference tree not present"))


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))
    return load


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """Make sure the C-ABI library exists (nvcc cross-compiles without a GPU)."""
    sys.path.insert(0, os.path.join(ROOT, "python-qinfer_b200"))
    import build as qb_build
    qb_build.build_library()
    yield
