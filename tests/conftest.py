import json
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # the torch statements the kernels are checked against must be true fp32 (no TF32 in cuDNN / cuBLAS)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def layouts():
    with open(os.path.join(GOLDEN, "state_dict_layouts.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def built_lib():
    """Compile (if stale) and load libdv_b200.so — on CPU this only checks that it builds, loads and exports."""
    from dalle2_video_b200 import build as b
    b.build()
    from dalle2_video_b200 import _lib
    return _lib


def golden(name):
    return torch.load(os.path.join(GOLDEN, name), weights_only=False)
