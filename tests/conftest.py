import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200, sm_100a)")
    # tests build random-init models on purpose; the product warns about that (the reference starts from ImageNet weights)
    config.addinivalue_line("filterwarnings", "ignore:CutoffVGG16. no pretrained VGG16 weights supplied")


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """Build libabl_b200.so if it is missing or stale (nvcc cross-compiles without a GPU)."""
    from ab_line_classifier_b200 import build
    build.build()


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container (run with gpurun)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
