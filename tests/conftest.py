import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    if have_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="module", params=["default", "block-only"])
def engine(request):
    """The C-ABI engine in its two kernel configurations: default (short logs on the warp-per-log kernel, the rest and its
    deferrals on the CTA-per-log kernel) and with the warp-per-log bin disabled — results must be identical.  Modules that
    define their own `engine` fixture run in the default configuration only."""
    from peritext_b200.engine import BatchEngine
    if request.param == "block-only":
        os.environ["PT_WARP"] = "0"
    else:
        os.environ.pop("PT_WARP", None)
    e = BatchEngine(0)
    yield e
    e.close()
    os.environ.pop("PT_WARP", None)
