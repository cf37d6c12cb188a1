import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run on the GPU box with -m gpu)")


@pytest.fixture(scope="session")
def cuda_lib():
    """Builds/loads libb200seg.so and pins fp32 reference math (no TF32) for GPU comparisons."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from dasemanticsegmentationaml_b200 import _lib, build
    build.build()
    _lib.ensure_device(0)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return _lib.lib()
