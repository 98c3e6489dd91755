import os
import sys

import pytest

# tests run on the seeded synthetic stand-in weights when the upstream .pt files are absent (weights.py); the product
# entry point run() never sets this
os.environ.setdefault("TRUELY_ALLOW_SYNTHETIC", "1")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def analyzer():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    import truely_b200  # noqa: F401
    from truely_b200.model import Analyzer
    return Analyzer(device=0)


@pytest.fixture(scope="session")
def analyzer_simt():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    import truely_b200  # noqa: F401
    from truely_b200.model import Analyzer
    return Analyzer(device=0, facenet_impl=1)
