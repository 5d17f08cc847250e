import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box: pytest -m gpu)")


def _cuda_available():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def engine_factory():
    """Creates dumux_b200 engines on cuda:0; fails loudly (no fallback) if the CUDA library is missing."""
    from dumux_b200.binding import Engine
    made = []

    def make(spec=None):
        e = Engine(spec, device=0)
        made.append(e)
        return e

    yield make
    for e in made:
        e.close()


def pytest_collection_modifyitems(config, items):
    # gpu tests are selected with -m gpu on the GPU box; nothing is skipped silently there.
    if not _cuda_available():
        skip = pytest.mark.skip(reason="no CUDA device in this container (gpu tests run under gpurun)")
        for item in items:
            if "gpu" in item.keywords:
                item.add_marker(skip)
