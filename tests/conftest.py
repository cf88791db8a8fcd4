import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def gpu_ctx():
    """One GPU context for the whole session (fails loudly when the extension/GPU is missing)."""
    import optimet_b200 as ob
    ctx = ob.Context(0)
    yield ctx
    ctx.close()


@pytest.fixture(autouse=True)
def _oracle_threads_reset():
    """The oracle's thread count is process-wide; the reference's AMOS (oracle/_ref) is not thread-safe: every test starts
    and ends single-threaded unless it asks otherwise."""
    yield
    try:
        from oracle import oracle as O
        O.set_threads(1)
    except Exception:
        pass
