import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ctx():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import hmse_b200
    return hmse_b200.default_context(0)


@pytest.fixture(scope="session")
def corpus8():
    """8 MiB of the seed-42 procedural wiki corpus (oracle twin generator)."""
    from oracle import corpus
    return corpus.generate(8 << 20)


@pytest.fixture(autouse=True)
def _scratch_guard_bands(request):
    """HMSE_GUARD=1 (checked run of the GPU suite: compute-sanitizer is closed on the B200 pool): after every GPU test the
    guard bands around every scratch slot of the shared context must be intact."""
    yield
    import os
    if os.environ.get("HMSE_GUARD") and request.node.get_closest_marker("gpu") is not None:
        import torch
        if torch.cuda.is_available():
            import hmse_b200
            c = hmse_b200.default_context(0)
            c.check(c.lib.hmse_guard_check(c.h))
