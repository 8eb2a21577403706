import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) GPU; run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def ctx():
    """One dgp_ctx for the whole GPU session (fails loudly without the CUDA library / a GPU)."""
    import dynaml_b200 as dg
    return dg.default_context()
