import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built_oracle():
    from oracle import oracle
    oracle.build()


# Test bodies written against tests/backends.py run on the CPU oracle (here, `-m "not gpu"`)
# and on the CUDA path through the C-ABI (`-m gpu`, on the B200 box).
@pytest.fixture(params=["oracle", pytest.param("b200", marks=pytest.mark.gpu)])
def backend(request):
    return request.param
