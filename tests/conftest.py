import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'video-long-term-feature-banks_b200', 'lib')
for p in (LIB, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


@pytest.fixture(autouse=True)
def _fresh_cfg():
    from core import config as C
    C.reset_cfg()
    yield
    C.reset_cfg()
