import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "ttl-test-time-low-rank-adaptation_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")
    config.addinivalue_line("markers", "slow: long CPU test")


@pytest.fixture(scope="session")
def b16_weights():
    from oracle import ttl_oracle as O
    return O.make_synthetic_weights(O.ARCHS["ViT-B/16"], 1234)


@pytest.fixture(scope="session")
def b16_views():
    from oracle import ttl_oracle as O
    return O.make_synthetic_views(64, 224, seed=7)
