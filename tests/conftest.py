import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (runs on the B200 box)")


@pytest.fixture(scope="session")
def synth():
    from deepflow_b200.synth import Synth
    s = Synth(seed=20260921, n_pods=2000)
    yield s
    s.close()


@pytest.fixture(scope="session")
def built():
    """Build the test-side libraries once (oracle + generator)."""
    from deepflow_b200 import build
    build.build_synth()
    build.build_oracle()
    return True
