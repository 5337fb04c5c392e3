import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def built_library():
    """libedf_b200.so, (re)built when nvcc is present and the sources are newer."""
    from elasticdeform_b200 import build
    return build.build_library()


@pytest.fixture(scope="session")
def oracle_ref_built():
    """oracle/_ref (the compiled unmodified reference); built here when /root/reference exists."""
    from oracle import build_ref
    return build_ref.build()
