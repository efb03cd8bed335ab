import os
import sys

import pytest

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def port():
    """CPU oracle: our restatement (oracle/libpainty_oracle.so)."""
    from oracle import cpu

    cpu.build()
    return cpu.Cpu("port")


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference headers (oracle/_ref); skipped where it was never built."""
    from oracle import cpu

    cpu.build()
    if not cpu.have_ref():
        pytest.skip("oracle/_ref/libpainty_ref.so not present (needs /root/reference at build time)")
    return cpu.Cpu("ref")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    return dict(np.load(os.path.join(ROOT, "tests", "golden", "ref_golden.npz")))


@pytest.fixture(scope="session")
def built_lib():
    from painty_b200 import build

    return build.build()


@pytest.fixture(scope="session")
def ctx32(built_lib):
    from painty_b200 import api

    return api.Context(0, api.F32)


@pytest.fixture(scope="session")
def ctx64(built_lib):
    from painty_b200 import api

    return api.Context(0, api.F64)
