import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """The C-ABI library and the oracle's C tape VM are build products; build them once per session."""
    import subprocess

    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "optas_b200", "csrc")], check=True)
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)


def has_gpu() -> bool:
    from optas_b200 import _capi

    return _capi.device_count() > 0
