"""pytest configuration: `-m "not gpu"` runs on the CPU-only container, `-m gpu` on a B200."""
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with gpurun on a B200)")


@pytest.fixture(scope="session", autouse=True)
def _built_libraries():
    """Make sure both shared libraries exist (they are build artefacts, not in git)."""
    from custos_b200 import build as cb_build
    from oracle import oracle as orc
    cb_build.build()
    orc.build()


@pytest.fixture(scope="session")
def raw_device():
    from custos_b200.raw import RawDevice
    dev = RawDevice(0)
    yield dev
    dev.close()
