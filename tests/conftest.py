import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PKG = "ntt-based-polynomial-multiplier-fpga_b200"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _pkg():
    mod = importlib.import_module(PKG)
    sys.modules.setdefault("nttb200", mod)
    return mod


@pytest.fixture(scope="session")
def nttb200():
    mod = _pkg()
    if not os.path.exists(mod.lib_path()):
        mod.build_library()
    return mod


@pytest.fixture(scope="session")
def oracle():
    from oracle import loader
    return loader.Oracle()


@pytest.fixture(scope="session")
def loader():
    from oracle import loader as ld
    return ld


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_256_12289.npz"))


@pytest.fixture(scope="session")
def hw_golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "hw_256_7681.npz"))


@pytest.fixture(scope="session")
def gpu(nttb200):
    if nttb200.device_count() < 1:
        pytest.fail("no CUDA device visible to libnttb200.so (gpu-marked test on a CPU box?)")
    return nttb200
