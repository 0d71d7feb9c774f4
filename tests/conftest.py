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


def oracle_product_mt(oracle, n, q, a, b, variant=10, psi=0, threads=None):
    """oracle.product over the rows of (a, b), split over the host cores (the ctypes call releases
    the GIL and the oracle's plans are read-only)."""
    from concurrent.futures import ThreadPoolExecutor
    a = np.ascontiguousarray(a, dtype=np.int32).reshape(-1, n)
    b = np.ascontiguousarray(b, dtype=np.int32).reshape(-1, n)
    rows = a.shape[0]
    threads = threads or min(32, len(os.sched_getaffinity(0)))
    oracle.plan(n, q, psi)                       # create the plan once, outside the pool
    if rows < 2 * threads:
        return oracle.product(n, q, a, b, variant, psi)
    cuts = [rows * i // threads for i in range(threads + 1)]
    with ThreadPoolExecutor(threads) as ex:
        parts = list(ex.map(lambda i: oracle.product(n, q, a[cuts[i]:cuts[i + 1]], b[cuts[i]:cuts[i + 1]], variant, psi),
                            range(threads)))
    return np.concatenate(parts)


@pytest.fixture(scope="session")
def oracle_mt(oracle):
    return lambda n, q, a, b, variant=10, psi=0: oracle_product_mt(oracle, n, q, a, b, variant, psi)
