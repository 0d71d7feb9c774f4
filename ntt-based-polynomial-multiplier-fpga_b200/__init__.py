"""nttb200 -- B200-native batched NTT polynomial multiplication.

The product is ``libnttb200.so`` (hand-written CUDA for sm_100a behind the C ABI of
``include/nttb200.h``); this Python package is only the ctypes binding used by the test
suite and ``bench.py``.  The package directory name contains hyphens, so import it with
``importlib.import_module("ntt-based-polynomial-multiplier-fpga_b200")`` (done by
``__graft_entry__`` and ``tests/conftest.py``, which alias it as ``nttb200``).
"""
from .capi import (  # noqa: F401
    NttError, Plan, lib, lib_path, build_library, device_count, set_device, make_table,
    find_psi, find_omega, is_prime, host_alloc, measure_int_peak, last_launch_count,
    TRANSFORMS, DATAFLOWS, TABLES, ntt_table_batch, legacy, MultiPlan, shard_bounds,
)
from . import inputs  # noqa: F401,E402
