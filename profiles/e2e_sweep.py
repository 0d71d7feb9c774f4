import importlib, sys, time, os
sys.path.insert(0, '.')
import numpy as np
mod = importlib.import_module("ntt-based-polynomial-multiplier-fpga_b200")
n, q, batch = 256, 12289, 1 << 16
plan = mod.Plan(n, q, 1002)
ha, hb, hc = mod.host_alloc((batch, n)), mod.host_alloc((batch, n)), mod.host_alloc((batch, n))
rng = np.random.default_rng(1)
ha.array[:] = rng.integers(0, q, (batch, n)); hb.array[:] = rng.integers(0, q, (batch, n))
for _ in range(3): plan.polymul_host_ptr(hc.ptr, ha.ptr, hb.ptr, batch)
t0 = time.perf_counter(); K = 20
for _ in range(K): plan.polymul_host_ptr(hc.ptr, ha.ptr, hb.ptr, batch)
dt = time.perf_counter() - t0
print(os.environ.get("NTTB200_NSLOT"), os.environ.get("NTTB200_SLOT_MB"), f"{batch*K/dt/1e6:.2f} M polymul/s  H2D {2*batch*n*4*K/dt/1e9:.1f} GB/s")
