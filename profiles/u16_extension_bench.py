import importlib, sys, time
sys.path.insert(0, '.')
import numpy as np, torch
mod = importlib.import_module("ntt-based-polynomial-multiplier-fpga_b200")
n, q, batch = 256, 12289, 1 << 17
plan = mod.Plan(n, q, 1002)
# host path
nb = batch * n * 2
ha, hb, hc = [mod.host_alloc((batch, n // 2)) for _ in range(3)]       # int32 pinned buffers viewed as u16
rng = np.random.default_rng(1)
ha.array.view(np.uint16)[:] = rng.integers(0, q, (batch, n)); hb.array.view(np.uint16)[:] = rng.integers(0, q, (batch, n))
L = mod.lib()
for _ in range(3): L.nttb200_polymul_batch_u16(plan._h, hc.ptr, ha.ptr, hb.ptr, batch)
t0 = time.perf_counter(); K = 20
for _ in range(K): L.nttb200_polymul_batch_u16(plan._h, hc.ptr, ha.ptr, hb.ptr, batch)
dt = time.perf_counter() - t0
print(f"u16 e2e (host buffers): {batch*K/dt/1e6:.2f} M polymul/s  H2D {2*nb*K/dt/1e9:.1f} GB/s")
# device path
a = torch.randint(0, q, (4, batch, n), dtype=torch.int16, device='cuda'); b = torch.randint(0, q, (4, batch, n), dtype=torch.int16, device='cuda'); c = torch.empty_like(a)
st = torch.cuda.current_stream().cuda_stream
for i in range(8): plan.polymul_u16_dev(c[i%4].data_ptr(), a[i%4].data_ptr(), b[i%4].data_ptr(), batch, st)
torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(100): plan.polymul_u16_dev(c[i%4].data_ptr(), a[i%4].data_ptr(), b[i%4].data_ptr(), batch, st)
e1.record(); torch.cuda.synchronize()
print(f"u16 device-resident: {batch*100/(e0.elapsed_time(e1)*1e-3)/1e6:.1f} M polymul/s")
