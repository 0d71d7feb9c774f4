#!/usr/bin/env python3
"""n = 512 (q = 12289), device-resident products on one stream: the three-layout kernel against the
one-layout-per-phase kernel (NTTB200_PLANT_N1024=0) and the unsigned kernel (NTTB200_PLANT_SIGNED=0)."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
m = importlib.import_module("ntt-based-polynomial-multiplier-fpga_b200")
n, q, rows, sets = 512, 12289, 1 << 19, 6
st = torch.cuda.current_stream().cuda_stream
bufs = [tuple(torch.randint(0, q, (rows, n), dtype=torch.int32, device="cuda") for _ in range(3)) for _ in range(sets)]
for tag, env in (("three-layout", {"NTTB200_PLANT_N1024": "2"}), ("one-layout", {"NTTB200_PLANT_N1024": "1"}),
                 ("unsigned", {"NTTB200_PLANT_SIGNED": "0"})):
    for k in ("NTTB200_PLANT_N1024", "NTTB200_PLANT_SIGNED"):
        os.environ.pop(k, None)
    os.environ.update(env)
    p = m.Plan(n, q)
    for i in range(6):
        a, b, c = bufs[i % sets]
        p.polymul_dev(c.data_ptr(), a.data_ptr(), b.data_ptr(), rows, st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 100
    e0.record()
    for i in range(steps):
        a, b, c = bufs[i % sets]
        p.polymul_dev(c.data_ptr(), a.data_ptr(), b.data_ptr(), rows, st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"n=512 {tag:13s} {rows * steps / ms / 1e3:8.1f} M polymul/s   hbm frac {rows * steps / (ms * 1e-3) * 12 * n / 6537.3e9:.3f}  {p.describe()}")
    p.close()
