"""Device-resident throughput of the Shoup-arithmetic product kernels (n <= 1024, moduli that do
not fit the Plantard kernel) -- the classes bench.py's workloads do not cover.
Usage: python profiles/shoup_small_bench.py   (GPU box; NTTB200_LIB selects another build)"""
import importlib
import sys

sys.path.insert(0, ".")
import torch

mod = importlib.import_module("ntt-based-polynomial-multiplier-fpga_b200")
CASES = [(256, 12289, True, "LAZY (Plantard off)"), (256, 8380417, False, "HARVEY 23-bit"),
         (256, 998244353, False, "HARVEY 30-bit"), (256, 2013265921, False, "CANON 31-bit"),
         (1024, 8380417, False, "HARVEY 23-bit"), (1024, 2013265921, False, "CANON 31-bit")]
for n, q, noplant, name in CASES:
    batch = (1 << 26) // n
    plan = mod.Plan(n, q, 0, no_plantard=noplant)
    sets = 3
    a = torch.randint(0, q, (sets, batch, n), dtype=torch.int32, device="cuda")
    b = torch.randint(0, q, (sets, batch, n), dtype=torch.int32, device="cuda")
    c = torch.empty_like(a)
    st = torch.cuda.current_stream().cuda_stream
    for i in range(6):
        plan.polymul_dev(c[i % sets].data_ptr(), a[i % sets].data_ptr(), b[i % sets].data_ptr(), batch, st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 60
    e0.record()
    for i in range(K):
        plan.polymul_dev(c[i % sets].data_ptr(), a[i % sets].data_ptr(), b[i % sets].data_ptr(), batch, st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"n={n} q={q} {name}: {batch * K / (ms * 1e-3) / 1e6:.1f} M polymul/s  "
          f"({12 * n * batch * K / (ms * 1e-3) / 1e9:.0f} GB/s algorithmic)  [{plan.describe()}]")
    plan.close()
    del a, b, c
