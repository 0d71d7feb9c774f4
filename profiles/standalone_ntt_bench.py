import importlib, sys, os
sys.path.insert(0, '.')
import torch
mod = importlib.import_module("ntt-based-polynomial-multiplier-fpga_b200")
n, q, batch = 256, 12289, 1 << 18
p = mod.Plan(n, q, 1002)
a = torch.randint(0, q, (batch, n), dtype=torch.int32, device='cuda')
st = torch.cuda.current_stream().cuda_stream
for kind in ("mulntt_std2rev", "inttmul_rev2std_scaled"):
    for _ in range(5): p.transform_dev(kind, a.data_ptr(), batch, st)
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): p.transform_dev(kind, a.data_ptr(), batch, st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    print(os.environ.get("NTTB200_NTT_SHOUP", "plantard"), kind, f"{batch/ms/1e3:.1f} M transforms/s  {8*n*batch/ms/1e6:.0f} GB/s")
