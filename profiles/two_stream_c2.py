"""c2 (batch 2^16 per call) issued on ONE stream vs alternately on TWO streams: how much of the
per-launch start-up / drain (the difference between c2 and c3 in DESIGN.md section 4) a caller
recovers by keeping two independent calls in flight.  Not a bench.py number."""
import importlib
import sys

sys.path.insert(0, ".")
import torch

mod = importlib.import_module("ntt-based-polynomial-multiplier-fpga_b200")
logb = int(sys.argv[1]) if len(sys.argv) > 1 else 16
n, q, batch = 256, 12289, 1 << logb
sets, K = (6, 600) if logb <= 17 else (3, 60)
plan = mod.Plan(n, q, 1002)
a = torch.randint(0, q, (sets, batch, n), dtype=torch.int32, device="cuda")
b = torch.randint(0, q, (sets, batch, n), dtype=torch.int32, device="cuda")
c = torch.empty_like(a)
streams = [torch.cuda.Stream() for _ in range(3)]
for ns in (1, 2, 3):
    for rep in range(2):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in streams[:ns]:
            s.wait_event(e0)
        for i in range(K):
            st = streams[i % ns]
            plan.polymul_dev(c[i % sets].data_ptr(), a[i % sets].data_ptr(), b[i % sets].data_ptr(), batch, st.cuda_stream)
        for s in streams[:ns]:
            ev = torch.cuda.Event()
            ev.record(s)
            torch.cuda.current_stream().wait_event(ev)
        e1.record()
        torch.cuda.synchronize()
    print(f"batch 2^{logb}, {ns} stream(s): {batch * K / (e0.elapsed_time(e1) * 1e-3) / 1e6:.1f} M polymul/s")
