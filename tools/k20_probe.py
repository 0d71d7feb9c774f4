import importlib, sys, os, torch, statistics
sys.path.insert(0, os.getcwd())
import bench
mod = importlib.import_module(bench.PKG); sh = importlib.import_module(bench.PKG + ".sharding")
torch.cuda.set_device(0); mod.set_device(0); dev = torch.device("cuda", 0)
W = bench.Workload(mod, sh, torch, "c2", dev, 0, 1)
for trial in range(6):
    ms, ms_local, _ = W.timed(20, 5, 0, clocks=False)
    print("trial", trial, "K=20", round(65536 * 20 / (ms * 1e-3) / 1e6, 1), "M/s")
# per-kernel timeline inside one K=20 region: events after each launch
for trial in range(2):
    for i in range(5): W.step(i)
    torch.cuda.synchronize()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(21)]
    evs[0].record()
    for i in range(20):
        W.step(i); evs[i + 1].record()
    torch.cuda.synchronize()
    print([round(evs[0].elapsed_time(evs[i + 1]) * 1e3) for i in range(20)])
W.close()
