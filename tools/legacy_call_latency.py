"""Per-call latency of the one-polynomial-per-call legacy surface (GPU box)."""
import importlib, sys, time
sys.path.insert(0, '.')
import numpy as np
mod = importlib.import_module("ntt-based-polynomial-multiplier-fpga_b200")
g = np.load("tests/golden/ref_256_12289.npz")
lg = mod.legacy
a, b = g["fixture_a"].astype(np.int32), g["fixture_b"].astype(np.int32)
def t(f, k=300):
    f(); f()
    t0 = time.perf_counter()
    for _ in range(k): f()
    return (time.perf_counter() - t0) / k * 1e6
print(f"ntt256_product1          {t(lambda: lg.product('ntt256_product1', a, b)):8.1f} us/call")
print(f"ntt_ct_std2rev (n=256)   {t(lambda: lg.transform('ntt_ct_std2rev', a, g['table_4'])):8.1f} us/call")
print(f"ntt_red_ct_std2rev       {t(lambda: lg.red_transform('ntt_red_ct_std2rev', a, g['red_table_4'])):8.1f} us/call")
print(f"reduce_array             {t(lambda: lg.red_helper('reduce_array', a)):8.1f} us/call")
print(f"mul_array                {t(lambda: lg.mul_array(a, b)):8.1f} us/call")
