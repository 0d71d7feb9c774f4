"""Small driver for compute-sanitizer (memcheck / racecheck / synccheck), run on the GPU box:
   compute-sanitizer --tool racecheck python tests/sanitize_small.py
One product per kernel family at a tiny batch; exits non-zero on a parity failure."""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
mod = importlib.import_module("ntt-based-polynomial-multiplier-fpga_b200")
from oracle import loader  # noqa: E402

O = loader.Oracle()
bad = 0
cases = [(256, 12289, False), (1024, 12289, False), (64, 257, False), (256, 12289, True),
         (256, 8380417, False), (256, 2013265921, False), (2048, 12289, False), (4096, 2013265921, False)]
for n, q, no_plant in cases:
    p = mod.Plan(n, q, no_plantard=no_plant)
    a, b = O.random((5, n), q, n + q), O.random((5, n), q, n + q + 1)
    a[0], b[0] = q - 1, q - 1
    ok = bool((p.polymul(a, b) == O.product(n, q, a, b, 10)).all())
    t = p.transform("inttmul_rev2std_scaled", p.transform("mulntt_std2rev", a))
    ok = ok and bool((t == a).all())
    print(p.describe(), "OK" if ok else "MISMATCH")
    bad += not ok
    p.close()
sys.exit(1 if bad else 0)
