#!/usr/bin/env python3
"""Butterfly-level issue rates on the B200 integer pipes (nttb200_measure_int_peak): how many
lane-butterflies per second each candidate instruction mix sustains, 8 independent chains per
thread.  python profiles/microbench_butterflies.py > profiles/<round>_microbench_butterflies.txt"""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
m = importlib.import_module("ntt-based-polynomial-multiplier-fpga_b200")
NAMES = {0: "IMAD", 1: "IMAD.HI", 2: "IADD", 12: "SHF + IADD pairs (x2)", 15: "LEA.HI.SX32",
         3: "Shoup LAZY butterfly (IMAD.HI, 2 IMAD, 2 IADD)", 13: "unsigned Plantard, full word (IMAD, IMAD.HI, 2 IADD3)",
         11: "Plantard half word, 6 instructions (IMAD, SHF, IMAD, SHF, 2 IADD)",
         14: "signed Plantard half word, 5 instructions (IMAD, SHF, IMAD, LEA.HI.SX32, IADD3)",
         16: "signed Plantard full word, 4 instructions (IMAD, IMAD.HI, LEA.HI.SX32, IADD3)",
         17: "14 and 16 alternating",
         18: "14 with the second leg as mad.lo (X, 2, -X'): ptxas levels IMAD / LEA over the two pipes",
         19: "14 and 18 alternating",
         23: "14, 14, 18 in turn",
         20: "signed Plantard half word, Gentleman-Sande, 6 instructions (2 IADD, IMAD, SHF, IMAD, SHF)",
         21: "20 with the sum as mad.lo (X, 1, Y)",
         24: "31-bit class (q = 2013265921), Cooley-Tukey, canonical results, 8 instructions",
         25: "31-bit class, Cooley-Tukey, results left in [0, 2q) for the next multiplication, 6 instructions",
         26: "31-bit class, Gentleman-Sande, 7 instructions",
         22: "signed Gentleman-Sande, last shift of the product left pending (LEA.HI.SX32, IADD3, IMAD, SHF, IMAD)"}
sms, clk = 148, 1.965e9
for k, name in NAMES.items():
    r = m.measure_int_peak(k)
    print(f"{k:3d} {name:90s} {r / 1e12:7.3f} T lane-ops/s  = {r / sms / clk:6.2f} lanes/clk/SM")
