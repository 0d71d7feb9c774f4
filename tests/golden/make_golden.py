#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the UNMODIFIED reference, in the dev container.

Run from the repo root:  python tests/golden/make_golden.py
Needs /root/reference (read-only) and gcc; builds oracle/_ref via oracle/Makefile and
records outputs of the compiled reference C (there is no committed expected output for the
C path in the reference -- SURVEY.md section 4), plus the reference's own committed vectors:

* N256/coeficientes_{a,b}.txt            (inputs of time_testing256.c:139)
* HW/simulation/modelsim/test/*.txt       (q=7681 golden vectors of the Verilog testbench)
* the known-answer products quoted in the reference's drivers / testbenches
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import loader  # noqa: E402

REF = "/root/reference/Multiplier_NTT_Based"
N256 = f"{REF}/NTT_Software/NTT_Software_Evaluations/NTT-256"
HWT = f"{REF}/Hardware_Multiplier/simulation/modelsim/test"
OUT = os.path.join(ROOT, "tests", "golden")


def read_ints(path):
    with open(path) as f:
        return np.array([int(t) for t in f.read().split()], dtype=np.int32)


def read_hex(path):
    with open(path) as f:
        vals = []
        for line in f:
            line = line.split("//")[0].strip()      # PARAM.txt carries trailing // comments
            if line:
                vals.append(int(line, 16))
        return np.array(vals, dtype=np.int64)


def main():
    loader.build(ref=True)
    R = loader.Reference()
    O = loader.Oracle()
    n, q = 256, 12289
    g = {}

    # (5) the reference benchmark's fixture pair and its product under all variants
    fa, fb = read_ints(f"{N256}/coeficientes_a.txt"), read_ints(f"{N256}/coeficientes_b.txt")
    assert fa.size == 256 and fb.size == 256
    g["fixture_a"], g["fixture_b"] = fa, fb
    outs = {v: R.product(fa, fb, v)[0] for v in (1, 4, 101, 104, 10, 110)}
    for v, c in outs.items():
        assert (c == outs[1]).all(), v
    g["fixture_c"] = outs[1]

    # (1)-(4) known-answer inputs quoted in the reference's drivers
    def poly(d):
        p = np.zeros(n, dtype=np.int32)
        for k, v in d.items():
            p[k] = v
        return p
    kat_a = np.stack([poly({0: 1, 1: 2}), poly({0: 1, 1: 2, 2: 3}), poly({0: 1, 1: 2, 2: 3}),
                      poly({0: 1, 1: 2, 4: 2})])
    kat_b = np.stack([poly({0: 3}), poly({0: 2}), poly({0: 2, 1: 2}), poly({0: 3, 1: 3, 3: 1})])
    kat_c = R.product(kat_a, kat_b, 1)
    # stated expectations: test_prod_nttred256.c:48-57 (3+6x); NTT_PCIECommunicationv2.c:157-158
    # ({2,4,6}); NTT_PolyMul_test.v:167-192 ({2,6,10,6})
    assert kat_c[0][:3].tolist() == [3, 6, 0]
    assert kat_c[1][:4].tolist() == [2, 4, 6, 0]
    assert kat_c[2][:5].tolist() == [2, 6, 10, 6, 0]
    assert kat_c[3][:8].tolist() == [3, 9, 6, 1, 8, 6, 0, 2]
    g["kat_a"], g["kat_b"], g["kat_c"] = kat_a, kat_b, kat_c

    # random rows + edge rows, every product variant and every standalone transform
    rows = 96
    ra = O.random((rows, n), q, 0x4E54544232303001)
    rb = O.random((rows, n), q, 0x4E54544232303002)
    ra[0], rb[0] = 0, 0
    ra[1], rb[1] = q - 1, q - 1
    ra[2], rb[2] = 0, q - 1
    ra[2][0] = 1                       # delta_0 * (q-1)
    ra[3], rb[3] = 0, 0
    ra[3][n - 1], rb[3][n - 1] = 1, 1  # x^(n-1) * x^(n-1) = -x^(n-2)
    g["rand_a"], g["rand_b"] = ra, rb
    rc = R.product(ra, rb, 1)
    for v in (4, 101, 104, 10, 110):
        assert (R.product(ra, rb, v) == rc).all(), v
    g["rand_c"] = rc
    # a/b post-state of ntt256_product1 ("arrays a and b are modified", ntt256.h:80)
    a0, b0, c0 = ra[4].copy(), rb[4].copy(), np.zeros(n, dtype=np.int32)
    R.lib.ref_product(1, c0, a0, b0)
    g["clobber_a_after_product1"], g["clobber_b_after_product1"] = a0, b0
    tin = ra[:16]
    for tid in list(range(13)):
        g[f"transform_{tid}"] = R.transform(tid, tin)
    cin = tin.copy()
    cin[cin > 6144] -= q               # RED transforms take centred inputs
    g["red_transform_in"] = cin
    for tid in range(100, 112):
        g[f"transform_{tid}"] = R.transform(tid, cin)
    # reference tables (both sets) so the table generator is pinned without /root/reference
    for k in range(11):
        g[f"table_{k}"] = R.table(k)
        g[f"red_table_{k}"] = R.table(k, red=True)
    g["red_table_100"] = R.table(100, red=True)
    g["params"] = np.array([R.param(i) for i in range(8)], dtype=np.int64)
    # the generic entry points run with an ARBITRARY caller table (entries in [0, q), p[t] != 1):
    # pins the j = 0 peel of the un-merged functions (ntt.C:313-317 ...), n = 256 and n = 64
    rng = np.random.default_rng(20261019)
    for nn in (256, 64):
        tab = rng.integers(2, q, size=nn, dtype=np.int64).astype(np.uint16)
        g[f"arb_table_{nn}"] = tab
        ain = ra[:8, :nn].copy()
        g[f"arb_in_{nn}"] = ain
        for tid in range(9):
            g[f"arb_transform_{nn}_{tid}"] = R.transform_tab(tid, ain, tab)
        rtab = (tab.astype(np.int64) % 12289)
        rtab[rtab > 6144] -= 12289                      # |p| <= 6144 (ntt_red.h:169-173)
        rtab = rtab.astype(np.int16)
        g[f"arb_red_table_{nn}"] = rtab
        rin = ain.copy()
        rin[rin > 6144] -= q
        g[f"arb_red_in_{nn}"] = rin
        for tid in range(1, 9):
            g[f"arb_red_transform_{nn}_{tid}"] = R.transform_tab(tid, rin, rtab, red=True)
    # a/b post-state of the optimized products ("a and b are modified", ntt_red256.h:77-86)
    for v in (101, 104):
        c1, a1, b1 = R.product_post_state(ra[4], rb[4], v)
        assert (c1 == rc[4]).all()
        g[f"clobber_a_after_product{v}"], g[f"clobber_b_after_product{v}"] = a1, b1
    np.savez_compressed(os.path.join(OUT, "ref_256_12289.npz"), **g)

    # (6) hardware golden vectors, q = 7681 (PARAM.txt: N, q, w, w_inv, psi, psi_inv, n_inv*R, R)
    h = {nm: read_hex(f"{HWT}/{nm}.txt") for nm in
         ("PARAM", "NTT_DIN", "NTT_DOUT", "INTT_DIN", "INTT_DOUT", "POLY_A_HEX", "POLY_B_HEX", "W", "WINV")}
    assert h["PARAM"][0] == 256 and h["PARAM"][1] == 7681
    np.savez_compressed(os.path.join(OUT, "hw_256_7681.npz"), **h)
    print("wrote", os.listdir(OUT))


if __name__ == "__main__":
    main()
