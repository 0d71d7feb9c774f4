#!/usr/bin/env python3
"""Golden data for the "next" rows: parameter generator and text formats.

Run from the repo root in the dev container:  python tests/golden/make_golden_next.py
Uses oracle/_ref/libgen_ref.so and oracle/_ref/time_testing256_ref, both compiled by
oracle/Makefile from the UNMODIFIED reference sources, and the reference's committed
hardware vectors.  Writes
  tests/golden/gen_params.npz              generate_params() / generate_twiddles() outputs,
                                           W.txt / WINV.txt / PARAM.txt of the q=7681 testbench
  tests/golden/time_testing256_stdout.txt  stdout of the reference benchmark on its own
                                           coeficientes_{a,b}.txt (the time line removed)
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference/Multiplier_NTT_Based"
N256 = f"{REF}/NTT_Software/NTT_Software_Evaluations/NTT-256"
HWT = f"{REF}/Hardware_Multiplier/simulation/modelsim/test"
OUT = os.path.join(ROOT, "tests", "golden")


def read_hex(path):
    vals = []
    with open(path) as f:
        for line in f:
            line = line.split("//")[0].strip()
            if line:
                vals.append(int(line, 16))
    return np.array(vals, dtype=np.int64)


def main():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"], check=True)
    G = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libgen_ref.so"))
    vals = [C.c_int(0) for _ in range(8)]
    G.generate_params(*[C.byref(v) for v in vals])
    psi, psi_inv, w, w_inv, R, n_inv, PE, q = [v.value for v in vals]
    W = np.zeros(272, np.uint32)
    WI = np.zeros(272, np.uint32)
    G.generate_twiddles(W.ctypes.data_as(C.c_void_p), WI.ctypes.data_as(C.c_void_p),
                        C.c_uint32(w), C.c_uint32(w_inv), C.c_uint32(q), C.c_uint32(R))
    g = {"ref_params": np.array([psi, psi_inv, w, w_inv, R, n_inv, PE, q], dtype=np.int64),
         "ref_W": W, "ref_W_INV": WI}
    # the same generator at the hardware testbench's parameters (q=7681: PARAM.txt)
    hw = read_hex(f"{HWT}/PARAM.txt")
    g["hw_PARAM"] = hw
    g["hw_W"] = read_hex(f"{HWT}/W.txt")
    g["hw_WINV"] = read_hex(f"{HWT}/WINV.txt")
    W2 = np.zeros(272, np.uint32)
    WI2 = np.zeros(272, np.uint32)
    hq, hw_w, hw_winv, hR = int(hw[1]), int(hw[2]), int(hw[3]), 1 << 18
    G.generate_twiddles(W2.ctypes.data_as(C.c_void_p), WI2.ctypes.data_as(C.c_void_p),
                        C.c_uint32(hw_w), C.c_uint32(hw_winv), C.c_uint32(hq), C.c_uint32(hR))
    g["ref_W_q7681"], g["ref_W_INV_q7681"] = W2, WI2
    print("reference generate_twiddles(q=7681) == W.txt:", bool((W2 == g["hw_W"]).all()),
          " == WINV.txt:", bool((WI2 == g["hw_WINV"]).all()))
    g["modexp_cases"] = np.array([[b, e, m, pow(b, e, m)] for b, e, m in
                                  [(3, 256, 12289), (1002, 255, 12289), (62, 511, 7681), (7, 12288, 12289),
                                   (2, 0, 97), (123456, 65537, 2013265921)]], dtype=np.int64)
    np.savez_compressed(os.path.join(OUT, "gen_params.npz"), **g)

    out = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "time_testing256_ref")], cwd=N256,
                         capture_output=True, text=True, check=True).stdout
    keep = [ln for ln in out.splitlines() if not ln.startswith("Tempo total")]
    with open(os.path.join(OUT, "time_testing256_stdout.txt"), "w") as f:
        f.write("\n".join(keep) + "\n")
    # the reference's KAT program NTT/test_prod_ntt256.c, compiled from its own sources
    exe = os.path.join(ROOT, "oracle", "_ref", "test_prod_ntt256_ref")
    subprocess.run(["gcc", "-O2", "-I", N256, "-o", exe, "-x", "c", f"{N256}/NTT/test_prod_ntt256.c",
                    f"{N256}/NTT/ntt.C", f"{N256}/NTT/ntt256.C", f"{N256}/NTT/ntt256_tables.C"], check=True)
    kat = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
    with open(os.path.join(OUT, "test_prod_ntt256_stdout.txt"), "w") as f:
        f.write(kat)
    print("params:", g["ref_params"], "lines of stdout kept:", len(keep))


if __name__ == "__main__":
    sys.exit(main())
