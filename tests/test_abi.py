"""CPU: the C-ABI library loads and exports every function include/*.h declares; the
host-side (no-GPU) parts of the ABI -- table generator, root finder, parameter checks --
match the reference's tables and the oracle.  No compute call is made without a GPU."""
import ctypes
import glob
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N, Q, PSI = 256, 12289, 1002


def declared_functions():
    names = []
    for path in sorted(glob.glob(os.path.join(ROOT, "include", "*.h"))):
        text = open(path).read()
        text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
        text = re.sub(r"//[^\n]*", " ", text)
        text = "\n".join(l for l in text.splitlines() if not l.strip().startswith("#"))
        for m in re.finditer(r"\b([A-Za-z_]\w*)\s*\(", text):
            if m.group(1) not in ("sizeof", "defined"):
                names.append((os.path.basename(path), m.group(1)))
    return names


def test_headers_declare_something():
    names = [n for _, n in declared_functions()]
    for must in ("nttb200_plan_create", "nttb200_polymul_batch", "nttb200_polymul_batch_dev",
                 "nttb200_ntt_batch", "ntt256_product1", "ntt_red256_product4", "ntt_gs_std2rev"):
        assert must in names


def test_library_exports_every_declared_symbol(nttb200):
    lib = ctypes.CDLL(nttb200.lib_path())
    missing = [f"{h}:{n}" for h, n in declared_functions() if not hasattr(lib, n)]
    assert not missing, missing


def test_library_links_no_oracle_and_no_torch(nttb200):
    """The product must not route through the checker or through torch."""
    import subprocess
    out = subprocess.run(["ldd", nttb200.lib_path()], capture_output=True, text=True).stdout
    assert "oracle" not in out and "torch" not in out and "libcudart" in out


@pytest.mark.parametrize("kind", range(11))
def test_host_tables_match_reference(nttb200, golden, kind):
    assert (nttb200.make_table(kind, N, Q, PSI).astype(np.int64) == golden[f"table_{kind}"]).all()


TABLE_NAMES = ["psi_powers", "inv_psi_powers", "scaled_inv_psi_powers", "omega_powers", "omega_powers_rev",
               "inv_omega_powers", "inv_omega_powers_rev", "mixed_powers", "mixed_powers_rev",
               "inv_mixed_powers", "inv_mixed_powers_rev"]


def test_library_exports_the_reference_table_symbols(nttb200, golden):
    """ntt256_tables.h:29-44 / ntt_red256_tables.h:34-51: the 11 + 12 `extern const` arrays the
    reference's header-only n = 256 wrappers read, word for word the reference's literals."""
    lib = ctypes.CDLL(nttb200.lib_path())
    for k, nm in enumerate(TABLE_NAMES):
        t = (ctypes.c_uint16 * 256).in_dll(lib, "ntt256_" + nm)
        assert (np.ctypeslib.as_array(t).astype(np.int64) == golden[f"table_{k}"]).all(), nm
        r = (ctypes.c_int16 * 256).in_dll(lib, "ntt_red256_" + nm)
        assert (np.ctypeslib.as_array(r).astype(np.int64) == golden[f"red_table_{k}"]).all(), nm
    r = (ctypes.c_int16 * 256).in_dll(lib, "ntt_red256_scaled_inv_psi_powers_var")
    assert (np.ctypeslib.as_array(r).astype(np.int64) == golden["red_table_100"]).all()
    # every table name the legacy header declares is really exported
    text = open(os.path.join(ROOT, "include", "nttb200_legacy.h")).read()
    declared = re.findall(r"\b(ntt(?:_red)?256_\w+)\[256\]", text)
    assert len(set(declared)) == 23
    for nm in declared:
        (ctypes.c_int16 * 256).in_dll(lib, nm)


@pytest.mark.parametrize("kind", list(range(11)) + [100])
def test_red_table_generator_matches_reference(nttb200, golden, oracle, kind):
    """nttb200_make_red_table: value x 3^-1 centred; n^-1 3^-8 / 3^-6 for the two scaled tables."""
    L = nttb200.lib()
    out = np.zeros(256, np.int32)
    assert L.nttb200_make_red_table(kind, 256, PSI, out.ctypes.data) == 0
    assert (out.astype(np.int64) == golden[f"red_table_{kind}"]).all()
    for n in (8, 64, 1024):
        psi = oracle.psi(n, Q, 0)
        out = np.zeros(n, np.int32)
        assert L.nttb200_make_red_table(kind, n, psi, out.ctypes.data) == 0
        assert (out == oracle.red_table(kind, n, psi)).all()
    if kind == 2:
        assert golden["params"][6] == 8822 and out.dtype == np.int32


def test_reference_headers_link_against_the_library_alone(nttb200, tmp_path):
    """A program written against the reference's n = 256 headers (static inline wrappers over
    extern tables) links with -lnttb200 and nothing else (where the reference tree exists)."""
    import subprocess
    n256 = "/root/reference/Multiplier_NTT_Based/NTT_Software/NTT_Software_Evaluations/NTT-256"
    if not os.path.isdir(n256):
        pytest.skip("reference tree not present")
    exe = str(tmp_path / "wrappers")
    pkg = os.path.dirname(nttb200.lib_path())
    for opt in ("-O0", "-O2"):
        subprocess.run(["gcc", opt, "-I", n256, "-o", exe, os.path.join(ROOT, "tests", "ref_wrappers_client.c"),
                        "-L", pkg, "-lnttb200", f"-Wl,-rpath,{pkg}"], check=True)
    dyn = subprocess.run(["nm", "-D", exe], capture_output=True, text=True).stdout
    for sym in ("ntt256_omega_powers_rev", "ntt_red256_mixed_powers_rev", "ntt_ct_std2rev", "mulntt_red_ct_std2rev"):
        assert sym in dyn, sym          # resolved from libnttb200.so at load time


@pytest.mark.parametrize("n,q", [(64, 257), (256, 7681), (1024, 12289), (4096, 469762049), (65536, 2013265921)])
def test_host_tables_match_oracle(nttb200, oracle, n, q):
    psi = nttb200.find_psi(n, q)
    assert psi == oracle.psi(n, q, 0) and pow(psi, n, q) == q - 1
    for kind in range(12):
        assert (nttb200.make_table(kind, n, q, psi) == oracle.table(kind, n, q, psi)).all(), kind


def test_root_finder_and_primes(nttb200):
    assert nttb200.find_psi(256, 12289) == 3
    assert nttb200.find_psi(256, 7681) == 62          # HW/simulation/modelsim/test/PARAM.txt
    assert nttb200.find_psi(1024, 12289) == 7
    assert nttb200.find_psi(65536, 2013265921) == 37318
    assert nttb200.find_psi(256, 3329) == 0           # 512 does not divide 3328
    assert nttb200.find_omega(256, 3329) != 0
    for q, want in ((12289, True), (7681, True), (3329, True), (2013265921, True), (12287, False),
                    (1, False), (561, False), (2147483647, True), (4294967291, True), (4294967295, False)):
        assert nttb200.is_prime(q) is want, q


def test_bad_parameters_are_rejected_before_touching_the_gpu(nttb200):
    for n, q, psi in ((255, 12289, 0), (4, 12289, 0), (256, 12288, 0), (256, 3329, 0), (256, 12289, 5),
                      (1 << 18, 2013265921, 0)):
        with pytest.raises(nttb200.NttError) as e:
            nttb200.Plan(n, q, psi)
        assert "error -1" in str(e.value), (n, q, psi, str(e.value))


def test_no_cpu_fallback(nttb200):
    """Without a CUDA device the product path must fail loudly, never compute on the host."""
    if nttb200.device_count() > 0:
        pytest.skip("a GPU is visible; the no-fallback behaviour is exercised on CPU boxes")
    with pytest.raises(nttb200.NttError) as e:
        nttb200.Plan(N, Q, PSI)
    assert "error -2" in str(e.value) and "no CPU fallback" in str(e.value)
    with pytest.raises(nttb200.NttError):
        nttb200.ntt_table_batch(N, Q, "ct_std2rev", np.ones(N, np.uint32), np.zeros((1, N), np.int32))
