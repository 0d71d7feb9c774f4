"""Generator_Params parity, text formats, the time_testing256-compatible CLI and
the Terasic-ABI plugin.  CPU tests pin the host code against golden data generated from the
compiled reference (tests/golden/make_golden_next.py); gpu tests run the CLI and the plugin
protocol on the GPU box."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
PKG_DIR = os.path.join(ROOT, "ntt-based-polynomial-multiplier-fpga_b200")


@pytest.fixture(scope="module")
def L(nttb200):
    lib = C.CDLL(nttb200.lib_path())
    lib.nttb200_modexp.restype = C.c_uint32
    lib.nttb200_modexp.argtypes = [C.c_uint32] * 3
    lib.nttb200_gen_prime.restype = C.c_uint32
    lib.nttb200_gen_prime.argtypes = [C.c_uint32, C.c_uint32, C.c_uint64]
    lib.nttb200_gen_twiddle_count.restype = C.c_size_t
    lib.nttb200_gen_twiddle_count.argtypes = [C.c_uint32, C.c_uint32]
    lib.nttb200_gen_twiddles.argtypes = [C.c_void_p, C.c_void_p] + [C.c_uint32] * 6
    lib.nttb200_gen_params.argtypes = [C.c_uint32] * 4 + [C.c_void_p]
    lib.nttb200_modinv.argtypes = [C.c_int32, C.c_int32]
    lib.nttb200_write_coeff_file.argtypes = [C.c_char_p, C.c_void_p, C.c_size_t]
    lib.nttb200_read_coeff_file.argtypes = [C.c_char_p, C.c_void_p, C.c_size_t]
    lib.nttb200_read_coeff_file.restype = C.c_long
    lib.nttb200_write_hex_file.argtypes = [C.c_char_p, C.c_void_p, C.c_size_t]
    lib.nttb200_read_hex_file.argtypes = [C.c_char_p, C.c_void_p, C.c_size_t]
    lib.nttb200_read_hex_file.restype = C.c_long
    return lib


@pytest.fixture(scope="module")
def gen_golden():
    return np.load(os.path.join(GOLD, "gen_params.npz"))


def _params(L, n, K, P, q):
    out = (C.c_uint32 * 9)()
    rc = L.nttb200_gen_params(n, K, P, q, out)
    return rc, list(out)


def test_generate_params_matches_reference(L, gen_golden):
    """G/generate_params.C:12-52 at its own parameters (N=256, K=13, P=8, q forced to 12289)."""
    psi, psi_inv, w, w_inv, R, n_inv, PE, q = [int(x) for x in gen_golden["ref_params"]]
    rc, g = _params(L, 256, 13, 8, 0)
    assert rc == 0
    assert g == [256, q, psi, psi_inv, w, w_inv, R, n_inv, PE]
    vals = [C.c_int(0) for _ in range(8)]
    L.generate_params(*[C.byref(v) for v in vals])           # the reference's own signature
    assert [v.value for v in vals] == [psi, psi_inv, w, w_inv, R, n_inv, PE, q]


def test_generate_twiddles_matches_reference_and_hw_vectors(L, gen_golden):
    psi, psi_inv, w, w_inv, R, n_inv, PE, q = [int(x) for x in gen_golden["ref_params"]]
    assert L.nttb200_gen_twiddle_count(256, 8) == 272
    W, WI = np.zeros(272, np.uint32), np.zeros(272, np.uint32)
    assert L.nttb200_gen_twiddles(W.ctypes.data, WI.ctypes.data, 256, 8, w, w_inv, q, R) == 0
    assert (W == gen_golden["ref_W"]).all() and (WI == gen_golden["ref_W_INV"]).all()
    W2, WI2 = np.zeros(272, np.uint32), np.zeros(272, np.uint32)
    L.generate_twiddles(W2.ctypes.data, WI2.ctypes.data, C.c_uint32(w), C.c_uint32(w_inv), C.c_uint32(q), C.c_uint32(R))
    assert (W2 == W).all() and (WI2 == WI).all()
    # the Verilog testbench's committed stream (q=7681, w=0xf04, R=2^18): W.txt / WINV.txt
    hw = gen_golden["hw_PARAM"]
    hq, hw_w, hw_winv = int(hw[1]), int(hw[2]), int(hw[3])
    assert L.nttb200_gen_twiddles(W.ctypes.data, WI.ctypes.data, 256, 8, hw_w, hw_winv, hq, 1 << 18) == 0
    assert (W == gen_golden["hw_W"]).all() and (WI == gen_golden["hw_WINV"]).all()
    # and the parameters of that testbench are what the generator derives for q = 7681
    rc, g = _params(L, 256, 13, 8, 7681)
    assert rc == 0 and g[2] == int(hw[4]) and g[4] == hw_w and g[5] == hw_winv and g[6] == 1 << 18


def test_generator_number_theory(L, gen_golden):
    for b, e, m, want in gen_golden["modexp_cases"]:
        assert L.nttb200_modexp(int(b), int(e), int(m)) == int(want)
    for a, m in ((3, 12289), (256, 12289), (1002, 12289), (62, 7681), (12288, 12289)):
        assert L.nttb200_modinv(a, m) == pow(a, -1, m)
    assert L.nttb200_modinv(6, 9) == -1
    primes = [p for p in range(2, 2000) if all(p % d for d in range(2, int(p ** 0.5) + 1))]
    assert [p for p in range(2, 2000) if L.nttb200_miller_rabin(p)] == primes
    for p in (12289, 7681, 3329, 8380417, 2013265921, 4294967291):
        assert L.nttb200_miller_rabin(p)
    for c in (12289 * 7681, 3215031751, 4294967295, 2013265921 + 2):
        assert not L.nttb200_miller_rabin(c)
    for k, n in ((13, 256), (14, 256), (14, 1024), (23, 256), (31, 65536)):
        for seed in range(5):
            p = L.nttb200_gen_prime(k, n, seed)
            assert p and 1 << (k - 1) <= p < 1 << k and p % (2 * n) == 1 and L.nttb200_miller_rabin(p)
    assert _params(L, 256, 13, 8, 3329)[0] != 0            # no 512-th root of unity
    assert _params(L, 256, 13, 8, 12291)[0] != 0           # not prime


def test_generate_coeff_command_line_twin(L, tmp_path):
    """tools/nttb200_generate_coeff.c: the reference's generate_coeff.c:12-59 with N, Q and the seed as
    optional arguments -- same messages, same two files, 10 decimals per line, every value in [0, Q); the
    files are what the library's reader (and time_testing256.c:17-44) reads back."""
    import subprocess
    exe = os.path.join(PKG_DIR, "nttb200_generate_coeff")
    assert os.path.exists(exe), "build() makes it (csrc/Makefile)"
    for args, n, q in (([], 256, 12289), (["1024", "7681", "5"], 1024, 7681)):
        r = subprocess.run([exe] + args, cwd=tmp_path, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert r.stdout.splitlines() == [f"Gerando coeficientes aleatorios com Q={q} e N={n}.",
                                         "Arquivos 'coeficientes_a.txt' e 'coeficientes_b.txt' gerados com sucesso!"]
        got = []
        for name in ("coeficientes_a.txt", "coeficientes_b.txt"):
            text = open(tmp_path / name).read()
            lines = text.split("\n")
            assert all(len(ln.split()) == 10 and ln.endswith(" ") for ln in lines[:n // 10])
            back = np.full(n + 1, -1, np.int32)
            assert L.nttb200_read_coeff_file(str(tmp_path / name).encode(), back.ctypes.data, n + 1) == n
            assert back[:n].min() >= 0 and back[:n].max() < q
            got.append(back[:n].copy())
        assert not (got[0] == got[1]).all()
    again = subprocess.run([exe, "1024", "7681", "5"], cwd=tmp_path, capture_output=True, text=True)
    assert again.returncode == 0
    back = np.zeros(1024, np.int32)
    L.nttb200_read_coeff_file(str(tmp_path / "coeficientes_b.txt").encode(), back.ctypes.data, 1024)
    assert (back == got[1]).all()                    # a given seed reproduces the files
    assert subprocess.run([exe, "0"], cwd=tmp_path, capture_output=True).returncode == 2


def test_text_formats(L, golden, tmp_path):
    a = golden["fixture_a"].astype(np.int32)
    p = str(tmp_path / "coef.txt").encode()
    assert L.nttb200_write_coeff_file(p, a.ctypes.data, a.size) == 0
    text = open(p).read()
    lines = text.split("\n")
    assert all(len(ln.split()) == 10 and ln.endswith(" ") for ln in lines[:25]) and lines[25].split() == [str(x) for x in a[250:]]
    ref_file = "/root/reference/Multiplier_NTT_Based/NTT_Software/NTT_Software_Evaluations/NTT-256/coeficientes_a.txt"
    if os.path.exists(ref_file):                             # byte-identical to the reference's own file
        assert text == open(ref_file).read()
    back = np.zeros(256, np.int32)
    assert L.nttb200_read_coeff_file(p, back.ctypes.data, 256) == 256 and (back == a).all()
    assert L.nttb200_read_coeff_file(b"/nonexistent/x", back.ctypes.data, 256) == -1
    h = str(tmp_path / "v.txt").encode()
    u = golden["fixture_b"].astype(np.uint32)
    assert L.nttb200_write_hex_file(h, u.ctypes.data, u.size) == 0
    assert open(h).read().split("\n")[:3] == [format(int(x), "x") for x in u[:3]]
    back_u = np.zeros(256, np.uint32)
    assert L.nttb200_read_hex_file(h, back_u.ctypes.data, 256) == 256 and (back_u == u).all()


def test_plugin_exports_the_twelve_terasic_symbols(nttb200):
    """C/PCIE.c:78-103 dlsym()s exactly these and fails if any is missing."""
    path = os.path.join(PKG_DIR, "terasic_pcie_qsys.so")
    assert os.path.exists(path), "run __graft_entry__.build()"
    lib = C.CDLL(path)
    for s in ("PCIE_Open", "PCIE_Close", "PCIE_Read32", "PCIE_Write32", "PCIE_Read16", "PCIE_Write16",
              "PCIE_Read8", "PCIE_Write8", "PCIE_DmaWrite", "PCIE_DmaRead", "PCIE_DmaFifoWrite", "PCIE_DmaFifoRead"):
        assert hasattr(lib, s), s


def test_reference_host_program_links_against_the_plugin_abi(tmp_path):
    """Where the reference tree exists: its unmodified v2 host program + loader compile, and its
    PCIE_Load() finds all twelve symbols in our plugin (it then stops at PCIE_Open on a box
    without a GPU -- the same message it prints without a board)."""
    comm = "/root/reference/Multiplier_NTT_Based/Software_Hardware_Comunnicator/linux_app"
    if not os.path.isdir(comm):
        pytest.skip("reference tree not present")
    gen = "/root/reference/Multiplier_NTT_Based/NTT_Software/Generator_Params"
    exe = str(tmp_path / "ntt_pcie_v2")
    # the program includes "NTT_Software/Generator_Params/generate_params.h": give it that path
    inc = tmp_path / "inc" / "NTT_Software"
    inc.mkdir(parents=True)
    os.symlink(gen, inc / "Generator_Params")
    subprocess.run(["gcc", "-O1", "-w", "-I", comm, "-I", str(tmp_path / "inc"), "-o", exe,
                    f"{comm}/NTT_PCIECommunicationv2.c", f"{comm}/PCIE.c",
                    "-x", "c", f"{gen}/generate_params.C", f"{gen}/prime_generate.C", f"{gen}/helper.C",
                    "-x", "none", "-ldl", "-lm"], check=True)
    os.symlink(os.path.join(PKG_DIR, "terasic_pcie_qsys.so"), tmp_path / "terasic_pcie_qsys.so")
    env = dict(os.environ, LD_LIBRARY_PATH=PKG_DIR)
    out = subprocess.run([exe], cwd=tmp_path, capture_output=True, text=True, env=env, timeout=120)
    text = out.stdout + out.stderr
    assert "PCIE_Load failed" not in text and "Load ./terasic_pcie_qsys.so error" not in text
    import torch
    if torch.cuda.is_available():
        assert "0 erros encontrados" in text
    else:
        assert "PCIE_Open failed" in text


# ------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_cli_stdout_matches_the_reference_binary(gpu, golden, tmp_path):
    exe = os.path.join(PKG_DIR, "nttb200_time_testing256")
    np.savetxt(tmp_path / "coeficientes_a.txt", golden["fixture_a"][None], fmt="%d")
    np.savetxt(tmp_path / "coeficientes_b.txt", golden["fixture_b"][None], fmt="%d")
    want = open(os.path.join(GOLD, "time_testing256_stdout.txt")).read()
    for variant in ("4", "1", "red1", "red4"):
        out = subprocess.run([exe, "--variant", variant, "--batch", "4096"], cwd=tmp_path, capture_output=True,
                             text=True, timeout=300)
        assert out.returncode == 0, out.stderr
        got = "\n".join(ln for ln in out.stdout.splitlines() if not ln.startswith("Tempo total")) + "\n"
        if variant == "4":
            assert got == want
        else:   # same C block, other banner
            assert got.split("Polinômio C")[1] == want.split("Polinômio C")[1]
        assert "rows equal the single result" in out.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("n,q,cyclic", [(256, 12289, False), (256, 7681, False), (1024, 12289, False),
                                        (256, 3329, True), (4096, 40961, False)])
def test_terasic_plugin_protocol(gpu, oracle, tmp_path, n, q, cyclic):
    """mode 0 (params) / 1 (A) / 2 (B) / 3 (GO) / FIFO read through dlopen, as the reference host does."""
    exe = str(tmp_path / "terasic_client")
    subprocess.run(["gcc", "-O1", "-o", exe, os.path.join(ROOT, "tests", "terasic_client.c"), "-ldl"], check=True)
    plugin = os.path.join(PKG_DIR, "terasic_pcie_qsys.so")
    env = dict(os.environ)
    if cyclic:
        env["NTTB200_SHIM_CYCLIC"] = "1"
    cases = []
    a = np.zeros(n, np.int64); b = np.zeros(n, np.int64)
    a[:3] = (1, 2, 3); b[0] = 2                              # KAT of ...v2.c:157-158, 232-238: C = 2,4,6
    cases.append((a, b))
    a2 = np.zeros(n, np.int64); b2 = np.zeros(n, np.int64)
    a2[:3] = (1, 2, 3); b2[:2] = (2, 2)                      # KAT of HW/NTT_PolyMul_test.v:110-192: 2,6,10,6
    cases.append((a2, b2))
    cases.append((oracle.random((1, n), q, 11)[0].astype(np.int64), oracle.random((1, n), q, 12)[0].astype(np.int64)))
    for i, (x, y) in enumerate(cases):
        fin, fout = tmp_path / f"in{i}.txt", tmp_path / f"out{i}.txt"
        np.savetxt(fin, np.concatenate([x, y])[None], fmt="%d")
        r = subprocess.run([exe, plugin, str(n), str(q), str(fin), str(fout)], env=env, capture_output=True,
                           text=True, timeout=300)
        assert r.returncode == 0, (r.returncode, r.stderr)
        got = np.loadtxt(fout, dtype=np.int64)
        want = oracle.product(n, q, x[None].astype(np.int32), y[None].astype(np.int32), 30 if cyclic else 10)[0]
        assert (got == want).all()
        if i == 0:
            assert list(got[:4]) == [2, 4, 6, 0]
        if i == 1:
            assert list(got[:5]) == [2, 6, 10, 6, 0]


# ------------------------------------------------------------------------------------------
# the reference's OWN programs, unmodified, on the GPU library (prebuilt by oracle/Makefile into
# oracle/_ref/, which travels to the GPU box; /root/reference is not needed at run time)
REFBIN = os.path.join(ROOT, "oracle", "_ref")


def _refbin(name):
    path = os.path.join(REFBIN, name)
    assert os.path.exists(path), f"{path} missing: run __graft_entry__.build() where /root/reference exists"
    return path


def _write_coeffs(L, path, v):
    v = np.ascontiguousarray(v, dtype=np.int32)
    assert L.nttb200_write_coeff_file(str(path).encode(), v.ctypes.data, v.size) == 0


@pytest.mark.gpu
def test_reference_benchmark_main_relinked_against_the_gpu_library(gpu, L, golden, tmp_path):
    """NTT-256/time_testing256.c (its main(), file reader and printer, unmodified) linked with
    -lnttb200 instead of the reference's C: same stdout as the reference binary, the time aside."""
    _write_coeffs(L, tmp_path / "coeficientes_a.txt", golden["fixture_a"])
    _write_coeffs(L, tmp_path / "coeficientes_b.txt", golden["fixture_b"])
    out = subprocess.run([_refbin("time_testing256_gpu")], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    got = "\n".join(ln for ln in out.stdout.splitlines() if not ln.startswith("Tempo total")) + "\n"
    assert got == open(os.path.join(GOLD, "time_testing256_stdout.txt")).read()
    assert "Tempo total" in out.stdout


@pytest.mark.gpu
def test_reference_kat_program_relinked_against_the_gpu_library(gpu, tmp_path):
    """NTT/test_prod_ntt256.c:48-57: (1 + 2x) * 3 = 3 + 6x, printed by the reference's own main()."""
    out = subprocess.run([_refbin("test_prod_ntt256_gpu")], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    assert out.stdout == open(os.path.join(GOLD, "test_prod_ntt256_stdout.txt")).read()
    nums = [int(t) for t in out.stdout.split("Polin")[1].split(":")[1].split()]
    assert nums[:3] == [3, 6, 0] and len(nums) == 256 and not any(nums[2:])


@pytest.mark.gpu
def test_reference_fpga_host_program_runs_on_the_plugin(gpu, tmp_path):
    """linux_app/NTT_PCIECommunicationv2.c + PCIE.c, unmodified: PCIE_Load() dlopens
    ./terasic_pcie_qsys.so (our plugin), streams W/W_INV/q/n_inv, A, B, pulses GO, polls STATUS,
    reads C back and checks {2, 4, 6} itself (v2.c:232-238)."""
    os.symlink(os.path.join(PKG_DIR, "terasic_pcie_qsys.so"), tmp_path / "terasic_pcie_qsys.so")
    # the plugin finds libnttb200.so next to itself ($ORIGIN); through the symlink it needs the path
    env = dict(os.environ, LD_LIBRARY_PATH=PKG_DIR + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""))
    out = subprocess.run([_refbin("ntt_pcie_v2")], cwd=tmp_path, capture_output=True, text=True, timeout=300, env=env)
    text = out.stdout + out.stderr
    assert out.returncode == 0 and "PCIE_Load failed" not in text and "PCIE_Open failed" not in text, text
    assert "TB: Sinal 'done_all' recebido!" in text
    assert "0 erros encontrados" in text


WRAPPER_IDS = list(range(13)) + list(range(100, 112))


@pytest.mark.gpu
def test_reference_header_wrappers_on_the_gpu_library(gpu, golden, tmp_path):
    """tests/ref_wrappers_client.c, compiled against the reference's ntt256.h / ntt_red256.h and
    linked with -lnttb200 only: every `static inline` wrapper (ntt256.h:20-69, ntt_red256.h:21-70)
    binds one of OUR exported tables to one of OUR generic transforms; outputs = the compiled
    reference's (golden transform_*), products = fixture_c."""
    exe = _refbin("ref_wrappers_gpu")
    fin, fout = tmp_path / "in.txt", tmp_path / "out.txt"
    for tid in WRAPPER_IDS:
        row = 5
        src = golden["rand_a"][row] if tid < 100 else golden["red_transform_in"][row]
        np.savetxt(fin, src[None], fmt="%d")
        r = subprocess.run([exe, str(tid), str(fin), str(fout)], capture_output=True, text=True, timeout=120)
        assert r.returncode == 0, (tid, r.stderr)
        assert (np.loadtxt(fout, dtype=np.int64) == golden[f"transform_{tid}"][row]).all(), tid
    np.savetxt(fin, np.concatenate([golden["fixture_a"], golden["fixture_b"]])[None], fmt="%d")
    for pid in (200, 201, 202, 203):
        r = subprocess.run([exe, str(pid), str(fin), str(fout)], capture_output=True, text=True, timeout=120)
        assert r.returncode == 0, (pid, r.stderr)
        assert (np.loadtxt(fout, dtype=np.int64) == golden["fixture_c"]).all(), pid
