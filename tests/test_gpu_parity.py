"""GPU (B200): the CUDA path, called through the C ABI, bit-exact against
  * the committed golden vectors produced by the compiled reference C,
  * the CPU oracle on the same seeded inputs (sizes the oracle finishes in seconds),
  * the compiled reference itself (oracle/_ref travels to the box) on a 2^16 batch,
and, at BASELINE.json's full batch sizes, through size-independent properties
(linearity, delta rows, commutativity, checksum against sampled oracle rows)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N, Q, PSI = 256, 12289, 1002
SEED = 0x4E545442323030

SMALL_CASES = [(8, 17), (16, 97), (32, 193), (64, 257), (128, 3329), (256, 12289), (256, 7681),
               (512, 12289), (1024, 12289),
               (256, 8380417), (1024, 8380417),          # Dilithium prime: HARVEY class
               (256, 469762049), (1024, 998244353),      # 29/30-bit primes: HARVEY class
               (256, 2013265921), (1024, 2013265921)]    # 31-bit prime: CANON class


@pytest.fixture(scope="module")
def plan256(gpu):
    p = gpu.Plan(N, Q, PSI)
    yield p
    p.close()


def test_native_library_is_the_one_running(gpu, plan256):
    assert "fused-small" in plan256.describe() and "plantard" in plan256.describe()
    a = np.zeros((2, N), np.int32)
    plan256.polymul(a, a)
    assert gpu.last_launch_count() >= 1


def test_golden_fixture_and_kats(golden, plan256):
    assert (plan256.polymul(golden["fixture_a"], golden["fixture_b"])[0] == golden["fixture_c"]).all()
    assert (plan256.polymul(golden["kat_a"], golden["kat_b"]) == golden["kat_c"]).all()


def test_golden_random_rows(golden, plan256):
    assert (plan256.polymul(golden["rand_a"], golden["rand_b"]) == golden["rand_c"]).all()


def test_product_does_not_depend_on_psi(gpu, golden):
    for psi in (0, 3, 1002, 10805):
        p = gpu.Plan(N, Q, psi)
        assert (p.polymul(golden["rand_a"][:8], golden["rand_b"][:8]) == golden["rand_c"][:8]).all()
        p.close()


@pytest.mark.parametrize("batch", [1, 2, 3, 15, 16, 17, 31, 33, 1000])
def test_ragged_batches(plan256, oracle, batch):
    a, b = oracle.random((batch, N), Q, SEED + batch), oracle.random((batch, N), Q, SEED - batch)
    assert (plan256.polymul(a, b) == oracle.product(N, Q, a, b, 10, PSI)).all()


@pytest.mark.parametrize("n,q", [(256, 12289), (1024, 2013265921), (2048, 12289)])
def test_range_check_on_request(gpu, oracle, n, q):
    """NTTB200_PLAN_CHECK_RANGE: the reference's unchecked precondition "elements in [0 .. Q-1]"
    (ntt256.h:82-83) is verified first and NTTB200_ERANGE names the first offender -- host and
    device buffers, products and transforms; without the flag nothing is looked at."""
    import torch
    p = gpu.Plan(n, q, check_range=True)
    a, b = oracle.random((5, n), q, 1), oracle.random((5, n), q, 2)
    assert (p.polymul(a, b) == oracle.product(n, q, a, b, 10)).all()
    for arr, name, val in ((a, "a", q), (b, "b", -1)):
        bad = arr.copy()
        bad[3, 7] = val
        with pytest.raises(gpu.NttError) as e:
            p.polymul(bad if name == "a" else a, bad if name == "b" else b)
        assert "error -4" in str(e.value) and f"{name}[{3 * n + 7}]" in str(e.value), str(e.value)
        with pytest.raises(gpu.NttError) as e:
            p.transform("ntt_std2rev", bad)
        assert "error -4" in str(e.value)
        da, db = torch.from_numpy(bad if name == "a" else a).cuda(), torch.from_numpy(bad if name == "b" else b).cuda()
        dc = torch.empty_like(da)
        with pytest.raises(gpu.NttError) as e:
            p.polymul_dev(dc.data_ptr(), da.data_ptr(), db.data_ptr(), 5)
        assert "error -4" in str(e.value) and f"{name}[{3 * n + 7}]" in str(e.value), str(e.value)
    da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    dc = torch.empty_like(da)
    p.polymul_dev(dc.data_ptr(), da.data_ptr(), db.data_ptr(), 5)
    torch.cuda.synchronize()
    assert (dc.cpu().numpy() == oracle.product(n, q, a, b, 10)).all()
    p.close()


def test_empty_batch(plan256):
    out = plan256.polymul(np.zeros((0, N), np.int32), np.zeros((0, N), np.int32))
    assert out.shape == (0, N)


def test_inputs_are_left_untouched(plan256, oracle):
    a, b = oracle.random((4, N), Q, 1), oracle.random((4, N), Q, 2)
    a0, b0 = a.copy(), b.copy()
    plan256.polymul(a, b)
    assert (a == a0).all() and (b == b0).all()


def test_config2_batch_2e16_vs_all_four_reference_variants(plan256, oracle, loader, nttb200, golden):
    """BASELINE config 2: batch 2^16 at the reference default (n,q) on the SURVEY 8d batch (splitmix64
    stream, edge rows 0..7, the reference's coefficient files in row 8), EVERY row bit-exact vs all
    four C variants of the compiled, unmodified reference."""
    batch = 1 << 16
    ta, tb = nttb200.inputs.survey_batch(N, Q, batch, 2, fixture=(golden["fixture_a"], golden["fixture_b"]))
    a, b = ta.numpy(), tb.numpy()
    assert (a[8] == golden["fixture_a"]).all() and not a[0].any() and (b[1] == Q - 1).all()
    got = plan256.polymul(a, b)
    assert (got[8] == golden["fixture_c"]).all() and (got[4:8] == golden["kat_c"]).all()
    if loader.reference_available():
        ref = loader.Reference()
        for v in (loader.REF_CT, loader.REF_GS, loader.REF_RED_CT, loader.REF_RED_GS):
            assert (ref.product(a, b, v) == got).all(), v
    else:                                   # never on the build box; keeps the test meaningful elsewhere
        assert (oracle.product(N, Q, a[:4096], b[:4096], 10, PSI) == got[:4096]).all()
    assert got.min() >= 0 and got.max() < Q


@pytest.mark.parametrize("n,q", [(8, 17), (16, 97), (32, 193), (64, 257), (128, 3329), (256, 12289),
                                 (256, 7681), (512, 12289), (1024, 12289), (256, 10753), (128, 12289)])
@pytest.mark.parametrize("signed", [0, 1, 2])
def test_half_word_moduli_plantard_vs_shoup_kernels(gpu, oracle, monkeypatch, n, q, signed):
    """q <= 12385: the product runs a Plantard kernel -- signed = 1: the default signed kernels
    (ntt_small_splant.cuh, ntt_splant_wide.cuh at n = 1024); 0: the unsigned kernel of
    ntt_small_plant.cuh (NTTB200_PLANT_SIGNED=0); 2: at n >= 512 the OTHER signed kernel of the size --
    the same plan with NTTB200_PLAN_NO_PLANTARD runs the Shoup/Montgomery kernel.  All against the oracle, with worst-case rows (all q-1 and alternating 0 / q-1: every
    lazy bound is attained) and ragged batch sizes."""
    if signed == 2:          # n >= 512: the other signed kernel of each size (one layout per phase at n = 1024,
        if n < 512:          # three layouts at n = 512)
            pytest.skip("one signed kernel below n = 512")
        monkeypatch.setenv("NTTB200_PLANT_N1024", "0" if n == 1024 else "2")
    monkeypatch.setenv("NTTB200_PLANT_SIGNED", "1" if signed else "0")
    pl, sh = gpu.Plan(n, q), gpu.Plan(n, q, no_plantard=True)
    assert "plantard" in pl.describe() and "plantard" not in sh.describe()
    for batch in (1, 2, 31, 32, 33, 257, 1031):
        a, b = oracle.random((batch, n), q, SEED + n + batch), oracle.random((batch, n), q, SEED + q + batch)
        a[0], b[0] = q - 1, q - 1
        if batch > 2:
            a[1] = q - 1
            b[2] = q - 1
        if batch > 8:
            alt = np.where(np.arange(n) % 2 == 0, q - 1, 0).astype(np.int32)
            a[3], b[3] = alt, alt
            a[4], b[4] = alt, (q - 1) - alt
            a[5] = np.where(np.arange(n) < n // 2, q - 1, 0)
            b[5] = np.where(np.arange(n) % 4 < 2, q - 1, 0)
        want = oracle.product(n, q, a, b, 10)
        assert (pl.polymul(a, b) == want).all(), (pl.describe(), batch)
        assert (sh.polymul(a, b) == want).all(), (sh.describe(), batch)
    pl.close()
    sh.close()


def test_signed_plantard_kernel_on_config2_vs_the_compiled_reference(gpu, loader, nttb200, golden, monkeypatch):
    """NTTB200_PLANT_SIGNED=1 (ntt_small_splant.cuh) on the whole SURVEY 8d batch of config 2: every
    one of the 2^16 rows equals the compiled reference (optimized CT and plain GS variants), device-
    resident and through the host-buffer call (16-bit wire included)."""
    import torch
    monkeypatch.setenv("NTTB200_PLANT_SIGNED", "1")
    batch = 1 << 16
    p = gpu.Plan(N, Q, PSI)
    a, b = nttb200.inputs.survey_batch(N, Q, batch, 2, device="cuda", fixture=(golden["fixture_a"], golden["fixture_b"]))
    c = torch.empty_like(a)
    p.polymul_dev(c.data_ptr(), a.data_ptr(), b.data_ptr(), batch, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    ha, hb, got = a.cpu().numpy(), b.cpu().numpy(), c.cpu().numpy()
    assert (got[8] == golden["fixture_c"]).all() and (got[4:8] == golden["kat_c"]).all()
    if loader.reference_available():
        ref = loader.Reference()
        for v in (loader.REF_RED_CT, loader.REF_GS):
            assert (ref.product(ha, hb, v) == got).all(), v
    assert (p.polymul(ha, hb) == got).all()
    p.close()


@pytest.mark.parametrize("n,q", SMALL_CASES)
def test_products_other_sizes_and_moduli(gpu, oracle, n, q):
    batch = 67
    p = gpu.Plan(n, q)
    a, b = oracle.random((batch, n), q, SEED + n), oracle.random((batch, n), q, SEED + q)
    a[0], b[0] = 0, 0
    a[1], b[1] = q - 1, q - 1
    a[2] = 0
    a[2][0] = 1                                          # delta_0 * b = b
    a[3], b[3] = 0, 0
    a[3][n - 1], b[3][n - 1] = 1, 1                      # x^(n-1) x^(n-1) = -x^(n-2)
    got = p.polymul(a, b)
    want = oracle.product(n, q, a, b, 10)
    assert (got == want).all(), p.describe()
    assert (got[2] == b[2]).all()
    assert got[3][n - 2] == q - 1 and got[3].sum() == q - 1
    if n <= 256:
        assert (oracle.product(n, q, a[:8], b[:8], 20) == got[:8]).all()     # O(n^2) definition
    p.close()


LARGE_CASES = [(2048, 12289), (4096, 40961), (8192, 65537), (16384, 65537),      # LAZY class
               (2048, 8380417), (32768, 786433), (65536, 786433), (65536, 469762049),  # HARVEY class
               (2048, 2013265921), (65536, 2013265921), (131072, 2013265921)]     # CANON class


FUSED_N = (32768, 65536)      # one persistent cluster kernel (ntt_large_fused.cuh)


@pytest.mark.parametrize("fused", [2, 1, 0])
@pytest.mark.parametrize("n,q", LARGE_CASES)
def test_large_n_multipass_products(gpu, oracle, monkeypatch, n, q, fused):
    """n > 1024: column pass / row pass / column pass (ntt_large.cuh) against the oracle's
    merged CT-fwd/GS-inv pipeline; ragged batch (odd, not a multiple of the CTA's 16 rows).
    n = 2^15, 2^16 can also run the three passes inside ONE persistent kernel (1 launch): the ticket
    dataflow kernel (NTTB200_LARGE_FUSED=2) or the cluster kernel (=1) -- both measured slower than the
    three-launch pipeline, which is the default (NTTB200_LARGE_FUSED=0)."""
    if fused != 2 and n not in FUSED_N:
        pytest.skip("only n = 2^15, 2^16 have the fused kernels")
    monkeypatch.setenv("NTTB200_LARGE_FUSED", str(fused))
    batch = 21 if n <= 16384 else 5
    p = gpu.Plan(n, q)
    assert "large(multi-pass)" in p.describe()
    a, b = oracle.random((batch, n), q, SEED + n), oracle.random((batch, n), q, SEED + q)
    a[0], b[0] = 0, 0
    a[1], b[1] = q - 1, q - 1
    a[2] = 0
    a[2][0] = 1
    a[3], b[3] = 0, 0
    a[3][n - 1], b[3][n - 1] = 1, 1
    got = p.polymul(a, b)
    want = oracle.product(n, q, a, b, 10)
    assert (got == want).all(), p.describe()
    assert (got[2] == b[2]).all()
    assert got[3][n - 2] == q - 1 and int(got[3].astype(np.int64).sum()) == q - 1
    assert gpu.last_launch_count() == (1 if (fused and n in FUSED_N) else 3)
    if fused == 2 and n == 65536:
        # more polynomials than scratch slots (48) and than the delay between a polynomial's passes
        batch = 131
        a, b = oracle.random((batch, n), q, SEED + 5), oracle.random((batch, n), q, SEED + 6)
        assert (p.polymul(a, b) == oracle.product(n, q, a, b, 10)).all()
    p.close()


@pytest.mark.parametrize("n,q", [(65536, 2013265921), (65536, 469762049), (65536, 786433), (32768, 2013265921),
                                 (32768, 786433)])
@pytest.mark.parametrize("mode", [2, 1])
def test_fused_kernels_many_polynomials(gpu, oracle_mt, nttb200, monkeypatch, n, q, mode):
    """More polynomials than resident clusters / scratch slots (every cluster loops and the split
    barrier between one polynomial's pass 3 and the next one's pass 1 is exercised; the dataflow
    kernel recycles every scratch slot several times), device-resident, every row against the
    oracle, twice in a row on the same scratch, and equal to the three-launch pipeline."""
    import os
    import torch
    monkeypatch.setenv("NTTB200_LARGE_FUSED", str(mode))
    batch = 301
    p = gpu.Plan(n, q)
    a, b = nttb200.inputs.survey_batch(n, q, batch, 5, device="cuda")
    c = torch.empty_like(a)
    st = torch.cuda.current_stream().cuda_stream
    want = oracle_mt(n, q, a.cpu().numpy(), b.cpu().numpy(), 10)
    for _ in range(2):
        c.zero_()
        p.polymul_dev(c.data_ptr(), a.data_ptr(), b.data_ptr(), batch, st)
        torch.cuda.synchronize()
        assert gpu.last_launch_count() == 1
        got = c.cpu().numpy()
        assert (got == want).all(), np.nonzero((got != want).any(axis=1))[0][:8]
    monkeypatch.setenv("NTTB200_LARGE_FUSED", "0")
    c2 = torch.empty_like(a)
    p.polymul_dev(c2.data_ptr(), a.data_ptr(), b.data_ptr(), batch, st)
    torch.cuda.synchronize()
    assert gpu.last_launch_count() > 1 and bool((c2 == c).all())
    p.close()


@pytest.mark.parametrize("n,q", [(2048, 12289), (16384, 65537), (65536, 469762049), (65536, 2013265921)])
def test_large_n_standalone_transforms(gpu, oracle, loader, n, q):
    p = gpu.Plan(n, q)
    psi = p.psi
    a = oracle.random((3, n), q, SEED + 3 * n + q)
    a[1] = q - 1
    T = lambda k: oracle.table(k, n, q, psi)
    checks = [
        ("ntt_std2rev", "ntt_ct_std2rev", loader.OMEGA_POWERS_REV),
        ("mulntt_std2rev", "mulntt_ct_std2rev", loader.MIXED_POWERS_REV),
        ("intt_rev2std", "ntt_gs_rev2std", loader.INV_OMEGA_POWERS_REV),
        ("inttmul_rev2std", "nttmul_gs_rev2std", loader.INV_MIXED_POWERS_REV),
        ("intt_std2rev", "ntt_ct_std2rev", loader.INV_OMEGA_POWERS_REV),
        ("ntt_rev2std", "ntt_gs_rev2std", loader.OMEGA_POWERS_REV),
    ]
    for kind, fn, tab in checks:
        assert (p.transform(kind, a) == oracle.transform(fn, a, T(tab), q)).all(), (kind, fn)
    back = p.transform("inttmul_rev2std_scaled", p.transform("mulntt_std2rev", a))
    assert (back == a).all()
    p.close()


def test_config5_full_batch_every_row(gpu, oracle_mt, nttb200):
    """BASELINE config 5: n=2^16, 31-bit prime, batch 2^10, device-resident, on the SURVEY 8d batch
    (splitmix64 stream, edge rows 0..7): EVERY row against the oracle, plus commutativity."""
    import torch
    n, q, batch = 65536, 2013265921, 1 << 10
    p = gpu.Plan(n, q)
    a, b = nttb200.inputs.survey_batch(n, q, batch, 5, device="cuda")
    c = torch.empty_like(a)
    st = torch.cuda.current_stream().cuda_stream
    p.polymul_dev(c.data_ptr(), a.data_ptr(), b.data_ptr(), batch, st)
    torch.cuda.synchronize()
    assert int(c.min()) >= 0 and int(c.max()) < q
    assert bool((c[0] == 0).all()) and bool((c[2] == b[2]).all())
    assert int(c[3, n - 2]) == q - 1 and int(c[3].long().sum()) == q - 1
    assert c[5, :4].tolist() == [2, 4, 6, 0] and c[6, :5].tolist() == [2, 6, 10, 6, 0]
    want = oracle_mt(n, q, a.cpu().numpy(), b.cpu().numpy(), 10)
    got = c.cpu().numpy()
    assert (got == want).all(), np.nonzero((got != want).any(axis=1))[0][:8]
    c2 = torch.empty_like(a)
    p.polymul_dev(c2.data_ptr(), b.data_ptr(), a.data_ptr(), batch, st)
    torch.cuda.synchronize()
    assert bool((c == c2).all())
    p.close()


@pytest.mark.parametrize("n,q", [(256, 12289), (256, 7681), (1024, 12289), (512, 12289), (128, 3329),
                                 (256, 998244353), (256, 2013265921)])
def test_standalone_transforms_match_reference_functions(gpu, oracle, loader, n, q):
    psi = 1002 if (n, q) == (256, 12289) else 0
    p = gpu.Plan(n, q, psi)
    psi = p.psi
    a = oracle.random((37, n), q, SEED + 5 * n + q)
    a[0] = 0
    a[1] = q - 1
    T = lambda k: oracle.table(k, n, q, psi)
    checks = [
        ("ntt_std2rev", "ntt_ct_std2rev", loader.OMEGA_POWERS_REV),
        ("ntt_std2rev", "ntt_gs_std2rev", loader.OMEGA_POWERS),
        ("mulntt_std2rev", "mulntt_ct_std2rev", loader.MIXED_POWERS_REV),
        ("intt_rev2std", "ntt_gs_rev2std", loader.INV_OMEGA_POWERS_REV),
        ("intt_rev2std", "ntt_ct_rev2std", loader.INV_OMEGA_POWERS),
        ("inttmul_rev2std", "nttmul_gs_rev2std", loader.INV_MIXED_POWERS_REV),
        ("intt_std2rev", "ntt_ct_std2rev", loader.INV_OMEGA_POWERS_REV),
        ("intt_std2rev", "ntt_gs_std2rev", loader.INV_OMEGA_POWERS),
        ("ntt_rev2std", "ntt_ct_rev2std", loader.OMEGA_POWERS),
        ("ntt_rev2std", "ntt_gs_rev2std", loader.OMEGA_POWERS_REV),
    ]
    for kind, fn, tab in checks:
        assert (p.transform(kind, a) == oracle.transform(fn, a, T(tab), q)).all(), (kind, fn)
    ninv = pow(n, q - 2, q)
    for kind, fn, tab in (("intt_rev2std_scaled", "ntt_gs_rev2std", loader.INV_OMEGA_POWERS_REV),
                          ("inttmul_rev2std_scaled", "nttmul_gs_rev2std", loader.INV_MIXED_POWERS_REV)):
        want = oracle.transform(fn, a, T(tab), q).astype(np.int64) * ninv % q
        assert (p.transform(kind, a) == want).all(), kind
    # round trip: intt(ntt(a)) = n * a   (R/NTT/ntt256.h:16-17)
    back = p.transform("intt_rev2std", p.transform("ntt_std2rev", a))
    assert (back == a.astype(np.int64) * n % q).all()
    back = p.transform("inttmul_rev2std_scaled", p.transform("mulntt_std2rev", a))
    assert (back == a).all()
    p.close()


def test_standalone_transforms_golden(golden, plan256):
    a = golden["rand_a"][:16]
    assert (plan256.transform("ntt_std2rev", a) == golden["transform_2"]).all()
    assert (plan256.transform("ntt_std2rev", a) == golden["transform_3"]).all()
    assert (plan256.transform("mulntt_std2rev", a) == golden["transform_9"]).all()
    assert (plan256.transform("intt_rev2std", a) == golden["transform_5"]).all()
    assert (plan256.transform("intt_rev2std", a) == golden["transform_4"]).all()
    assert (plan256.transform("inttmul_rev2std", a) == golden["transform_10"]).all()
    assert (plan256.transform("intt_std2rev", a) == golden["transform_6"]).all()
    assert (plan256.transform("ntt_rev2std", a) == golden["transform_0"]).all()
    assert (plan256.transform("ntt_rev2std", a) == golden["transform_1"]).all()


def test_hw_golden_vectors_q7681(gpu, hw_golden):
    """Verilog testbench vectors (q=7681, omega=0xf04): NTT_DOUT = gs_std2rev(NTT_DIN)."""
    n, q, w = 256, 7681, int(hw_golden["PARAM"][2])
    psi = int(hw_golden["PARAM"][4])
    p = gpu.Plan(n, q, psi)
    assert p.psi * p.psi % q == w
    out = p.transform("ntt_std2rev", hw_golden["NTT_DIN"].astype(np.int32))
    assert (out[0] == hw_golden["NTT_DOUT"]).all()
    iout = p.transform("intt_std2rev", hw_golden["INTT_DIN"].astype(np.int32)).astype(np.int64)
    assert (iout[0] * pow(n, q - 2, q) % q == hw_golden["INTT_DOUT"]).all()
    p.close()


# dataflow -> (un-merged reference function, psi-merged twin, ids of ref_shim.c:ref_transform_tab)
TABLE_DATAFLOWS = {"ct_rev2std": ("ntt_ct_rev2std", "mulntt_ct_rev2std", 1, 2),
                   "ct_std2rev": ("ntt_ct_std2rev", "mulntt_ct_std2rev", 3, 4),
                   "gs_rev2std": ("ntt_gs_rev2std", "nttmul_gs_rev2std", 5, 6),
                   "gs_std2rev": ("ntt_gs_std2rev", "nttmul_gs_std2rev", 7, 8)}


@pytest.mark.parametrize("df", sorted(TABLE_DATAFLOWS))
def test_table_driven_dataflows(gpu, oracle, golden, df):
    """nttb200_ntt_table_batch: the reference dataflow with an ARBITRARY caller table (p[t] != 1),
    against the compiled reference's outputs (golden arb_*) for both the un-merged (j = 0 peeled,
    p[t] never read -- R/NTT/ntt.C:313-317) and the psi-merged entry points, and against the
    oracle at other (n, q)."""
    plain, merged, id_plain, id_merged = TABLE_DATAFLOWS[df]
    for nn in (256, 64):
        tab, a = golden[f"arb_table_{nn}"].astype(np.uint32), golden[f"arb_in_{nn}"]
        assert (gpu.ntt_table_batch(nn, Q, df, tab, a, skip_j0=True) == golden[f"arb_transform_{nn}_{id_plain}"]).all()
        assert (gpu.ntt_table_batch(nn, Q, df, tab, a, skip_j0=False) == golden[f"arb_transform_{nn}_{id_merged}"]).all()
        # the reference's own names (legacy surface) with the same caller table
        for fn, tid in ((plain, id_plain), (merged, id_merged)):
            got = gpu.legacy.transform(fn, a[3], tab)
            assert (got == golden[f"arb_transform_{nn}_{tid}"][3]).all(), fn
    for n, q in ((256, 12289), (64, 257), (2048, 12289), (8192, 12289)):
        tab = (oracle.random((n,), q - 2, SEED + n) + 2).astype(np.uint32)      # not even roots of unity
        a = oracle.random((5, n), q, SEED + 7)
        assert (gpu.ntt_table_batch(n, q, df, tab, a, skip_j0=True) == oracle.transform(plain, a, tab, q)).all(), (n, q)
        assert (gpu.ntt_table_batch(n, q, df, tab, a, skip_j0=False) == oracle.transform(merged, a, tab, q)).all(), (n, q)


def test_v1_table_layout_with_an_arbitrary_table(gpu, golden):
    """ntt_ct_rev2std_v1 reads psi-power entries p[j*l] and peels j = 0 (R/NTT/ntt.C:168-197)."""
    for nn in (256, 64):
        got = gpu.legacy.transform("ntt_ct_rev2std_v1", golden[f"arb_in_{nn}"][2], golden[f"arb_table_{nn}"])
        assert (got == golden[f"arb_transform_{nn}_0"][2]).all()


def test_legacy_surface(gpu, golden, oracle):
    """The reference's own function names, one polynomial per call."""
    fa, fb, fc = golden["fixture_a"], golden["fixture_b"], golden["fixture_c"]
    for name in ("ntt256_product1", "ntt256_product4", "ntt_red256_product1", "ntt_red256_product4"):
        c, a_after, b_after = gpu.legacy.product(name, fa, fb)
        assert (c == fc).all(), name
        assert (a_after == fa).all() and (b_after == fb).all()
    # opt-in reference post-state of a, b
    c, a_after, b_after = gpu.legacy.product("ntt256_product1", golden["rand_a"][4], golden["rand_b"][4], clobber=True)
    assert (c == golden["rand_c"][4]).all()
    assert (a_after == golden["clobber_a_after_product1"]).all()
    assert (b_after == golden["clobber_b_after_product1"]).all()
    # ... and of the optimized pair: the UNREDUCED representative their pipeline leaves behind
    # ("a and b are modified", ntt_red256.h:77-86), golden from the compiled reference
    for name, v in (("ntt_red256_product1", 101), ("ntt_red256_product4", 104)):
        c, a_after, b_after = gpu.legacy.product(name, golden["rand_a"][4], golden["rand_b"][4], clobber=True)
        assert (c == golden["rand_c"][4]).all(), name
        assert (a_after == golden[f"clobber_a_after_product{v}"]).all(), name
        assert (b_after == golden[f"clobber_b_after_product{v}"]).all(), name
    assert (golden["clobber_a_after_product101"] != golden["clobber_a_after_product1"]).any()
    a = golden["rand_a"][0:16]
    names = {0: ("ntt_ct_rev2std", 3), 1: ("ntt_gs_rev2std", 4), 2: ("ntt_ct_std2rev", 4), 3: ("ntt_gs_std2rev", 3),
             4: ("ntt_ct_rev2std", 5), 5: ("ntt_gs_rev2std", 6), 6: ("ntt_ct_std2rev", 6), 7: ("ntt_gs_std2rev", 5),
             8: ("mulntt_ct_rev2std", 7), 9: ("mulntt_ct_std2rev", 8), 10: ("nttmul_gs_rev2std", 10),
             11: ("nttmul_gs_std2rev", 9), 12: ("ntt_ct_rev2std_v1", 0)}
    for tid, (fn, kind) in names.items():
        got = gpu.legacy.transform(fn, a[5], golden[f"table_{kind}"])
        assert (got == golden[f"transform_{tid}"][5]).all(), fn
    x, y = golden["rand_a"][7], golden["rand_b"][7]
    assert (gpu.legacy.mul_array(x, y) == x.astype(np.int64) * y % Q).all()
    assert (gpu.legacy.mul_array16(x, golden["table_0"]) == x.astype(np.int64) * golden["table_0"] % Q).all()
    assert (gpu.legacy.scalar_mul_array(x, 12241) == x.astype(np.int64) * 12241 % Q).all()


RED_TRANSFORMS = {   # golden transform id -> (reference function, red table kind); R/NTT-RED/ntt_red256.h:21-70
    100: ("ntt_red_ct_rev2std", 3), 101: ("ntt_red_gs_rev2std", 4), 102: ("ntt_red_ct_std2rev", 4),
    103: ("ntt_red_gs_std2rev", 3), 104: ("ntt_red_ct_rev2std", 5), 105: ("ntt_red_gs_rev2std", 6),
    106: ("ntt_red_ct_std2rev", 6), 107: ("ntt_red_gs_std2rev", 5), 108: ("mulntt_red_ct_rev2std", 7),
    109: ("mulntt_red_ct_std2rev", 8), 110: ("nttmul_red_gs_rev2std", 10), 111: ("nttmul_red_gs_std2rev", 9)}


def test_red_surface_exact_unreduced_outputs(gpu, golden, oracle):
    """The Longa-Naehrig functions return UNREDUCED signed int32 values (ntt_red256.h:18); the GPU
    emulation must reproduce them bit for bit: golden outputs of the compiled reference, then the
    oracle's restatement on other sizes and on worst-case inputs."""
    cin = golden["red_transform_in"]
    for tid, (fn, kind) in RED_TRANSFORMS.items():
        for r in (0, 5, 15):
            got = gpu.legacy.red_transform(fn, cin[r], golden[f"red_table_{kind}"])
            assert (got == golden[f"transform_{tid}"][r]).all(), fn
    # batch entry point, other sizes, extreme admissible inputs (|a| <= 21499, ntt_red.h:169-173)
    L = gpu.lib()
    for n in (8, 64, 256, 1024):
        psi = oracle.psi(n, Q, 0)
        a = oracle.random((9, n), 2 * 21499 + 1, SEED + n).astype(np.int32) - 21499
        a[0] = 21499
        a[1] = -21499
        for df, name, skip0, kind in ((0, "ct_std2rev", 1, 4), (0, "mulntt_ct_std2rev", 0, 8),
                                      (1, "gs_rev2std", 1, 6), (1, "nttmul_gs_rev2std", 0, 10),
                                      (2, "ct_rev2std", 1, 3), (2, "mulntt_ct_rev2std", 0, 7),
                                      (3, "gs_std2rev", 1, 5), (3, "nttmul_gs_std2rev", 0, 9)):
            tab = oracle.red_table(kind, n, psi).astype(np.int32)
            x = a.copy()
            assert L.nttb200_red_ntt_table_batch(n, df, skip0, tab.ctypes.data, x.ctypes.data, x.shape[0]) == 0
            assert (x == oracle.red_transform(name, a, tab)).all(), (n, name)


def test_red_helpers_and_permutations(gpu, golden, oracle):
    rng = np.random.default_rng(5)
    x = rng.integers(-(1 << 30), 1 << 30, 1000).astype(np.int32)
    x[:4] = (0, -1, 12288, -12289)
    Qc = 12289
    assert (gpu.legacy.red_helper("normalize", x) == x.astype(np.int64) % Qc).all()
    assert (gpu.legacy.red_helper("normalize_inv3", x) == x.astype(np.int64) * 8193 % Qc).all()
    red = lambda v: 3 * (v & 4095) - (v >> 12)
    assert (gpu.legacy.red_helper("reduce_array", x) == red(x.astype(np.int64))).all()
    assert (gpu.legacy.red_helper("reduce_array_twice", x) == red(red(x.astype(np.int64)))).all()
    c = rng.integers(0, Qc, 777).astype(np.int32)
    assert (gpu.legacy.red_helper("shift_array", c) == np.where(c > 6144, c - Qc, c)).all()
    r = rng.integers(-Qc, 2 * Qc, 777).astype(np.int32)
    assert (gpu.legacy.red_helper("correct", r) == r.astype(np.int64) % Qc).all()
    a, b = cen(rng, 500), cen(rng, 500)
    z = a.astype(np.int64) * b
    want = (3 * (z & 4095) - (z >> 12)).astype(np.int32)
    L = gpu.lib()
    out = np.zeros_like(a)
    L.mul_reduce_array(out.ctypes.data, a.size, a.ctypes.data, b.ctypes.data)
    assert (out == want).all()
    y = a.copy()
    L.scalar_mul_reduce_array(y.ctypes.data, y.size, 8822)       # ntt_red256_rescale8
    z = a.astype(np.int64) * 8822
    assert (y == 3 * (z & 4095) - (z >> 12)).all()
    t16 = golden["red_table_0"][:256].astype(np.int16)
    y = a[:256].copy()
    L.mul_reduce_array16(y.ctypes.data, 256, t16.ctypes.data)
    z = a[:256].astype(np.int64) * t16
    assert (y == 3 * (z & 4095) - (z >> 12)).all()
    # the reference's optimized product, composed from the exact pieces (ntt_red256.C:5-27)
    fa, fb = golden["fixture_a"].astype(np.int32), golden["fixture_b"].astype(np.int32)
    lg = gpu.legacy
    def fwd(v):
        v = lg.red_helper("shift_array", v)
        L.mul_reduce_array16(v.ctypes.data, 256, golden["red_table_0"].astype(np.int16).ctypes.data)
        v = lg.red_transform("ntt_red_ct_std2rev", v, golden["red_table_4"])
        return lg.red_helper("reduce_array", v)
    ha, hb = fwd(fa.copy()), fwd(fb.copy())
    hc = np.zeros(256, np.int32)
    L.mul_reduce_array(hc.ctypes.data, 256, ha.ctypes.data, hb.ctypes.data)
    hc = lg.red_helper("reduce_array_twice", hc)
    hc = lg.red_transform("ntt_red_ct_rev2std", hc, golden["red_table_5"])
    L.mul_reduce_array16(hc.ctypes.data, 256, golden["red_table_2"].astype(np.int16).ctypes.data)
    hc = lg.red_helper("correct", lg.red_helper("reduce_array_twice", hc))
    assert (hc == golden["fixture_c"]).all()
    # permutations
    for n in (2, 8, 256, 4096):
        v = rng.integers(0, 1 << 30, n).astype(np.int32)
        bits = n.bit_length() - 1
        rev = np.array([int(format(i, f"0{bits}b")[::-1], 2) for i in range(n)])
        assert (gpu.legacy.red_helper("bitrev_shuffle", v) == v[rev]).all()
    v = rng.integers(0, 1 << 30, 64).astype(np.int32)
    pairs = np.array([[1, 32], [2, 16], [3, 48], [5, 40]], dtype=np.uint16)
    w = v.copy()
    L.shuffle_with_table(w.ctypes.data, pairs.ctypes.data, len(pairs))
    e = v.copy()
    for j, k in pairs:
        e[j], e[k] = e[k], e[j]
    assert (w == e).all()


def cen(rng, n):
    return rng.integers(-6144, 6145, n).astype(np.int32)


def test_elementwise_batch(gpu, oracle):
    for n, q in ((256, 12289), (1024, 2013265921)):
        p = gpu.Plan(n, q)
        a, b = oracle.random((9, n), q, 3), oracle.random((9, n), q, 4)
        assert (p.mul_array(a, b) == a.astype(np.int64) * b % q).all()
        assert (p.scalar_mul_array(a, q - 2) == a.astype(np.int64) * (q - 2) % q).all()
        p.close()


def test_cyclic_plan_q3329(gpu, oracle):
    """q=3329 (Kyber): no 512-th root of unity -> the psi-free surface, product mod x^256 - 1."""
    with pytest.raises(gpu.NttError):
        gpu.Plan(256, 3329)
    p = gpu.Plan(256, 3329, cyclic=True)
    a, b = oracle.random((33, 256), 3329, 5), oracle.random((33, 256), 3329, 6)
    assert (p.polymul(a, b) == oracle.product(256, 3329, a, b, 30)).all()
    p.close()


@pytest.mark.parametrize("n,q", [(256, 12289), (1024, 12289), (256, 2013265921), (8192, 65537), (65536, 2013265921)])
def test_repeatability_under_load(gpu, oracle, n, q):
    """compute-sanitizer is closed on this pool, so races in the shared-memory exchanges / the
    cp.async prefetch are hunted the blunt way: the same device-resident batch, multiplied 12
    times while the SMs are fully loaded, must give bit-identical results every time, and the
    first/last rows must equal the oracle."""
    import torch
    batch = max(64, (1 << 22) // n)
    p = gpu.Plan(n, q)
    g = torch.Generator(device="cuda").manual_seed(7 + n)
    a = torch.randint(0, q, (batch, n), dtype=torch.int32, device="cuda", generator=g)
    b = torch.randint(0, q, (batch, n), dtype=torch.int32, device="cuda", generator=g)
    st = torch.cuda.current_stream().cuda_stream
    outs = []
    for _ in range(12):
        c = torch.empty_like(a)
        p.polymul_dev(c.data_ptr(), a.data_ptr(), b.data_ptr(), batch, st)
        outs.append(c)
    torch.cuda.synchronize()
    for c in outs[1:]:
        assert bool((c == outs[0]).all())
    idx = [0, 1, batch // 2, batch - 2, batch - 1]
    ti = torch.tensor(idx, device="cuda")
    want = oracle.product(n, q, a[ti].cpu().numpy(), b[ti].cpu().numpy(), 10)
    assert (outs[0][ti].cpu().numpy() == want).all()
    p.close()


@pytest.mark.parametrize("n,q", [(256, 12289), (1024, 12289), (256, 2013265921), (4096, 40961), (65536, 2013265921)])
def test_in_place_result_over_an_operand(gpu, oracle, n, q):
    """c == a or c == b (exactly the same buffer) is allowed on the device path: every kernel has
    consumed a polynomial's operand rows before it stores that polynomial's result."""
    import torch
    batch = max(37, (1 << 20) // n + 5)
    p = gpu.Plan(n, q)
    g = torch.Generator(device="cuda").manual_seed(11 + n)
    a = torch.randint(0, q, (batch, n), dtype=torch.int32, device="cuda", generator=g)
    b = torch.randint(0, q, (batch, n), dtype=torch.int32, device="cuda", generator=g)
    st = torch.cuda.current_stream().cuda_stream
    c = torch.empty_like(a)
    p.polymul_dev(c.data_ptr(), a.data_ptr(), b.data_ptr(), batch, st)
    a2, b2 = a.clone(), b.clone()
    p.polymul_dev(a2.data_ptr(), a2.data_ptr(), b.data_ptr(), batch, st)      # c over a
    p.polymul_dev(b2.data_ptr(), a.data_ptr(), b2.data_ptr(), batch, st)      # c over b
    torch.cuda.synchronize()
    assert bool((a2 == c).all()) and bool((b2 == c).all())
    want = oracle.product(n, q, a[:3].cpu().numpy(), b[:3].cpu().numpy(), 10)
    assert (c[:3].cpu().numpy() == want).all()
    p.close()


def _dev_buffers(torch, *arrays):
    return [torch.from_numpy(x).cuda() for x in arrays]


@pytest.mark.parametrize("n,q,logb,cfg", [(256, 12289, 20, 2), (256, 7681, 20, 3), (1024, 12289, 18, 4),
                                           (256, 3329, 20, 3)])
def test_full_size_batches_properties(gpu, oracle_mt, nttb200, golden, n, q, logb, cfg):
    """BASELINE configs 3/4 at full batch, device-resident, on the SURVEY 8d batch (splitmix64
    stream, edge rows 0..7, the fixture pair in row 8): 2^13 sampled rows + every edge row + the
    last rows against the oracle, plus size-independent properties over the WHOLE batch
    (commutativity, linearity in a).  q = 3329 has no 512-th root of unity: that case is the
    cyclic product."""
    import torch
    batch = 1 << logb
    cyclic = (q - 1) % (2 * n) != 0
    variant = 30 if cyclic else 10
    p = gpu.Plan(n, q, cyclic=cyclic)
    a, b = nttb200.inputs.survey_batch(n, q, batch, cfg, device="cuda",
                                       fixture=(golden["fixture_a"], golden["fixture_b"]))
    c = torch.empty_like(a)
    st = torch.cuda.current_stream().cuda_stream
    p.polymul_dev(c.data_ptr(), a.data_ptr(), b.data_ptr(), batch, st)
    torch.cuda.synchronize()
    assert int(c.min()) >= 0 and int(c.max()) < q
    assert bool((c[0] == 0).all()) and bool((c[2] == b[2]).all())
    if not cyclic:
        assert int(c[3, n - 2]) == q - 1 and int(c[3].long().sum()) == q - 1
        assert c[4, :3].tolist() == [3, 6, 0] and c[5, :4].tolist() == [2, 4, 6, 0]
        assert c[6, :5].tolist() == [2, 6, 10, 6, 0] and c[7, :8].tolist() == [3, 9, 6, 1, 8, 6, 0, 2]
    if (n, q) == (256, 12289):
        assert (c[8].cpu().numpy() == golden["fixture_c"]).all()
    idx = np.unique(np.concatenate([np.arange(0, 16), np.random.default_rng(1).integers(0, batch, 1 << 13),
                                    np.arange(batch - 16, batch)]))
    assert idx.size >= 1 << 12
    ti = torch.from_numpy(idx).cuda()
    want = oracle_mt(n, q, a[ti].cpu().numpy(), b[ti].cpu().numpy(), variant)
    assert (c[ti].cpu().numpy() == want).all()
    # commutativity on the whole batch
    c2 = torch.empty_like(a)
    p.polymul_dev(c2.data_ptr(), b.data_ptr(), a.data_ptr(), batch, st)
    torch.cuda.synchronize()
    assert bool((c == c2).all())
    # linearity: (a + a') * b = a*b + a'*b  with a' = a rolled by one row
    a2 = torch.roll(a, 1, 0)
    s = ((a.long() + a2.long()) % q).int()
    cs = torch.empty_like(a)
    p.polymul_dev(cs.data_ptr(), s.data_ptr(), b.data_ptr(), batch, st)
    p.polymul_dev(c2.data_ptr(), a2.data_ptr(), b.data_ptr(), batch, st)
    torch.cuda.synchronize()
    assert bool((cs.long() == (c.long() + c2.long()) % q).all())
    p.close()


def test_kernels_cross_check_fused_vs_literal_dataflow(gpu, oracle):
    """The fused register/smem kernel against the one-stage-per-launch literal dataflow
    kernels (both on the GPU), n=1024."""
    n, q = 1024, 12289
    p = gpu.Plan(n, q)
    a = oracle.random((300, n), q, 77)
    fused = p.transform("mulntt_std2rev", a)
    literal = gpu.ntt_table_batch(n, q, "ct_std2rev", gpu.make_table("mixed_powers_rev", n, q, p.psi), a)
    assert (fused == literal).all()
    p.close()


def test_multi_gpu_sharder_in_one_process(gpu, oracle):
    """nttb200_multi_*: every visible GPU gets a contiguous slice (1 GPU on the default box; the
    same test is run with --gpus 2)."""
    n, q = 256, 12289
    m = gpu.MultiPlan(n, q, 1002)
    assert m.gpus == gpu.device_count() >= 1
    for batch in (1, 7, 4099):
        a, b = oracle.random((batch, n), q, 1 + batch), oracle.random((batch, n), q, 2 + batch)
        assert (m.polymul(a, b) == oracle.product(n, q, a, b, 10)).all()
    m.close()


@pytest.mark.parametrize("n,q", [(8, 17), (16, 97), (64, 257), (128, 3329), (256, 12289), (256, 7681), (512, 12289),
                                 (1024, 12289)])
def test_packed_u16_extension(gpu, oracle, n, q):
    """nttb200_polymul_batch_u16 (outside the reference API): same products, 16-bit transport."""
    p = gpu.Plan(n, q)
    for batch in (1, 3, 33, 1027):
        a, b = oracle.random((batch, n), q, SEED + 7 * n + batch), oracle.random((batch, n), q, SEED + 9 * q + batch)
        a[0], b[0] = q - 1, q - 1
        got = p.polymul_u16(a.astype(np.uint16), b.astype(np.uint16))
        assert got.dtype == np.uint16
        assert (got.astype(np.int32) == oracle.product(n, q, a, b, 10)).all(), (n, q, batch)
    p.close()
    big = gpu.Plan(256, 8380417)
    with pytest.raises(gpu.NttError):
        big.polymul_u16(np.zeros((1, 256), np.uint16), np.zeros((1, 256), np.uint16))
    big.close()


def test_operands_not_16_byte_aligned_fall_back_to_the_shoup_kernel(gpu, oracle):
    """The Plantard kernel prefetches with 16-byte cp.async; operands at a 4-byte offset must still
    give the right products (served by the Shoup kernel of the same plan)."""
    import torch
    n, q, batch = 256, 12289, 301
    p = gpu.Plan(n, q)
    a = oracle.random((batch, n), q, 5)
    b = oracle.random((batch, n), q, 6)
    da = torch.zeros(batch * n + 1, dtype=torch.int32, device="cuda")
    db = torch.zeros(batch * n + 3, dtype=torch.int32, device="cuda")
    dc = torch.zeros(batch * n + 1, dtype=torch.int32, device="cuda")
    da[1:] = torch.from_numpy(a.reshape(-1)).cuda()
    db[3:] = torch.from_numpy(b.reshape(-1)).cuda()
    st = torch.cuda.current_stream().cuda_stream
    p.polymul_dev(dc.data_ptr() + 4, da.data_ptr() + 4, db.data_ptr() + 12, batch, st)
    torch.cuda.synchronize()
    assert (dc[1:].cpu().numpy().reshape(batch, n) == oracle.product(n, q, a, b, 10)).all()
    assert int(dc[0]) == 0
    p.close()


def test_large_n_calls_on_two_streams_share_the_scratch_safely(gpu, oracle):
    """Two device-resident calls on ONE large-n plan, issued back to back on two different
    streams, use the same scratch: the second must wait for the first on the device."""
    import torch
    n, q, batch = 8192, 65537, 96
    p = gpu.Plan(n, q)
    g = torch.Generator(device="cuda").manual_seed(3)
    xs = [torch.randint(0, q, (batch, n), dtype=torch.int32, device="cuda", generator=g) for _ in range(4)]
    outs = [torch.empty_like(xs[0]) for _ in range(2)]
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    for _ in range(3):
        p.polymul_dev(outs[0].data_ptr(), xs[0].data_ptr(), xs[1].data_ptr(), batch, s1.cuda_stream)
        p.polymul_dev(outs[1].data_ptr(), xs[2].data_ptr(), xs[3].data_ptr(), batch, s2.cuda_stream)
    torch.cuda.synchronize()
    idx = [0, 47, 95]
    for o, (u, v) in zip(outs, ((xs[0], xs[1]), (xs[2], xs[3]))):
        want = oracle.product(n, q, u[idx].cpu().numpy(), v[idx].cpu().numpy(), 10)
        assert (o[idx].cpu().numpy() == want).all()
    p.close()


@pytest.mark.parametrize("n,q,batch", [(256, 12289, 2048), (256, 12289, 20011), (256, 7681, 9000),
                                       (1024, 12289, 3001), (64, 257, 40000)])
def test_host_buffer_wire_formats_give_the_same_products(gpu, oracle, n, q, batch, monkeypatch):
    """nttb200_polymul_batch on half-word moduli: chunks cross PCIe as 16-bit or 32-bit words
    (csrc/hostwire.c, polymul_batch_wire).  Every mode -- forced 16, forced 32, automatic mix,
    pageable and pinned buffers -- must return the reference products, and the statistics must
    show which wire was used."""
    p = gpu.Plan(n, q)
    a, b = oracle.random((batch, n), q, SEED + batch), oracle.random((batch, n), q, SEED + 3 * batch)
    a[0], b[0], a[-1], b[-1] = q - 1, q - 1, q - 1, 1
    idx = np.unique(np.r_[0:40, batch - 40:batch, np.random.default_rng(batch).integers(0, batch, 300)])
    want = oracle.product(n, q, a[idx], b[idx], 10)
    ref = None
    for mode in ("32", "16", "auto"):
        monkeypatch.setenv("NTTB200_WIRE", mode)
        got = p.polymul(a, b)                               # pageable numpy buffers
        st = p.wire_stats()
        assert (got[idx] == want).all(), mode
        if mode == "32":
            assert st["rows16"] == 0
            ref = got
        else:
            assert st["rows16"] == batch and st["rows32"] == 0, st      # pageable: always narrowed
            assert st["result_rows16"] == batch, st                     # ... and widened by the pool
            assert (got == ref).all(), mode
    ha, hb, hc = (gpu.host_alloc((batch, n)) for _ in range(3))
    ha.array[:], hb.array[:] = a, b
    for mode in ("16", "auto"):
        monkeypatch.setenv("NTTB200_WIRE", mode)
        hc.array[:] = -1
        p.polymul_host_ptr(hc.ptr, ha.ptr, hb.ptr, batch)
        st = p.wire_stats()
        assert (hc.array == ref).all(), mode
        assert st["result_rows16"] == st["rows16"], st         # widened by the pool ...
        monkeypatch.setenv("NTTB200_WIRE_C32", "1")            # ... or int32 rows written by the kernel
        monkeypatch.setenv("NTTB200_WIRE_AHEAD", "1")          # and chunks as 32-bit words when the pool lags
        hc.array[:] = -1
        p.polymul_host_ptr(hc.ptr, ha.ptr, hb.ptr, batch)
        st = p.wire_stats()
        assert (hc.array == ref).all() and st["result_rows16"] == 0 and st["rows16"] + st["rows32"] == batch, (mode, st)
        monkeypatch.delenv("NTTB200_WIRE_C32")
        monkeypatch.delenv("NTTB200_WIRE_AHEAD")
        assert st["rows16"] + st["rows32"] == batch and (mode == "auto" or st["rows32"] == 0), st
    for h in (ha, hb, hc):
        h.free()
    p.close()


def test_words_that_do_not_fit_16_bits_travel_as_32_bit_words(gpu, oracle, monkeypatch):
    """Out-of-contract inputs (a word >= 2^16, a negative word) must give exactly what the 32-bit
    wire gives: the chunk that holds them is re-sent as 32-bit words."""
    n, q, batch = 256, 12289, 12000
    p = gpu.Plan(n, q)
    a, b = oracle.random((batch, n), q, 11), oracle.random((batch, n), q, 12)
    a[5000, 17] += 8 * q                                    # 8q + x >= 2^16: a lazy representative
    b[11999, 255] = -1
    monkeypatch.setenv("NTTB200_WIRE", "32")
    ref = p.polymul(a, b)
    for mode in ("16", "auto"):
        monkeypatch.setenv("NTTB200_WIRE", mode)
        got = p.polymul(a, b)
        st = p.wire_stats()
        assert (got == ref).all()
        assert st["rows32"] >= 2 and st["rows16"] >= 1 and st["rows16"] + st["rows32"] == batch, st
    rows = np.r_[0:8, 4990:5000, 5001:5010]
    assert (ref[rows] == oracle.product(n, q, a[rows], b[rows], 10)).all()
    p.close()


@pytest.mark.parametrize("n,q", [(2048, 12289), (8192, 998244353), (2048, 2013265921)])
def test_large_n_operands_at_a_4_byte_offset(gpu, oracle, n, q):
    """The column passes of the cheap-butterfly classes read two adjacent columns per lane with
    64-bit accesses; rows that are only 4-byte aligned must take the one-column kernels and
    give the same products."""
    import torch
    batch = 7
    p = gpu.Plan(n, q)
    a, b = oracle.random((batch, n), q, 21), oracle.random((batch, n), q, 22)
    da = torch.zeros(batch * n + 1, dtype=torch.int32, device="cuda")
    db = torch.zeros(batch * n + 1, dtype=torch.int32, device="cuda")
    dc = torch.zeros(batch * n + 1, dtype=torch.int32, device="cuda")
    da[1:] = torch.from_numpy(a.reshape(-1)).cuda()
    db[1:] = torch.from_numpy(b.reshape(-1)).cuda()
    st = torch.cuda.current_stream().cuda_stream
    p.polymul_dev(dc.data_ptr() + 4, da.data_ptr() + 4, db.data_ptr() + 4, batch, st)
    torch.cuda.synchronize()
    want = oracle.product(n, q, a, b, 10)
    assert (dc[1:].cpu().numpy().reshape(batch, n) == want).all()
    assert int(dc[0]) == 0
    with pytest.raises(gpu.NttError):                       # transforms: 16-byte aligned rows only
        p.transform_dev("mulntt_std2rev", da.data_ptr() + 4, batch, st)
    x = da[1:].clone()
    assert x.data_ptr() % 16 == 0
    p.transform_dev("mulntt_std2rev", x.data_ptr(), batch, st)
    p.transform_dev("inttmul_rev2std_scaled", x.data_ptr(), batch, st)
    torch.cuda.synchronize()
    assert (x.cpu().numpy().reshape(batch, n) == a).all()
    p.close()


@pytest.mark.parametrize("n,q,batch", [(256, 8380417, 9001), (1024, 2013265921, 1500), (256, 12289, 5000)])
def test_pageable_buffers_are_staged_by_the_host_pool(gpu, oracle, n, q, batch, monkeypatch):
    """malloc'd (pageable) caller buffers -- the reference's convention -- of plans that cannot use
    the 16-bit wire: the host pool copies 32-bit words through pinned staging instead of the
    driver's pageable path.  Same products as the plain DMA ring."""
    p = gpu.Plan(n, q)
    a, b = oracle.random((batch, n), q, 31 + batch), oracle.random((batch, n), q, 32 + batch)
    idx = np.unique(np.r_[0:16, batch - 16:batch, np.random.default_rng(n).integers(0, batch, 64)])
    want = oracle.product(n, q, a[idx], b[idx], 10)
    monkeypatch.setenv("NTTB200_WIRE", "32")
    monkeypatch.setenv("NTTB200_STAGE_PAGEABLE", "0")
    ref = p.polymul(a, b)                                   # driver-staged ring
    assert p.wire_stats()["rows32"] == 0 and (ref[idx] == want).all()
    monkeypatch.setenv("NTTB200_STAGE_PAGEABLE", "1")
    got = p.polymul(a, b)
    st = p.wire_stats()
    assert st["rows32"] == batch and st["rows16"] == 0 and st["host_threads"] >= 1, st
    assert (got == ref).all()
    monkeypatch.delenv("NTTB200_WIRE")
    got = p.polymul(a, b)                                   # default mode
    assert (got == ref).all()
    ha, hb, hc = (gpu.host_alloc((batch, n)) for _ in range(3))     # mixed: pinned operands, pageable result
    ha.array[:], hb.array[:] = a, b
    out = np.full((batch, n), -1, np.int32)
    p.polymul_host_ptr(out.ctypes.data, ha.ptr, hb.ptr, batch)
    assert (out == ref).all()
    for h in (ha, hb, hc):
        h.free()
    p.close()


@pytest.mark.parametrize("n,q,logb", [(256, 7681, 19), (1024, 12289, 16)])
def test_host_buffer_call_equals_device_resident_call_on_a_large_batch(gpu, oracle, n, q, logb):
    """Hundreds of chunks through the wire pipeline (ring reuse, ramp, taper): the int32 rows that
    come back into the caller's buffer must equal, row for row, what the device-resident call
    gives on the same operands, and sampled rows must equal the oracle."""
    import torch
    batch = (1 << logb) + 77                                # not a multiple of anything
    p = gpu.Plan(n, q)
    g = torch.Generator(device="cuda").manual_seed(99 + n)
    da = torch.randint(0, q, (batch, n), dtype=torch.int32, device="cuda", generator=g)
    db = torch.randint(0, q, (batch, n), dtype=torch.int32, device="cuda", generator=g)
    dc = torch.empty_like(da)
    p.polymul_dev(dc.data_ptr(), da.data_ptr(), db.data_ptr(), batch, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    a, b = da.cpu().numpy(), db.cpu().numpy()               # pageable host copies
    got = p.polymul(a, b)
    st = p.wire_stats()
    assert st["rows16"] + st["rows32"] == batch, st
    assert (got == dc.cpu().numpy()).all()
    idx = np.unique(np.r_[0:4, batch - 4:batch, np.random.default_rng(3).integers(0, batch, 120)])
    assert (got[idx] == oracle.product(n, q, a[idx], b[idx], 10)).all()
    p.close()


@pytest.mark.parametrize("n,q,batch", [(256, 12289, 5003), (1024, 8380417, 700), (4096, 2013265921, 301)])
def test_batch_transforms_from_pageable_memory_are_staged_by_the_pool(gpu, oracle, loader, n, q, batch, monkeypatch):
    """nttb200_ntt_batch on a large pageable array: the host pool stages the rows through pinned
    memory; same output as the plain DMA ring, forward then scaled inverse is the identity, and
    sampled rows of the forward transform equal the oracle's."""
    p = gpu.Plan(n, q)
    a = oracle.random((batch, n), q, 77 + batch)
    monkeypatch.setenv("NTTB200_STAGE_PAGEABLE", "0")
    ref = p.transform("mulntt_std2rev", a)
    monkeypatch.setenv("NTTB200_STAGE_PAGEABLE", "1")
    got = p.transform("mulntt_std2rev", a)
    assert (got == ref).all()
    back = p.transform("inttmul_rev2std_scaled", got)
    assert (back == a).all()
    idx = np.unique(np.r_[0:3, batch - 3:batch, np.random.default_rng(5).integers(0, batch, 20)])
    want = oracle.transform("mulntt_ct_std2rev", a[idx], oracle.table(loader.MIXED_POWERS_REV, n, q, p.psi), q)
    assert (got[idx] == want).all()
    p.close()


def test_independent_launches_skip_the_wait_dependent_ones_do_not(gpu, oracle, nttb200):
    """Launches on one stream whose operands share nothing with the launches in flight start before
    those finish (programmatic dependent launch without the initial wait); a launch that reads the
    previous result, or writes over anything in flight, waits.  Chains of both kinds, interleaved
    with a foreign kernel, give the oracle's values."""
    import torch
    n, q, batch = 256, 12289, 1 << 14
    p = gpu.Plan(n, q)
    st = torch.cuda.current_stream().cuda_stream
    sets = []
    for k in range(4):
        a, b = nttb200.inputs.survey_batch(n, q, batch, 2, device="cuda", row_offset=k * batch)
        sets.append((a, b, torch.empty_like(a)))
    for rep in range(3):                               # independent: rotating buffer sets
        for a, b, c in sets:
            p.polymul_dev(c.data_ptr(), a.data_ptr(), b.data_ptr(), batch, st)
    torch.cuda.synchronize()
    rows = np.r_[0:9, batch - 3:batch, np.random.default_rng(3).integers(0, batch, 200)]
    for a, b, c in sets:
        want = oracle.product(n, q, a[rows].cpu().numpy(), b[rows].cpu().numpy(), 10)
        assert (c[rows].cpu().numpy() == want).all()
    # dependent chain: c1 = a*b, c2 = c1*b, c3 = c2*c1 (reads two earlier results), then overwrite a
    a, b, c1 = sets[0]
    c2, c3 = sets[1][2], sets[2][2]
    a0 = a.clone()
    p.polymul_dev(c1.data_ptr(), a.data_ptr(), b.data_ptr(), batch, st)
    p.polymul_dev(c2.data_ptr(), c1.data_ptr(), b.data_ptr(), batch, st)      # RAW on c1
    c1.add_(0)                                                                 # a foreign kernel in between
    p.polymul_dev(c3.data_ptr(), c2.data_ptr(), c1.data_ptr(), batch, st)      # RAW on c2 and c1
    p.polymul_dev(a.data_ptr(), c3.data_ptr(), b.data_ptr(), batch, st)        # WAR on a (read by the first)
    torch.cuda.synchronize()
    r = rows[:64]
    w1 = oracle.product(n, q, a0[r].cpu().numpy(), b[r].cpu().numpy(), 10)
    w2 = oracle.product(n, q, w1, b[r].cpu().numpy(), 10)
    w3 = oracle.product(n, q, w2, w1, 10)
    w4 = oracle.product(n, q, w3, b[r].cpu().numpy(), 10)
    assert (c1[r].cpu().numpy() == w1).all() and (c2[r].cpu().numpy() == w2).all()
    assert (c3[r].cpu().numpy() == w3).all() and (a[r].cpu().numpy() == w4).all()
    # an event recorded after an independent launch fires only when every earlier launch is done too
    big = [nttb200.inputs.survey_batch(n, q, 1 << 18, 3, device="cuda") for _ in range(1)]
    ab, bb = big[0]
    cb = torch.empty_like(ab)
    small_a, small_b, small_c = sets[3]
    p.polymul_dev(cb.data_ptr(), ab.data_ptr(), bb.data_ptr(), 1 << 18, st)     # long
    small_c.zero_()
    p.polymul_dev(small_c.data_ptr(), small_a.data_ptr(), small_b.data_ptr(), 64, st)   # short, independent
    ev = torch.cuda.Event()
    ev.record()
    ev.synchronize()
    rr = np.r_[(1 << 18) - 64:(1 << 18)]
    want = oracle.product(n, q, ab[rr].cpu().numpy(), bb[rr].cpu().numpy(), 10)
    assert (cb[rr].cpu().numpy() == want).all()
    p.close()


@pytest.mark.parametrize("n,q", [(256, 12289), (256, 8380417), (1024, 12289)])
def test_random_launch_sequences_on_one_stream_equal_their_sequential_meaning(gpu, oracle, nttb200, n, q):
    """60 products issued back to back on ONE stream over a pool of six buffers, each on a random row
    range, with operands and results drawn at random (results over operands, results read by the next
    launch, ranges that overlap partly): whatever the library decides about which launches may start
    early, every buffer ends up as if the launches had run one after the other (replayed with the
    oracle on the sampled rows -- rows are independent)."""
    import torch
    rows = 1 << 14
    p = gpu.Plan(n, q)
    st = torch.cuda.current_stream().cuda_stream
    rng = np.random.default_rng(n + q)
    bufs = [nttb200.inputs.survey_batch(n, q, rows, 2, device="cuda", row_offset=k * rows)[0] for k in range(6)]
    watch = np.unique(np.r_[0:4, rows - 4:rows, rng.integers(0, rows, 40)])
    model = [b[torch.from_numpy(watch).cuda()].cpu().numpy() for b in bufs]
    plan = []
    for _ in range(60):
        ia, ib, ic = rng.integers(0, 6, 3)
        lo = int(rng.integers(0, rows - 1))
        hi = int(rng.integers(lo + 1, min(rows, lo + rng.choice([64, 2048, rows])) + 1))
        plan.append((int(ia), int(ib), int(ic), lo, hi))
    for ia, ib, ic, lo, hi in plan:
        off = lo * n * 4
        p.polymul_dev(bufs[ic].data_ptr() + off, bufs[ia].data_ptr() + off, bufs[ib].data_ptr() + off, hi - lo, st)
    torch.cuda.synchronize()
    for ia, ib, ic, lo, hi in plan:
        m = (watch >= lo) & (watch < hi)
        if m.any():
            model[ic][m] = oracle.product(n, q, model[ia][m], model[ib][m], 10)
    for k in range(6):
        got = bufs[k][torch.from_numpy(watch).cuda()].cpu().numpy()
        assert (got == model[k]).all(), k
    p.close()


@pytest.mark.parametrize("n,q", [(256, 12289), (256, 8380417), (1024, 12289)])
def test_launch_after_a_waiting_launch_does_not_overtake_older_launches(gpu, oracle, nttb200, n, q):
    """L1 writes X (long), L2 is tiny, reads one row of X (so it waits for L1) and writes elsewhere, L3
    reads a window of X and shares nothing with L2.  The library checks L3 only against the launches since
    the last one that waited (L2), so L2 must not let its dependents go before its own wait is over: L3
    would run next to L1 and read rows of X that L1 has not written yet.  (With the trigger ahead of the
    wait this fails on the windows that L1 writes last.)"""
    import torch
    rows = (1 << 16) if n == 256 else (1 << 14)
    p = gpu.Plan(n, q)
    st = torch.cuda.current_stream().cuda_stream
    a, b = nttb200.inputs.survey_batch(n, q, rows, 2, device="cuda")
    windows = [(rows * 3 // 4, rows), (rows * 27 // 32, rows * 28 // 32), (rows - 2048, rows), (rows // 2, rows * 5 // 8),
               (rows * 13 // 16, rows * 14 // 16), (rows - 256, rows), (0, rows)]
    for rep, (lo, hi) in enumerate(windows * 2):
        watch = np.unique(np.r_[lo:lo + 8, hi - 64:hi, np.random.default_rng(rep).integers(lo, hi, 64)])
        tw = torch.from_numpy(watch).cuda()
        hb = b[tw].cpu().numpy()
        hx = oracle.product(n, q, a[tw].cpu().numpy(), hb, 10)
        want = oracle.product(n, q, hx, hb, 10)
        x = torch.full_like(a, q - 1)                # what a too-early L3 would read
        z = torch.empty((1, n), dtype=a.dtype, device="cuda")
        d = torch.zeros_like(a)
        torch.cuda.synchronize()
        off = lo * n * 4
        p.polymul_dev(x.data_ptr(), a.data_ptr(), b.data_ptr(), rows, st)                               # L1
        p.polymul_dev(z.data_ptr(), x.data_ptr(), b.data_ptr(), 1, st)                                  # L2: waits for L1
        p.polymul_dev(d.data_ptr() + off, x.data_ptr() + off, b.data_ptr() + off, hi - lo, st)          # L3
        torch.cuda.synchronize()
        assert (x[tw].cpu().numpy() == hx).all()
        assert (d[tw].cpu().numpy() == want).all(), (rep, lo, hi)
    p.close()


def _friendly_primes(L, n, bits, count):
    out = []
    for seed in range(count * 3):
        q = L.nttb200_gen_prime(bits, n, seed)
        if q and q not in out:
            out.append(int(q))
        if len(out) == count:
            break
    return out


@pytest.mark.parametrize("logn", [3, 5, 8, 9, 10, 11, 13, 16, 17])
def test_products_at_generated_primes_of_every_width(gpu, oracle, logn):
    """NTT-friendly primes from the library's own generator (Generator_Params twin) at 13 .. 31 bits:
    every arithmetic class (Plantard, LAZY, HARVEY, CANON) at every kernel family, against the oracle."""
    n = 1 << logn
    import ctypes
    L = gpu.lib()
    L.nttb200_gen_prime.restype = ctypes.c_uint32
    L.nttb200_gen_prime.argtypes = [ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint64]
    seen = set()
    for bits in (13, 14, 17, 20, 23, 26, 29, 30, 31):
        if bits < logn + 2:
            continue
        for q in _friendly_primes(L, n, bits, 2 if logn <= 10 else 1):
            if q in seen or (q - 1) % (2 * n):
                continue
            seen.add(q)
            p = gpu.Plan(n, q)
            batch = 9 if logn <= 10 else 3
            a, b = oracle.random((batch, n), q, SEED + q), oracle.random((batch, n), q, SEED + q + 1)
            a[0], b[0] = q - 1, q - 1
            a[1] = np.where(np.arange(n) % 2 == 0, q - 1, 0)
            assert (p.polymul(a, b) == oracle.product(n, q, a, b, 10)).all(), p.describe()
            p.close()
    assert len(seen) >= 4


def test_small_product_right_after_a_large_n_product_on_the_same_stream(gpu, oracle, nttb200):
    """A large-n product (its last kernel triggers its dependents early) followed at once by an
    n = 256 product that reads its result viewed as rows of 256: the second one must wait."""
    import torch
    nl, q = 2048, 12289                                  # 2 nl = 4096 divides q - 1 = 2^12 3
    pl, ps = gpu.Plan(nl, q), gpu.Plan(256, q)
    st = torch.cuda.current_stream().cuda_stream
    a, b = nttb200.inputs.survey_batch(nl, q, 48, 4, device="cuda")
    warm = nttb200.inputs.survey_batch(256, q, 64, 2, device="cuda")
    wc = torch.empty_like(warm[0])
    for rep in range(3):
        c = torch.zeros_like(a)
        d = torch.empty_like(a)
        ps.polymul_dev(wc.data_ptr(), warm[0].data_ptr(), warm[1].data_ptr(), 64, st)     # history on the stream
        pl.polymul_dev(c.data_ptr(), a.data_ptr(), b.data_ptr(), 48, st)
        ps.polymul_dev(d.data_ptr(), c.data_ptr(), b.data_ptr(), 48 * nl // 256, st)      # reads c as rows of 256
        torch.cuda.synchronize()
        hc = oracle.product(nl, q, a.cpu().numpy(), b.cpu().numpy(), 10)
        assert (c.cpu().numpy() == hc).all()
        want = oracle.product(256, q, hc.reshape(-1, 256), b.cpu().numpy().reshape(-1, 256), 10)
        assert (d.cpu().numpy().reshape(-1, 256) == want).all()
    pl.close()
    ps.close()


@pytest.mark.parametrize("n,q", [(256, 12289), (1024, 12289), (256, 8380417)])
def test_asynchronous_host_buffer_products_equal_the_synchronous_call(gpu, oracle, nttb200, n, q):
    """nttb200_polymul_batch_async: a queue of host-buffer products of mixed sizes (wire-eligible and
    not, pinned and pageable, an empty one) run by the plan's worker as ONE stream of jobs through the
    pipeline of the synchronous call; every result equals the synchronous call's and the oracle's on
    sampled rows, tickets can be waited for in any order, and a synchronous call may come in between."""
    p = gpu.Plan(n, q)
    sizes = [(1 << 19) // n * 4, 3, (1 << 19) // n * 2 + 17, 0, (1 << 19) // n * 8, 70, (1 << 19) // n * 3]
    jobs = []
    for k, rows in enumerate(sizes):
        a, b = oracle.random((max(rows, 1), n), q, 100 + k)[:rows], oracle.random((max(rows, 1), n), q, 200 + k)[:rows]
        if k % 2 == 0 and rows:                      # pinned buffers for every other job
            ha, hb, hc = nttb200.host_alloc((rows, n)), nttb200.host_alloc((rows, n)), nttb200.host_alloc((rows, n))
            ha.array[:], hb.array[:] = a, b
            jobs.append((ha.array, hb.array, hc.array, (ha, hb, hc)))
        else:
            jobs.append((np.ascontiguousarray(a), np.ascontiguousarray(b), np.empty((rows, n), np.int32), None))
    for rep in range(2):
        for _, _, c, _ in jobs:
            c[...] = -1
        tickets = [p.polymul_async_ptr(c.ctypes.data, a.ctypes.data, b.ctypes.data, a.shape[0]) for a, b, c, _ in jobs]
        if rep == 1:
            mid = p.polymul(jobs[1][0], jobs[1][1])   # a synchronous call queues behind the stream
            assert (mid == oracle.product(n, q, jobs[1][0], jobs[1][1], 10)).all()
        for t in (reversed(tickets) if rep == 0 else tickets[::2]):
            p.wait(t)
        p.wait(0)
        for a, b, c, _ in jobs:
            rows = a.shape[0]
            if rows == 0:
                continue
            pick = np.unique(np.r_[0:min(rows, 3), max(0, rows - 3):rows, np.random.default_rng(rows).integers(0, rows, 16)])
            assert (c[pick] == oracle.product(n, q, a[pick], b[pick], 10)).all()
            assert c.min() >= 0 and c.max() < q
    with pytest.raises(nttb200.NttError):
        p.wait(tickets[0])                           # already waited for
    for *_, keep in jobs:
        if keep:
            for h in keep:
                h.free()
    p.close()


def test_asynchronous_products_from_several_threads(gpu, oracle):
    """Four threads queue and wait for host-buffer products on ONE plan at the same time (submission is
    serialised inside the library, the worker runs one stream of jobs), with synchronous calls thrown in:
    every thread gets its own results back."""
    import threading
    n, q = 256, 12289
    p = gpu.Plan(n, q)
    errors = []

    def work(tid):
        try:
            rng = np.random.default_rng(1000 + tid)
            for it in range(6):
                rows = int(rng.choice([5, 300, 2048, 4096, 9000]))
                a, b = oracle.random((rows, n), q, 7 * tid + it), oracle.random((rows, n), q, 900 + 7 * tid + it)
                c = np.full((rows, n), -1, np.int32)
                if it % 3 == 2:
                    c = p.polymul(a, b)
                else:
                    t = p.polymul_async_ptr(c.ctypes.data, a.ctypes.data, b.ctypes.data, rows)
                    p.wait(t)
                pick = np.unique(np.r_[0, rows - 1, rng.integers(0, rows, 12)])
                if not (c[pick] == oracle.product(n, q, a[pick], b[pick], 10)).all():
                    errors.append((tid, it, rows))
        except Exception as ex:                     # noqa: BLE001 -- reported by the main thread
            errors.append((tid, repr(ex)))

    threads = [threading.Thread(target=work, args=(t,)) for t in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    assert not errors, errors
    p.wait(0)
    p.close()
