"""CPU: the oracle restatement against the committed golden vectors (generated from the
compiled, unmodified reference by tests/golden/make_golden.py) and against the reference's
own committed vectors (hardware testbench, q = 7681)."""
import numpy as np
import pytest

N, Q, PSI = 256, 12289, 1002

TRANSFORM_IDS = {  # ref_shim.c ids -> (oracle function, oracle table kind)
    0: ("ntt_ct_rev2std", 3), 1: ("ntt_gs_rev2std", 4), 2: ("ntt_ct_std2rev", 4), 3: ("ntt_gs_std2rev", 3),
    4: ("ntt_ct_rev2std", 5), 5: ("ntt_gs_rev2std", 6), 6: ("ntt_ct_std2rev", 6), 7: ("ntt_gs_std2rev", 5),
    8: ("mulntt_ct_rev2std", 7), 9: ("mulntt_ct_std2rev", 8), 10: ("nttmul_gs_rev2std", 10),
    11: ("nttmul_gs_std2rev", 9), 12: ("ntt_ct_rev2std_v1", 0),
}
RED_IDS = {
    100: ("ct_rev2std", 3), 101: ("gs_rev2std", 4), 102: ("ct_std2rev", 4), 103: ("gs_std2rev", 3),
    104: ("ct_rev2std", 5), 105: ("gs_rev2std", 6), 106: ("ct_std2rev", 6), 107: ("gs_std2rev", 5),
    108: ("mulntt_ct_rev2std", 7), 109: ("mulntt_ct_std2rev", 8), 110: ("nttmul_gs_rev2std", 10),
    111: ("nttmul_gs_std2rev", 9),
}


def test_params(golden, oracle):
    psi, omega, ipsi, iomega, ninv, invk, r8, r6 = golden["params"].tolist()
    assert (psi, omega, ipsi, iomega, ninv) == (1002, 8595, 10805, 2525, 12241)
    assert omega == psi * psi % Q and psi * ipsi % Q == 1 and omega * iomega % Q == 1
    assert ninv * N % Q == 1 and invk * 3 % Q == 1
    assert r8 == ninv * pow(invk, 8, Q) % Q and r6 == ninv * pow(invk, 6, Q) % Q
    assert oracle.psi(N, Q, 0) == 3            # the generator's rule picks the smallest root


@pytest.mark.parametrize("kind", range(11))
def test_tables(golden, oracle, kind):
    assert (oracle.table(kind, N, Q, PSI).astype(np.int64) == golden[f"table_{kind}"]).all()
    assert (oracle.red_table(kind, N, PSI).astype(np.int64) == golden[f"red_table_{kind}"]).all()


def test_red_var_table(golden, oracle):
    assert (oracle.red_table(100, N, PSI).astype(np.int64) == golden["red_table_100"]).all()


@pytest.mark.parametrize("variant", [1, 4, 10, 20])
def test_products_fixture_and_kats(golden, oracle, variant):
    c = oracle.product(N, Q, golden["fixture_a"], golden["fixture_b"], variant, PSI)
    assert (c[0] == golden["fixture_c"]).all()
    assert golden["fixture_c"][:8].tolist() == [2562, 4542, 3303, 357, 2079, 10183, 1177, 7019]
    assert golden["fixture_c"][-4:].tolist() == [12120, 1856, 7067, 7618]
    assert (oracle.product(N, Q, golden["kat_a"], golden["kat_b"], variant, PSI) == golden["kat_c"]).all()


@pytest.mark.parametrize("variant", [1, 4, 10, 20])
def test_products_random(golden, oracle, variant):
    a, b = golden["rand_a"], golden["rand_b"]
    if variant == 20:
        a, b = a[:24], b[:24]
    assert (oracle.product(N, Q, a, b, variant, PSI) == golden["rand_c"][: a.shape[0]]).all()


def test_products_do_not_depend_on_psi(golden, oracle):
    a, b = golden["rand_a"][:16], golden["rand_b"][:16]
    assert (oracle.product(N, Q, a, b, 10, 0) == golden["rand_c"][:16]).all()


@pytest.mark.parametrize("variant", [1, 4])
def test_red_products(golden, oracle, variant):
    a, b = golden["rand_a"][:32], golden["rand_b"][:32]
    assert (oracle.red_product(N, PSI, a, b, variant) == golden["rand_c"][:32]).all()


def test_product_clobbers_like_reference(golden, oracle, loader):
    """ntt256.h:80 'arrays a and b are modified': post-state = psi-twisted NTT, rev order."""
    a, b = golden["rand_a"][4].copy(), golden["rand_b"][4].copy()
    c = np.zeros(N, dtype=np.int32)
    assert oracle.lib.orc_product(oracle.plan(N, Q, PSI), 1, c, a, b) == 0
    assert (a == golden["clobber_a_after_product1"]).all()
    assert (b == golden["clobber_b_after_product1"]).all()
    mixed = oracle.transform("mulntt_ct_std2rev", golden["rand_a"][4], oracle.table(8, N, Q, PSI), Q)
    assert (mixed == a).all()


@pytest.mark.parametrize("tid", sorted(TRANSFORM_IDS))
def test_transforms(golden, oracle, tid):
    name, kind = TRANSFORM_IDS[tid]
    got = oracle.transform(name, golden["rand_a"][:16], oracle.table(kind, N, Q, PSI), Q)
    assert (got == golden[f"transform_{tid}"]).all()


@pytest.mark.parametrize("tid", sorted(RED_IDS))
def test_red_transforms(golden, oracle, tid):
    name, kind = RED_IDS[tid]
    got = oracle.red_transform(name, golden["red_transform_in"], oracle.red_table(kind, N, PSI))
    assert (got == golden[f"transform_{tid}"]).all()


# ids of ref_shim.c:ref_transform_tab -> oracle function (same names as the reference)
TAB_IDS = {0: "ntt_ct_rev2std_v1", 1: "ntt_ct_rev2std", 2: "mulntt_ct_rev2std", 3: "ntt_ct_std2rev",
           4: "mulntt_ct_std2rev", 5: "ntt_gs_rev2std", 6: "nttmul_gs_rev2std", 7: "ntt_gs_std2rev",
           8: "nttmul_gs_std2rev"}
RED_TAB_IDS = {1: "ct_rev2std", 2: "mulntt_ct_rev2std", 3: "ct_std2rev", 4: "mulntt_ct_std2rev",
               5: "gs_rev2std", 6: "nttmul_gs_rev2std", 7: "gs_std2rev", 8: "nttmul_gs_std2rev"}


@pytest.mark.parametrize("nn", [256, 64])
@pytest.mark.parametrize("tid", sorted(TAB_IDS))
def test_transforms_with_an_arbitrary_caller_table(golden, oracle, nn, tid):
    """The un-merged entry points never read p[t] (j = 0 peel, ntt.C:313-317 ...): with a table
    whose p[t] != 1 they differ from the merged ones, and the oracle follows the reference."""
    tab = golden[f"arb_table_{nn}"].astype(np.uint32)
    assert (tab[[1, 2, 4, 8]] != 1).all()
    got = oracle.transform(TAB_IDS[tid], golden[f"arb_in_{nn}"], tab, Q)
    assert (got == golden[f"arb_transform_{nn}_{tid}"]).all()
    if tid in (1, 3, 5, 7):      # and the peel matters: the merged twin gives something else
        assert (golden[f"arb_transform_{nn}_{tid}"] != golden[f"arb_transform_{nn}_{tid + 1}"]).any()


@pytest.mark.parametrize("nn", [256, 64])
@pytest.mark.parametrize("tid", sorted(RED_TAB_IDS))
def test_red_transforms_with_an_arbitrary_caller_table(golden, oracle, nn, tid):
    got = oracle.red_transform(RED_TAB_IDS[tid], golden[f"arb_red_in_{nn}"], golden[f"arb_red_table_{nn}"])
    assert (got == golden[f"arb_red_transform_{nn}_{tid}"]).all()


def test_ct_and_gs_agree_and_invert(golden, oracle):
    a = golden["rand_a"][:8]
    f1 = golden["transform_2"][:8]      # ntt256_ct_std2rev
    f2 = golden["transform_3"][:8]      # ntt256_gs_std2rev
    assert (f1 == f2).all()
    back = oracle.transform("ntt_gs_rev2std", f1, oracle.table(6, N, Q, PSI), Q)
    assert (back == (a.astype(np.int64) * N % Q)).all()      # unscaled: intt(ntt(a)) = n*a


def test_hw_vectors_q7681(hw_golden, oracle):
    """HW/simulation/modelsim/test: NTT_DOUT = gs_std2rev(NTT_DIN) with omega = 0xf04, and
    INTT_DOUT = n^-1 * gs_std2rev(INTT_DIN, omega^-1)  (SURVEY.md facts)."""
    n, q, w, w_inv, psi, psi_inv, ninv_r, r = hw_golden["PARAM"][:8].tolist()
    assert (n, q, psi) == (256, 7681, 62) and w == psi * psi % q and w * w_inv % q == 1
    din = hw_golden["NTT_DIN"].astype(np.int32)
    tab = oracle.omega_table(3, n, q, w)                    # omega_powers
    out = oracle.transform("ntt_gs_std2rev", din, tab, q)
    assert (out.astype(np.int64) == hw_golden["NTT_DOUT"]).all()
    same = oracle.transform("ntt_ct_std2rev", din, oracle.omega_table(4, n, q, w), q)
    assert (same == out).all()
    idin = hw_golden["INTT_DIN"].astype(np.int32)
    itab = oracle.omega_table(5, n, q, w)                   # inv_omega_powers
    iout = oracle.transform("ntt_gs_std2rev", idin, itab, q).astype(np.int64)
    ninv = pow(n, q - 2, q)
    assert (iout * ninv % q == hw_golden["INTT_DOUT"]).all()
    assert ninv_r == ninv * r % q


@pytest.mark.parametrize("n,q", [(8, 17), (16, 97), (64, 257), (128, 3329), (256, 7681), (512, 12289),
                                 (1024, 12289), (2048, 12289), (1024, 2013265921), (4096, 469762049)])
def test_variants_agree_with_schoolbook(oracle, n, q):
    rows = 3
    a, b = oracle.random((rows, n), q, 11 + n), oracle.random((rows, n), q, 13 + q)
    a[0], b[0] = q - 1, q - 1
    ref = oracle.product(n, q, a, b, 20)
    for v in (1, 4, 10):
        assert (oracle.product(n, q, a, b, v) == ref).all(), v


def test_cyclic_surface_q3329(oracle):
    """q = 3329 has no 512-th root of unity: only the psi-free (cyclic) surface exists at n=256."""
    n, q = 256, 3329
    with pytest.raises(ValueError):
        oracle.product(n, q, np.zeros((1, n), np.int32), np.zeros((1, n), np.int32), 10)
    a, b = oracle.random((2, n), q, 1), oracle.random((2, n), q, 2)
    c = oracle.product(n, q, a, b, 30)
    full = np.array([np.convolve(a[i].astype(object), b[i].astype(object)) for i in range(2)])
    want = (full[:, :n] + np.pad(full[:, n:], ((0, 0), (0, 1)))) % q
    assert (c == want.astype(np.int64)).all()


def test_splitmix_generator_is_stable(oracle):
    x = oracle.random((4,), 12289, 0x4E545442323030)
    assert x.tolist() == oracle.random((4,), 12289, 0x4E545442323030).tolist()
    assert all(0 <= v < 12289 for v in x.tolist())
