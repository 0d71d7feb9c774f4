"""CPU: property tests of the oracle (hypothesis) -- the checker itself is checked against the
mathematical definition on randomly drawn (n, q, inputs): every product variant equals the
negacyclic schoolbook product (colab_programs/schoolbook.py:23-46), the cyclic variant equals the
cyclic convolution, forward/inverse transforms invert each other up to n (R/NTT/ntt256.h:16-17),
and products are bilinear."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import loader

O = loader.Oracle()
# (n, q) with 2n | q-1, all three arithmetic ranges
CASES = [(8, 17), (8, 97), (16, 97), (32, 193), (64, 257), (64, 7681), (128, 3329), (128, 12289), (32, 8380417),
         (64, 998244353), (16, 2013265921)]


def negacyclic(a, b, q):
    n = len(a)
    full = np.convolve(a.astype(object), b.astype(object))
    full = np.concatenate([full, np.zeros(2 * n - len(full), dtype=object)])
    return np.array([(int(full[k]) - int(full[k + n])) % q for k in range(n)], dtype=np.int64)


def cyclic(a, b, q):
    n = len(a)
    full = np.convolve(a.astype(object), b.astype(object))
    full = np.concatenate([full, np.zeros(2 * n - len(full), dtype=object)])
    return np.array([(int(full[k]) + int(full[k + n])) % q for k in range(n)], dtype=np.int64)


@settings(max_examples=60, deadline=None)
@given(st.sampled_from(CASES), st.integers(0, 2**32 - 1))
def test_every_product_variant_is_the_schoolbook_product(case, seed):
    n, q = case
    a, b = O.random((2, n), q, seed), O.random((2, n), q, seed + 1)
    a[1], b[1] = q - 1, q - 1
    want = np.stack([negacyclic(a[i], b[i], q) for i in range(2)])
    for variant in (loader.PRODUCT_CT, loader.PRODUCT_GS, loader.PRODUCT_MERGED, loader.PRODUCT_SCHOOLBOOK):
        assert (O.product(n, q, a, b, variant) == want).all(), variant
    assert (O.product(n, q, a, b, loader.PRODUCT_CYCLIC) == np.stack([cyclic(a[i], b[i], q) for i in range(2)])).all()


@settings(max_examples=40, deadline=None)
@given(st.sampled_from(CASES), st.integers(0, 2**32 - 1))
def test_transforms_invert_each_other_and_products_are_bilinear(case, seed):
    n, q = case
    psi = O.psi(n, q, 0)
    a, b, c = O.random((1, n), q, seed), O.random((1, n), q, seed + 7), O.random((1, n), q, seed + 9)
    T = lambda k: O.table(k, n, q, psi)
    f = O.transform("ntt_ct_std2rev", a, T(loader.OMEGA_POWERS_REV), q)
    assert (f == O.transform("ntt_gs_std2rev", a, T(loader.OMEGA_POWERS), q)).all()
    back = O.transform("ntt_gs_rev2std", f, T(loader.INV_OMEGA_POWERS_REV), q)
    assert (back == a.astype(np.int64) * n % q).all()
    assert (back == O.transform("ntt_ct_rev2std", f, T(loader.INV_OMEGA_POWERS), q)).all()
    s = ((a.astype(np.int64) + c) % q).astype(np.int32)
    lhs = O.product(n, q, s, b, loader.PRODUCT_MERGED).astype(np.int64)
    rhs = (O.product(n, q, a, b, loader.PRODUCT_MERGED).astype(np.int64) + O.product(n, q, c, b, loader.PRODUCT_MERGED)) % q
    assert (lhs == rhs).all()
    assert (O.product(n, q, a, b, loader.PRODUCT_MERGED) == O.product(n, q, b, a, loader.PRODUCT_MERGED)).all()
