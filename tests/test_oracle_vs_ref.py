"""CPU: the oracle restatement against the compiled, unmodified reference run live
(oracle/_ref, built from /root/reference by oracle/Makefile; the prebuilt object travels to
the GPU box).  Skipped only if the object was never built."""
import numpy as np
import pytest

N, Q, PSI = 256, 12289, 1002


@pytest.fixture(scope="module")
def ref(loader):
    if not loader.reference_available():
        pytest.skip("oracle/_ref not built and /root/reference absent")
    return loader.Reference()


def test_all_reference_variants_agree_with_oracle(ref, oracle):
    a, b = oracle.random((3000, N), Q, 101), oracle.random((3000, N), Q, 102)
    a[0], b[0] = 0, 0
    a[1], b[1] = Q - 1, Q - 1
    want = oracle.product(N, Q, a, b, 10, PSI)
    for v in (1, 4, 101, 104, 10, 110):
        assert (ref.product(a, b, v) == want).all(), v


def test_reference_tables(ref, oracle):
    for k in range(11):
        assert (oracle.table(k, N, Q, PSI).astype(np.int64) == ref.table(k)).all()
        assert (oracle.red_table(k, N, PSI).astype(np.int64) == ref.table(k, red=True)).all()


def test_reference_transforms_live(ref, oracle):
    from test_oracle_golden import RED_IDS, TRANSFORM_IDS
    a = oracle.random((64, N), Q, 7)
    for tid, (name, kind) in TRANSFORM_IDS.items():
        assert (oracle.transform(name, a, oracle.table(kind, N, Q, PSI), Q) == ref.transform(tid, a)).all(), name
    c = a.copy()
    c[c > 6144] -= Q
    for tid, (name, kind) in RED_IDS.items():
        assert (oracle.red_transform(name, c, oracle.red_table(kind, N, PSI)) == ref.transform(tid, c)).all(), name


def test_red_range_helpers(ref, oracle):
    x = oracle.random((N,), Q, 5)
    for which, fn in enumerate(["shift_array", "reduce_array", "reduce_array_twice"]):
        r, o = x.copy(), x.copy()
        ref.lib.ref_red_helper(which, r, N)
        getattr(oracle.lib, "orc_red_" + fn)(o, N)
        assert (r == o).all(), fn
    y = (np.arange(N, dtype=np.int32) * 144) - Q          # covers [-Q, 2Q)
    r, o = y.copy(), y.copy()
    ref.lib.ref_red_helper(3, r, N)
    oracle.lib.orc_red_correct(o, N)
    assert (r == o).all() and (r == y % Q).all()
    z = (oracle.random((N,), 2**31 - 1, 9).astype(np.int64) - 2**30).astype(np.int32)
    for which, fn in ((4, "normalize"), (5, "normalize_inv3")):
        r, o = z.copy(), z.copy()
        ref.lib.ref_red_helper(which, r, N)
        getattr(oracle.lib, "orc_red_" + fn)(o, N)
        assert (r == o).all(), fn


def test_bitrev_shuffle(ref, oracle):
    x = np.arange(N, dtype=np.int32)
    r, o = x.copy(), x.copy()
    ref.lib.ref_bitrev_shuffle(r, N)
    oracle.lib.orc_bitrev_shuffle(o, N)
    assert (r == o).all()


def test_generic_entry_points_with_arbitrary_tables_live(ref, oracle):
    """Every generic transform of ntt.h / ntt_red.h with random caller tables (entries in
    [0, q), p[t] != 1): the oracle restates the j = 0 peel of the un-merged functions."""
    from test_oracle_golden import RED_TAB_IDS, TAB_IDS
    rng = np.random.default_rng(7)
    for nn in (8, 32, 256):
        for trial in range(4):
            tab = rng.integers(2, Q, size=nn).astype(np.uint16)
            a = oracle.random((5, nn), Q, 100 + trial)
            for tid, name in TAB_IDS.items():
                assert (oracle.transform(name, a, tab.astype(np.uint32), Q) == ref.transform_tab(tid, a, tab)).all(), (nn, name)
            rtab = tab.astype(np.int64)
            rtab[rtab > 6144] -= Q
            c = a.copy()
            c[c > 6144] -= Q
            for tid, name in RED_TAB_IDS.items():
                assert (oracle.red_transform(name, c, rtab.astype(np.int32)) ==
                        ref.transform_tab(tid, c, rtab.astype(np.int16), red=True)).all(), (nn, name)
