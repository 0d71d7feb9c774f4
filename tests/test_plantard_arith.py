"""CPU check (numpy, exact integers) of the arithmetic claims the Plantard kernel
(ntt_small_plant.cuh) rests on -- no GPU, no library call: this pins the MATH; the kernel
itself is pinned against the oracle in the gpu tests.

  T = umulhi(Y * w~, q),  w~ = ((-w 2^32) mod q) q^-1 mod 2^32
      == Y w mod q, canonical, for every w in [0,q) and every Y < 28 q   (q <= 12385)
  umulhi(a b q^-1, q) == -a b 2^-32 mod q, canonical, whenever a b < 2^32
"""
import numpy as np
import pytest

M32 = (1 << 32) - 1


def qinv32(q):
    inv = 1
    for _ in range(6):
        inv = (inv * (2 - q * inv)) & M32
    assert (inv * q) & M32 == 1
    return inv


def plant_form(w, q, qinv):
    W = (q - ((w % q) << 32) % q) % q
    return (W * qinv) & M32


@pytest.mark.parametrize("q", [17, 97, 257, 3329, 7681, 12289, 12373])
def test_plantard_constant_multiplication_is_canonical_up_to_28q(q):
    qinv = qinv32(q)
    assert 28 * q * q < 1 << 32
    rng = np.random.default_rng(q)
    ws = np.arange(q, dtype=np.uint64) if q <= 3329 else np.unique(
        np.concatenate([rng.integers(0, q, 2000), [0, 1, 2, q - 1, q - 2, q // 2]])).astype(np.uint64)
    wt = np.array([plant_form(int(w), q, qinv) for w in ws], dtype=np.uint64)
    ys = np.unique(np.concatenate([rng.integers(0, 28 * q, 3000), np.arange(0, 40),
                                   28 * q - 1 - np.arange(0, 40), q * np.arange(1, 28), q * np.arange(1, 29) - 1]))
    ys = ys.astype(np.uint64)
    Y, WT = np.meshgrid(ys, wt, indexing="ij")
    _, Wc = np.meshgrid(ys, ws, indexing="ij")
    p = (Y * WT) & np.uint64(M32)                 # IMAD (low 32 bits)
    T = (p * np.uint64(q)) >> np.uint64(32)       # IMAD.HI.U32
    assert (T == (Y * Wc) % np.uint64(q)).all()


@pytest.mark.parametrize("q", [17, 97, 257, 3329, 7681, 12289, 12373, 12385])
def test_plantard_half_word_second_product_is_canonical_up_to_22q(q):
    """Variant B of the constant multiplication: only the upper half of p = Y w~ mod 2^32 meets q,
    T = ((p >> 16) + 1) q >> 16  (IMAD, SHF, IMAD, SHF: 2 multiplier slots instead of 3),
    == Y w mod q, canonical, whenever Y q + 2^16 q <= 2^32 -- every Y < 22 q for q <= 12385."""
    qinv = qinv32(q)
    assert 22 * q * q + (q << 16) <= 1 << 32
    rng = np.random.default_rng(q + 5)
    ws = np.arange(q, dtype=np.uint64) if q <= 3329 else np.unique(
        np.concatenate([rng.integers(0, q, 2000), [0, 1, 2, q - 1, q - 2, q // 2]])).astype(np.uint64)
    wt = np.array([plant_form(int(w), q, qinv) for w in ws], dtype=np.uint64)
    ys = np.unique(np.concatenate([rng.integers(0, 22 * q, 3000), np.arange(0, 40),
                                   22 * q - 1 - np.arange(0, 40), q * np.arange(1, 22), q * np.arange(1, 23) - 1]))
    ys = ys.astype(np.uint64)
    Y, WT = np.meshgrid(ys, wt, indexing="ij")
    _, Wc = np.meshgrid(ys, ws, indexing="ij")
    p = (Y * WT) & np.uint64(M32)                                   # IMAD (low 32 bits)
    t = ((p >> np.uint64(16)) * np.uint64(q) + np.uint64(q))        # IMAD on the upper half, + q
    assert (t < (1 << 32)).all()
    T = t >> np.uint64(16)
    assert (T == (Y * Wc) % np.uint64(q)).all()


@pytest.mark.parametrize("q", [3329, 7681, 12289])
def test_plantard_pointwise_product(q):
    qinv = qinv32(q)
    rng = np.random.default_rng(q + 1)
    r32inv = pow(1 << 32, -1, q)
    for ba, bb in ((5, 5), (3, 6), (1, 28), (4, 7)):
        assert ba * bb <= 28
        a = rng.integers(0, ba * q, 200000).astype(np.uint64)
        b = rng.integers(0, bb * q, 200000).astype(np.uint64)
        a[:4] = ba * q - 1
        b[:4] = bb * q - 1
        p = (((a * b) & np.uint64(M32)) * np.uint64(qinv)) & np.uint64(M32)
        T = (p * np.uint64(q)) >> np.uint64(32)
        want = (np.uint64(q) - (a * b % np.uint64(q)) * np.uint64(r32inv) % np.uint64(q)) % np.uint64(q)
        assert (T == want).all()


def test_gs_bound_closed_form_matches_simulation():
    """pl_gs_bound / PLANT_CAP logic (ntt_small_plant.cuh) re-stated and simulated: worst-case
    bounds of a register phase of GS stages with the cap-at-8 rule never exceed 8 at a stage
    input, so d = X - Y + b q < 16 q."""
    CAP = 8

    def bound(r, bit, b_in):
        h = r & ((1 << bit) - 1)
        if h == 0:
            return min(b_in << bit, CAP)
        return min(1 << (bit - 1 - (h.bit_length() - 1)), CAP)

    for bits in (1, 2, 3, 4, 5):
        for b_in in (1, 2, 4, 8):
            b = [b_in] * (1 << bits)
            for bit in range(bits):
                for r in range(1 << bits):
                    if r & (1 << bit):
                        continue
                    r2 = r | (1 << bit)
                    assert b[r] == b[r2] == bound(r, bit, b_in), (bits, b_in, bit, r)
                    assert b[r] <= CAP
                    s = 2 * b[r]
                    b[r] = s if s <= CAP else CAP          # one csub(s, 8q) when the inputs were at the cap
                    b[r2] = 1
            assert max(b) == min(b_in << bits, CAP)


# ------------------------------------------------------------------------------------------
# signed Plantard arithmetic of ntt_small_splant.cuh (five-instruction butterflies)
def splant_constants(q):
    """(D, C, Mmax) exactly as csrc/small_plant.cu computes them for the kernel parameters."""
    mmax = ((1 << 32) - 65536 * (q + 4)) // 2
    dd = (mmax + 65535 * q + 65535) // 65536
    cbar = ((1 << 26) + q // 2) // q
    return dd, cbar, mmax


def centred_form(w, q, qinv):
    W = (q - ((w % q) << 32) % q) % q
    if W > q // 2:
        W -= q
    return (W * qinv) & M32, W


def as_i32(x):
    x = np.asarray(x, dtype=np.int64) & M32
    return np.where(x >= 1 << 31, x - (1 << 32), x)


def splant_mul(y, wt, q, dd):
    """T = ((Y w~ as int32) >> 16) q + D) >> 16, arithmetic shifts, int64 emulation of the int32 code."""
    p = as_i32(np.asarray(y, dtype=np.int64) * np.asarray(wt, dtype=np.int64))
    u = (p >> 16) * q + dd
    assert (np.abs(u) < 1 << 31).all()
    return u >> 16


@pytest.mark.parametrize("q", [17, 97, 257, 3329, 7681, 12289, 12373, 12385])
def test_signed_plantard_multiplication_is_exact_and_centred(q):
    """T == Y w (mod q) and |T| <= (q-1)/2 for every centred W and every |Y| <= 16 q (the largest
    difference the kernel multiplies; the proof allows 23 q at q = 12289)."""
    qinv = qinv32(q)
    dd, _, mmax = splant_constants(q)
    ylim = 16 * q
    assert ylim * (q // 2) <= mmax and 2 * mmax + 65536 * (q + 4) <= 1 << 32
    rng = np.random.default_rng(q + 11)
    ws = np.arange(q) if q <= 3329 else np.unique(np.concatenate(
        [rng.integers(0, q, 1500), [0, 1, 2, q - 1, q - 2, q // 2, q // 2 + 1, q // 2 - 1]]))
    forms = [centred_form(int(w), q, qinv) for w in ws]
    wt = np.array([f[0] for f in forms], dtype=np.int64)
    ys = np.unique(np.concatenate([rng.integers(-ylim, ylim + 1, 3000), np.arange(-40, 41), ylim - np.arange(0, 40),
                                   -ylim + np.arange(0, 40), q * np.arange(-16, 17), q * np.arange(-16, 17) + 1,
                                   q * np.arange(-16, 17) - 1]))
    ys = ys[np.abs(ys) <= ylim]
    Y, WT = np.meshgrid(ys, wt, indexing="ij")
    _, Wn = np.meshgrid(ys, ws, indexing="ij")
    T = splant_mul(Y, WT, q, dd)
    assert (np.abs(T) <= (q - 1) // 2).all()
    assert ((T - Y * Wn) % q == 0).all()


@pytest.mark.parametrize("q", [17, 257, 3329, 7681, 12289, 12385])
def test_signed_barrett_step_and_pointwise_product(q):
    """sp_red: x - q ((x C + 2^25) >> 26) lands within q/2 + 20 of zero for |x| <= 16 q; the pointwise
    Plantard product of a reduced a and any |b| <= 6 q gives -a b 2^-32 mod q, centred."""
    qinv = qinv32(q)
    dd, cbar, mmax = splant_constants(q)
    rng = np.random.default_rng(q + 13)
    x = np.unique(np.concatenate([rng.integers(-16 * q, 16 * q + 1, 20000), q * np.arange(-16, 17),
                                  (q * np.arange(-31, 32)) // 2, (q * np.arange(-31, 32)) // 2 + 1]))
    assert (np.abs(x * cbar) < 1 << 31).all()
    r = x - q * ((x * cbar + (1 << 25)) >> 26)
    assert ((r - x) % q == 0).all() and (np.abs(r) <= q // 2 + 20).all()
    a = r[np.abs(x) <= 6 * q][:4000]
    b = rng.integers(-6 * q, 6 * q + 1, a.size)
    assert (np.abs(a * b) <= mmax).all()
    p = as_i32((as_i32(a * b) * qinv))
    u = (p >> 16) * q + dd
    T = u >> 16
    inv232 = pow(1 << 32, -1, q) if q > 2 else 1
    assert (np.abs(T) <= (q - 1) // 2).all()
    assert ((T + a * b * inv232) % q == 0).all()


def test_signed_bound_bookkeeping_of_the_inverse():
    """The compile-time bounds of ntt_small_splant.cuh (units of q/2), restated: no multiplied
    difference ever passes 32 units (16 q) and no Barrett input passes 32 units."""
    CAP = 16

    def leg(k, bit, b_in):
        b = b_in
        for s in range(bit):
            if (k >> s) & 1:
                b = 1
            else:
                b *= 2
                if b > CAP:
                    b = 1
        return b
    for H, R in ((4, 4), (3, 4), (3, 3), (2, 3), (2, 2), (1, 2)):
        worst = max(leg(k, H, 1) for k in range(1 << H))
        second = leg(1, H, 1)
        b_in = max(second, 2) if worst > second else worst
        for bit in range(H):
            for k in range(1 << H):
                assert 2 * leg(k, bit, 1) <= 2 * CAP
        for kb in range(R):
            for k in range(1 << R):
                assert 2 * leg(k, kb, b_in) <= 2 * CAP
