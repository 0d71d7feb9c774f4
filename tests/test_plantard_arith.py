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
