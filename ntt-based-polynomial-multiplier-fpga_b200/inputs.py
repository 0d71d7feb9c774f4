"""Synthetic operand batches exactly as SURVEY.md section 8d defines them.

Row-major ``int32 [batch][n]`` with coefficients i.i.d. on [0, q) from a fixed portable
PRNG -- splitmix64 seeded ``0x4E545442323030 + config_index`` (operand b: the stream continues
after a's ``batch * n`` draws), value = ``next() % q`` -- which is the reference's own practice
with another generator (``rand() % Q``, NTT-256/time_testing256.c:97-98,
Generator_Params/generate_coeff.c:47).  Rows 0..7 are then overwritten with edge rows:

    0  a = 0                      (c = 0)
    1  a = b = q-1 everywhere
    2  a = delta_0                (c = b)
    3  a = b = delta_{n-1}        (c = -x^(n-2))
    4  KAT (1)  (1+2x)(3)                     NTT-RED/test_prod_nttred256.c:48-57
    5  KAT (2)  (1+2x+3x^2)(2)                linux_app/NTT_PCIECommunicationv2.c:157-158
    6  KAT (3)  (1+2x+3x^2)(2+2x)             Hardware_Multiplier/NTT_PolyMul_test.v:110-145
    7  KAT (4)  (1+2x+2x^4)(3+3x+x^3)         NTT-256/time_testing256.c:73-79
    8  the reference's coefficient files (n = 256, q = 12289 only; caller passes them)

The generator is written with torch int64 tensor arithmetic so the same code fills a host
array (CPU tests compare it word for word with the oracle's ``orc_fill_random``) or a device
buffer of a full-size batch in milliseconds.  This is plumbing, not the product.
"""
from __future__ import annotations

SEED0 = 0x4E545442323030
_GAMMA = 0x9E3779B97F4A7C15
_M1 = 0xBF58476D1CE4E5B9
_M2 = 0x94D049BB133111EB


def _s64(x: int) -> int:
    x &= (1 << 64) - 1
    return x - (1 << 64) if x >= (1 << 63) else x


def splitmix64_mod(count: int, q: int, seed: int, offset: int = 0, device="cpu", out=None):
    """Draws offset+1 .. offset+count of splitmix64(seed), each reduced mod q; int32 tensor."""
    import torch
    if out is None:
        out = torch.empty(count, dtype=torch.int32, device=device)
    step = 1 << 24
    g, m1, m2 = _s64(_GAMMA), _s64(_M1), _s64(_M2)
    lsr = lambda z, k: (z >> k) & ((1 << (64 - k)) - 1)
    two32 = (1 << 32) % q
    for lo in range(0, count, step):
        hi = min(count, lo + step)
        i = torch.arange(offset + lo + 1, offset + hi + 1, dtype=torch.int64, device=out.device)
        z = i * g + _s64(seed)
        z = (z ^ lsr(z, 30)) * m1
        z = (z ^ lsr(z, 27)) * m2
        z = z ^ lsr(z, 31)
        zh, zl = lsr(z, 32), z & 0xFFFFFFFF
        out[lo:hi] = (((zh % q) * two32 + (zl % q)) % q).to(torch.int32)
    return out


def _poly(torch, n, coeffs, device):
    p = torch.zeros(n, dtype=torch.int32, device=device)
    for k, v in coeffs.items():
        p[k] = v
    return p


KATS = (({0: 1, 1: 2}, {0: 3}),
        ({0: 1, 1: 2, 2: 3}, {0: 2}),
        ({0: 1, 1: 2, 2: 3}, {0: 2, 1: 2}),
        ({0: 1, 1: 2, 4: 2}, {0: 3, 1: 3, 3: 1}))
EDGE_ROWS = 8


def survey_batch(n: int, q: int, batch: int, config_index: int, device="cpu", fixture=None,
                 row_offset: int = 0):
    """(a, b) int32 [batch][n] of config `config_index`.  `row_offset`: this call holds rows
    [row_offset, row_offset + batch) of the whole batch (a rank's shard); the edge rows exist in
    the shard that owns them only.  `fixture` = (fa, fb) arrays for row 8 at (256, 12289)."""
    import torch
    seed = SEED0 + config_index
    a = splitmix64_mod(batch * n, q, seed, offset=row_offset * n, device=device).view(batch, n)
    # operand b continues the stream one whole-batch stride later; 2^40 draws keep it clear of
    # a for every configuration (largest: 2^28 words per operand)
    b = splitmix64_mod(batch * n, q, seed, offset=(1 << 40) + row_offset * n, device=device).view(batch, n)

    def put(row, ra=None, rb=None):
        r = row - row_offset
        if 0 <= r < batch:
            if ra is not None:
                a[r] = ra
            if rb is not None:
                b[r] = rb

    z = torch.zeros(n, dtype=torch.int32, device=a.device)
    full = torch.full((n,), q - 1, dtype=torch.int32, device=a.device)
    put(0, z)
    put(1, full, full)
    put(2, _poly(torch, n, {0: 1}, a.device))
    d = _poly(torch, n, {n - 1: 1}, a.device)
    put(3, d, d)
    for k, (ka, kb) in enumerate(KATS):
        if n >= 8 and q > 3:
            put(4 + k, _poly(torch, n, ka, a.device), _poly(torch, n, kb, a.device))
    if fixture is not None and n == 256 and q == 12289:
        fa, fb = fixture
        put(8, torch.as_tensor(fa, dtype=torch.int32).to(a.device), torch.as_tensor(fb, dtype=torch.int32).to(a.device))
    return a, b
