/*
 * host_tables.c -- number theory + twiddle tables on the host (plain C, no CUDA).
 * Implements nttb200_make_table / nttb200_find_psi / nttb200_find_omega /
 * nttb200_is_prime of include/nttb200.h.
 */
#include "host_tables.h"
#include "nttb200.h"

uint32_t ht_powmod(uint32_t base, uint64_t exp, uint32_t q) {
  uint64_t acc = 1, sq = base % q;
  for (; exp; exp >>= 1) {
    if (exp & 1) acc = acc * sq % q;
    sq = sq * sq % q;
  }
  return (uint32_t)acc;
}

uint32_t ht_invmod(uint32_t a, uint32_t q) { return ht_powmod(a, (uint64_t)q - 2, q); }

uint32_t ht_bitrev(uint32_t x, uint32_t bits) {
  uint32_t y = 0;
  while (bits--) { y = (y << 1) | (x & 1u); x >>= 1; }
  return y;
}

uint32_t ht_log2(uint32_t n) {
  uint32_t l = 0;
  while ((n >> l) > 1u) l++;
  return l;
}

/* deterministic Miller-Rabin, exact for 32-bit inputs with bases 2, 7, 61
 * (the reference's generator uses probabilistic rounds:
 *  Generator_Params/prime_generate.C:9-107) */
int nttb200_is_prime(uint32_t q) {
  static const uint32_t bases[3] = {2, 7, 61};
  if (q < 2) return 0;
  for (uint32_t p = 2; p < 62; p++) {
    if (q == p) return 1;
    if (q % p == 0) return 0;
  }
  uint32_t d = q - 1, r = 0;
  while ((d & 1u) == 0) { d >>= 1; r++; }
  for (int b = 0; b < 3; b++) {
    uint64_t x = ht_powmod(bases[b], d, q);
    if (x == 1 || x == q - 1) continue;
    int composite = 1;
    for (uint32_t i = 1; i < r; i++) {
      x = x * x % q;
      if (x == q - 1) { composite = 0; break; }
    }
    if (composite) return 0;
  }
  return 1;
}

/* smallest psi with psi^n = -1: order exactly 2n because n is a power of two
 * (same answer as the search in Generator_Params/generate_params.C:25-44) */
uint32_t nttb200_find_psi(uint32_t n, uint32_t q) {
  if (n < 1 || (n & (n - 1)) || q < 3 || ((uint64_t)(q - 1) % (2ull * n))) return 0;
  for (uint32_t g = 2; g < q - 1; g++)
    if (ht_powmod(g, n, q) == q - 1) return g;
  return 0;
}

uint32_t nttb200_find_omega(uint32_t n, uint32_t q) {
  if (n < 2 || (n & (n - 1)) || q < 3 || ((q - 1) % n)) return 0;
  for (uint32_t g = 2; g < q - 1; g++)
    if (ht_powmod(g, n / 2, q) == q - 1) return g;
  return 0;
}

void ht_level_table(uint32_t *out, uint32_t n, uint32_t q, uint32_t lead, uint32_t root, int rev) {
  out[0] = 0;
  uint32_t bits = 0;
  for (uint32_t t = 1; t < n; t <<= 1, bits++) {
    uint32_t stride = n / (2 * t);
    uint64_t lead_t = ht_powmod(lead, stride, q);
    uint32_t root_t = ht_powmod(root, stride, q);
    /* walk the powers of root_t once, scatter to (bit-reversed) slots */
    uint64_t cur = 1;
    for (uint32_t e = 0; e < t; e++) {
      uint32_t slot = rev ? ht_bitrev(e, bits) : e;
      out[t + slot] = (uint32_t)(lead_t * cur % q);
      cur = cur * root_t % q;
    }
  }
}

int nttb200_make_table(int kind, uint32_t n, uint32_t q, uint32_t psi, uint32_t *out) {
  if (!out || n < 2 || (n & (n - 1)) || q < 3 || psi == 0 || psi >= q) return NTTB200_EPARAM;
  uint32_t omega = (uint32_t)((uint64_t)psi * psi % q);
  uint32_t ipsi = ht_invmod(psi, q), iomega = ht_invmod(omega, q);
  uint64_t cur;
  switch (kind) {
    case NTTB200_PSI_POWERS:
      cur = 1;
      for (uint32_t i = 0; i < n; i++) { out[i] = (uint32_t)cur; cur = cur * psi % q; }
      return 0;
    case NTTB200_INV_PSI_POWERS:
      cur = 1;
      for (uint32_t i = 0; i < n; i++) { out[i] = (uint32_t)cur; cur = cur * ipsi % q; }
      return 0;
    case NTTB200_SCALED_INV_PSI_POWERS:
      cur = ht_invmod(n % q, q);
      for (uint32_t i = 0; i < n; i++) { out[i] = (uint32_t)cur; cur = cur * ipsi % q; }
      return 0;
    case NTTB200_INV_PSI_POWERS_REV: {
      uint32_t bits = ht_log2(n);
      cur = 1;
      for (uint32_t i = 0; i < n; i++) { out[ht_bitrev(i, bits)] = (uint32_t)cur; cur = cur * ipsi % q; }
      return 0;
    }
    case NTTB200_OMEGA_POWERS:         ht_level_table(out, n, q, 1, omega, 0); return 0;
    case NTTB200_OMEGA_POWERS_REV:     ht_level_table(out, n, q, 1, omega, 1); return 0;
    case NTTB200_INV_OMEGA_POWERS:     ht_level_table(out, n, q, 1, iomega, 0); return 0;
    case NTTB200_INV_OMEGA_POWERS_REV: ht_level_table(out, n, q, 1, iomega, 1); return 0;
    case NTTB200_MIXED_POWERS:         ht_level_table(out, n, q, psi, omega, 0); return 0;
    case NTTB200_MIXED_POWERS_REV:     ht_level_table(out, n, q, psi, omega, 1); return 0;
    case NTTB200_INV_MIXED_POWERS:     ht_level_table(out, n, q, ipsi, iomega, 0); return 0;
    case NTTB200_INV_MIXED_POWERS_REV: ht_level_table(out, n, q, ipsi, iomega, 1); return 0;
    default: return NTTB200_EPARAM;
  }
}

/* ---- Longa-Naehrig ("RED") table set, q = 12289 = 3 * 2^12 + 1 (R/NTT-RED/ntt_red256_tables.c:16-469,
 * SURVEY 8a-T): the NTT/ table of the same kind times 3^-1 mod q -- every mul_red() leaves a factor
 * 3 -- centred to (-q/2, q/2]; scaled_inv_psi_powers carries n^-1 3^-8 (the eight red()s of the
 * product, ntt_red256_tables.h:28 rescale8) and its _var twin n^-1 3^-6 (rescale6).  p[0] of the
 * level tables stays 0. */
#define RED_Q 12289u
int nttb200_make_red_table(int kind, uint32_t n, uint32_t psi, int32_t *out) {
  if (!out || n < 2 || (n & (n - 1)) || n > 65536) return NTTB200_EPARAM;
  const uint32_t inv3 = ht_invmod(3, RED_Q);
  uint32_t scale = inv3;
  int base = kind;
  if (kind == NTTB200_SCALED_INV_PSI_POWERS) scale = ht_powmod(inv3, 8, RED_Q);
  if (kind == NTTB200_RED_SCALED_INV_PSI_POWERS_VAR) { scale = ht_powmod(inv3, 6, RED_Q); base = NTTB200_SCALED_INV_PSI_POWERS; }
  uint32_t *u = (uint32_t *)out;                     /* same size: fill in place, then re-centre */
  const int rc = nttb200_make_table(base, n, RED_Q, psi, u);
  if (rc) return rc;
  const int level = (base >= NTTB200_OMEGA_POWERS && base <= NTTB200_INV_MIXED_POWERS_REV);
  for (uint32_t i = 0; i < n; i++) {
    const uint32_t v = (uint32_t)((uint64_t)u[i] * scale % RED_Q);
    out[i] = (level && i == 0) ? 0 : (v > (RED_Q - 1) / 2 ? (int32_t)v - (int32_t)RED_Q : (int32_t)v);
  }
  return 0;
}
