/*
 * gen_params.c -- include/nttb200_gen.h: the FPGA-datapath parameter generator and the text
 * formats of the reference, with runtime parameters.  Host C only.
 * G/ = Multiplier_NTT_Based/NTT_Software/Generator_Params/ of the reference.
 */
#include <stdio.h>
#include <stdlib.h>

#include "nttb200_gen.h"

uint32_t nttb200_modexp(uint32_t base, uint32_t exp, uint32_t mod) {
  uint64_t r = 1 % mod, b = base % mod;                     /* G/prime_generate.C:9-20 */
  for (; exp; exp >>= 1) {
    if (exp & 1) r = r * b % mod;
    b = b * b % mod;
  }
  return (uint32_t)r;
}

int32_t nttb200_modinv(int32_t a, int32_t m) {
  /* iterative extended Euclid; same result as G/helper.C:22-35 ((x % m + m) % m) */
  int64_t r0 = m, r1 = ((int64_t)a % m + m) % m, t0 = 0, t1 = 1;
  while (r1) {
    int64_t qd = r0 / r1, t;
    t = r0 - qd * r1; r0 = r1; r1 = t;
    t = t0 - qd * t1; t0 = t1; t1 = t;
  }
  if (r0 != 1) return -1;
  return (int32_t)((t0 % m + m) % m);
}

int nttb200_miller_rabin(uint32_t p) {
  /* the reference draws random witnesses (G/prime_generate.C:23-50); for 32-bit p the
   * bases 2, 7, 61 decide primality exactly (valid below 4,759,123,141) */
  static const uint32_t small[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37, 41, 43, 47, 53, 59, 61};
  if (p < 2) return 0;
  for (size_t i = 0; i < sizeof small / sizeof small[0]; i++) {
    if (p == small[i]) return 1;
    if (p % small[i] == 0) return 0;
  }
  uint32_t r = p - 1;
  int u = 0;
  while (!(r & 1)) { r >>= 1; u++; }
  static const uint32_t bases[] = {2, 7, 61};
  for (int i = 0; i < 3; i++) {
    uint64_t z = nttb200_modexp(bases[i], r, p);
    if (z == 1 || z == p - 1) continue;
    int j;
    for (j = 1; j < u; j++) {
      z = z * z % p;
      if (z == p - 1) break;
    }
    if (j == u) return 0;
  }
  return 1;
}

uint32_t nttb200_gen_prime(uint32_t k, uint32_t n, uint64_t seed) {
  if (k < 2 || k > 31 || n == 0) return 0;
  const uint64_t lo = 1ull << (k - 1), hi = 1ull << k, step = 2ull * n;
  /* candidates 1 (mod 2n) starting at a seed-dependent point of [2^(k-1), 2^k), wrapping once */
  const uint64_t slots = (hi - lo + step - 1) / step;
  uint64_t s = slots ? (seed * 0x9E3779B97F4A7C15ull >> 11) % slots : 0;
  for (uint64_t it = 0; it < slots; it++) {
    uint64_t c = ((lo + step - 1) / step + (s + it) % slots) * step + 1;
    if (c >= lo && c < hi && nttb200_miller_rabin((uint32_t)c)) return (uint32_t)c;
  }
  return 0;
}

static uint32_t ilog2u(uint32_t n) {
  uint32_t l = 0;
  while ((1u << l) < n) l++;
  return l;
}

int nttb200_gen_params(uint32_t n, uint32_t K, uint32_t P, uint32_t q, nttb200_gen_params_t *out) {
  if (!out || n < 2 || (n & (n - 1)) || K == 0 || P == 0) return -1;
  if (q == 0) q = 12289;                                      /* G/generate_params.C:22 */
  if (!nttb200_miller_rabin(q) || (uint64_t)(q - 1) % (2ull * n)) return -1;
  out->n = n;
  out->q = q;
  out->psi = 0;
  /* G/generate_params.C:25-44: the first i >= 2 with i^(2n) = 1, i^n = q-1 (and no smaller
   * power equal to 1, which i^n = -1 already implies for n a power of two) */
  for (uint32_t i = 2; i < q - 1; i++) {
    if (nttb200_modexp(i, n, q) == q - 1) { out->psi = i; break; }
  }
  if (!out->psi) return -1;
  out->psi_inv = (uint32_t)nttb200_modinv((int32_t)out->psi, (int32_t)q);
  out->w = (uint32_t)((uint64_t)out->psi * out->psi % q);
  out->w_inv = (uint32_t)nttb200_modinv((int32_t)out->w, (int32_t)q);
  const uint32_t logn = ilog2u(n);
  const uint32_t fator = (K + logn) / (logn + 1);             /* ceil(K / (logn+1)) */
  const uint32_t bits = (logn + 1) * fator;
  out->R = bits >= 32 ? 0 : (1u << bits);                     /* G/generate_params.C:47-49 */
  out->n_inv = (uint32_t)nttb200_modinv((int32_t)(n % q), (int32_t)q);
  out->PE = 2 * P;
  return 0;
}

size_t nttb200_gen_twiddle_count(uint32_t n, uint32_t P) {
  size_t total = 0;
  const uint32_t PE = 2 * P, logn = ilog2u(n);
  for (uint32_t j = 0; j < logn; j++) {
    size_t limit = (n / PE) >> j;
    if (limit < 1) limit = 1;
    total += limit * P;
  }
  return total;
}

int nttb200_gen_twiddles(uint32_t *W, uint32_t *W_INV, uint32_t n, uint32_t P, uint32_t w, uint32_t w_inv,
                         uint32_t q, uint32_t R) {
  if (!W || !W_INV || n < 2 || (n & (n - 1)) || P == 0 || q < 2) return -1;
  const uint32_t PE = 2 * P, logn = ilog2u(n);
  size_t idx = 0;
  for (uint32_t j = 0; j < logn; j++) {                       /* G/generate_params.C:57-72 */
    size_t limit = (n / PE) >> j;
    if (limit < 1) limit = 1;
    for (size_t k = 0; k < limit; k++)
      for (size_t i = 0; i < P; i++) {
        const uint32_t e = (uint32_t)((((size_t)P << j) * k + (i << j)) % (n / 2));
        W[idx] = (uint32_t)((uint64_t)nttb200_modexp(w, e, q) * R % q);
        W_INV[idx] = (uint32_t)((uint64_t)nttb200_modexp(w_inv, e, q) * R % q);
        idx++;
      }
  }
  return 0;
}

/* ---- reference signatures (N = 256, K = 13, P = 8, q = 12289) ------------------------- */
void generate_params(int *psi, int *psi_inv, int *w, int *w_inv, int *R, int *n_inv, int *PE, int *q) {
  nttb200_gen_params_t g;
  if (nttb200_gen_params(256, 13, 8, 12289, &g) != 0) abort();
  *psi = (int)g.psi; *psi_inv = (int)g.psi_inv; *w = (int)g.w; *w_inv = (int)g.w_inv;
  *R = (int)g.R; *n_inv = (int)g.n_inv; *PE = (int)g.PE; *q = (int)g.q;
}
void generate_twiddles(uint32_t W[], uint32_t W_INV[], uint32_t w, uint32_t w_inv, uint32_t q, uint32_t R) {
  nttb200_gen_twiddles(W, W_INV, 256, 8, w, w_inv, q, R);
}

/* ---- text formats ----------------------------------------------------------------------- */
int nttb200_write_coeff_file(const char *path, const int32_t *a, size_t n) {
  FILE *f = fopen(path, "w");
  if (!f) return -1;
  for (size_t i = 0; i < n; i++) {
    fprintf(f, "%d ", a[i]);
    if ((i + 1) % 10 == 0) fprintf(f, "\n");
  }
  return fclose(f) ? -1 : 0;
}
long nttb200_read_coeff_file(const char *path, int32_t *a, size_t n) {
  FILE *f = fopen(path, "r");
  if (!f) return -1;
  size_t got = 0;
  while (got < n) {
    int v;
    if (fscanf(f, "%d", &v) != 1) break;
    a[got++] = v;
  }
  fclose(f);
  return (long)got;
}
void nttb200_print_array(void *FILE_ptr, const int32_t *a, size_t n) {
  FILE *f = (FILE *)FILE_ptr;
  unsigned k = 0;
  for (size_t i = 0; i < n; i++) {
    if (k == 0) fprintf(f, "  ");
    fprintf(f, "%5d", a[i]);
    if (++k == 16) { fprintf(f, "\n"); k = 0; }
    else fprintf(f, " ");
  }
  if (k > 0) fprintf(f, "\n");
}
int nttb200_write_hex_file(const char *path, const uint32_t *a, size_t n) {
  FILE *f = fopen(path, "w");
  if (!f) return -1;
  for (size_t i = 0; i < n; i++) fprintf(f, "%x\n", a[i]);
  return fclose(f) ? -1 : 0;
}
long nttb200_read_hex_file(const char *path, uint32_t *a, size_t n) {
  FILE *f = fopen(path, "r");
  if (!f) return -1;
  size_t got = 0;
  while (got < n) {
    unsigned v;
    if (fscanf(f, "%x", &v) != 1) break;
    a[got++] = v;
  }
  fclose(f);
  return (long)got;
}
