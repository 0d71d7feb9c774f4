/*
 * ntt_oracle.c -- CPU oracle (TEST INFRASTRUCTURE ONLY; see ntt_oracle.h).
 *
 * Parametrised restatement of the reference's NTT_Software algorithms with runtime
 * (n, q, psi) and 32-bit tables.  Every function cites the reference lines it
 * follows; R/ stands for
 *   /root/reference/Multiplier_NTT_Based/NTT_Software/NTT_Software_Evaluations/NTT-256/
 * The reference's q-specific tricks (Barrett magic 178942409>>41, the >>14 sign
 * masks) are replaced by their mathematical meaning (x mod q on canonical inputs),
 * which is what makes other moduli possible; at (256, 12289, 1002) the results are
 * bit-identical to the compiled reference (tests/test_oracle_vs_ref.py).
 */
#define _POSIX_C_SOURCE 199309L
#include "ntt_oracle.h"

#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ------------------------------------------------------------------------- */
/* scalar arithmetic: R/NTT/ntt.C:69-107                                      */
/* ------------------------------------------------------------------------- */

/* modq(x) for 0 <= x <= (q-1)^2   (ntt.C:101-107: x - divq(x)*Q) */
static inline int32_t mod_q(uint64_t x, uint32_t q) { return (int32_t)(x % q); }
/* ntt.C:69-74 -- x,y in [0,q): x-y, plus q when negative */
static inline int32_t sub_q(int32_t x, int32_t y, uint32_t q) {
  int64_t d = (int64_t)x - y;
  return (int32_t)(d < 0 ? d + q : d);
}
/* ntt.C:76-81 -- x+y-q, plus q when negative */
static inline int32_t add_q(int32_t x, int32_t y, uint32_t q) {
  int64_t s = (int64_t)x + y - q;
  return (int32_t)(s < 0 ? s + q : s);
}
static inline int32_t mul_q(int32_t x, uint32_t w, uint32_t q) {
  return mod_q((uint64_t)(uint32_t)x * w, q);
}

uint32_t orc_powmod(uint32_t b, uint64_t e, uint32_t q) {
  uint64_t r = 1, x = b % q;
  while (e) {
    if (e & 1) r = r * x % q;
    x = x * x % q;
    e >>= 1;
  }
  return (uint32_t)r;
}

uint32_t orc_invmod(uint32_t a, uint32_t q) { return orc_powmod(a, (uint64_t)q - 2, q); }

int orc_is_prime(uint32_t q) {
  if (q < 2) return 0;
  if (q % 2 == 0) return q == 2;
  for (uint64_t d = 3; d * d <= q; d += 2)
    if (q % d == 0) return 0;
  return 1;
}

/* Generator_Params/generate_params.C:25-44: first i>=2 with i^n == q-1; since n is a
 * power of two that alone makes the order exactly 2n. */
uint32_t orc_smallest_psi(uint32_t n, uint32_t q) {
  if (n == 0 || (n & (n - 1)) || ((uint64_t)(q - 1) % (2ull * n)) != 0) return 0;
  for (uint32_t i = 2; i < q - 1; i++)
    if (orc_powmod(i, n, q) == q - 1) return i;
  return 0;
}

uint32_t orc_smallest_omega(uint32_t n, uint32_t q) {
  if (n == 0 || (n & (n - 1)) || ((q - 1) % n) != 0) return 0;
  if (n == 1) return 1;
  for (uint32_t i = 2; i < q - 1; i++)
    if (orc_powmod(i, n / 2, q) == q - 1) return i;
  return 0;
}

/* ------------------------------------------------------------------------- */
/* tables: definitions verified entry-for-entry against R/NTT/ntt256_tables.C  */
/* ------------------------------------------------------------------------- */

static uint32_t bitrev(uint32_t x, uint32_t bits) {
  uint32_t r = 0;
  for (uint32_t i = 0; i < bits; i++) r |= ((x >> i) & 1u) << (bits - 1 - i);
  return r;
}
static uint32_t ilog2(uint32_t n) {
  uint32_t l = 0;
  while ((1u << l) < n) l++;
  return l;
}

/* level table: p[t+j] = lead^(n/2t) * root^((n/2t) * (rev? rev_t(j) : j)) */
static void level_table(uint32_t *out, uint32_t n, uint32_t q, uint32_t lead, uint32_t root,
                        int rev) {
  out[0] = 0;
  for (uint32_t t = 1, lt = 0; t < n; t <<= 1, lt++) {
    uint32_t step = n / (2 * t);
    uint32_t lead_t = orc_powmod(lead, step, q);
    for (uint32_t j = 0; j < t; j++) {
      uint32_t e = rev ? bitrev(j, lt) : j;
      out[t + j] = (uint32_t)((uint64_t)lead_t * orc_powmod(root, (uint64_t)step * e, q) % q);
    }
  }
}

int orc_make_omega_table(int kind, uint32_t n, uint32_t q, uint32_t omega, uint32_t *out) {
  uint32_t iomega = orc_invmod(omega, q);
  switch (kind) {
    case ORC_OMEGA_POWERS:         level_table(out, n, q, 1, omega, 0); return 0;
    case ORC_OMEGA_POWERS_REV:     level_table(out, n, q, 1, omega, 1); return 0;
    case ORC_INV_OMEGA_POWERS:     level_table(out, n, q, 1, iomega, 0); return 0;
    case ORC_INV_OMEGA_POWERS_REV: level_table(out, n, q, 1, iomega, 1); return 0;
    default: return -1;
  }
}

int orc_make_table(int kind, uint32_t n, uint32_t q, uint32_t psi, uint32_t *out) {
  uint32_t omega = (uint32_t)((uint64_t)psi * psi % q);
  uint32_t ipsi = orc_invmod(psi, q), iomega = orc_invmod(omega, q);
  uint32_t ninv = orc_invmod(n % q, q);
  uint32_t ln = ilog2(n);
  switch (kind) {
    case ORC_PSI_POWERS:
      for (uint32_t i = 0; i < n; i++) out[i] = orc_powmod(psi, i, q);
      return 0;
    case ORC_INV_PSI_POWERS:
      for (uint32_t i = 0; i < n; i++) out[i] = orc_powmod(ipsi, i, q);
      return 0;
    case ORC_INV_PSI_POWERS_REV:
      for (uint32_t i = 0; i < n; i++) out[i] = orc_powmod(ipsi, bitrev(i, ln), q);
      return 0;
    case ORC_SCALED_INV_PSI_POWERS:
      for (uint32_t i = 0; i < n; i++)
        out[i] = (uint32_t)((uint64_t)ninv * orc_powmod(ipsi, i, q) % q);
      return 0;
    case ORC_OMEGA_POWERS: case ORC_OMEGA_POWERS_REV:
    case ORC_INV_OMEGA_POWERS: case ORC_INV_OMEGA_POWERS_REV:
      return orc_make_omega_table(kind, n, q, omega, out);
    case ORC_MIXED_POWERS:         level_table(out, n, q, psi, omega, 0); return 0;
    case ORC_MIXED_POWERS_REV:     level_table(out, n, q, psi, omega, 1); return 0;
    case ORC_INV_MIXED_POWERS:     level_table(out, n, q, ipsi, iomega, 0); return 0;
    case ORC_INV_MIXED_POWERS_REV: level_table(out, n, q, ipsi, iomega, 1); return 0;
    default: return -1;
  }
}

/* ------------------------------------------------------------------------- */
/* elementwise: R/NTT/ntt.C:119-153, bit reversal ntt.C:27-44                  */
/* ------------------------------------------------------------------------- */

void orc_mul_array_tab(int32_t *a, uint32_t n, const uint32_t *p, uint32_t q) {
  for (uint32_t i = 0; i < n; i++) a[i] = mul_q(a[i], p[i], q);           /* ntt.C:122-124 */
}
void orc_mul_array(int32_t *c, uint32_t n, const int32_t *a, const int32_t *b, uint32_t q) {
  for (uint32_t i = 0; i < n; i++) c[i] = mul_q(a[i], (uint32_t)b[i], q); /* ntt.C:134-136 */
}
void orc_scalar_mul_array(int32_t *a, uint32_t n, int32_t c, uint32_t q) {
  for (uint32_t i = 0; i < n; i++) a[i] = mul_q(a[i], (uint32_t)c, q);    /* ntt.C:150-152 */
}
void orc_bitrev_shuffle(int32_t *a, uint32_t n) {
  uint32_t bits = ilog2(n);                     /* same permutation as ntt.C:33-43 */
  for (uint32_t i = 0; i < n; i++) {
    uint32_t j = bitrev(i, bits);
    if (i < j) { int32_t x = a[i]; a[i] = a[j]; a[j] = x; }
  }
}

/* ------------------------------------------------------------------------- */
/* transforms.  The reference's un-merged variants peel the j=0 block and do   */
/* its butterflies WITHOUT a multiplication and without reading p[t]           */
/* (ntt.C:178-183, 226-231, 313-317, 401-405, 477-481); the psi-merged ones    */
/* multiply every block by p[t+j].  With the reference's own tables p[t] = 1   */
/* and the two agree; with a caller's table they do not, so each dataflow      */
/* takes `peel0` and the un-merged entry points pass 1.                        */
/* ------------------------------------------------------------------------- */

/* CT butterfly (ntt.C:323-326): x = a[hi]*w; a[hi] = a[lo]-x; a[lo] = a[lo]+x */
#define CT_BFLY(lo, hi, w)                      \
  do {                                          \
    int32_t x_ = mul_q(a[hi], (w), q);          \
    int32_t u_ = a[lo];                         \
    a[hi] = sub_q(u_, x_, q);                   \
    a[lo] = add_q(u_, x_, q);                   \
  } while (0)
/* the peeled j=0 butterfly of both families (e.g. ntt.C:313-317, 401-405):
 * x = a[hi]; a[hi] = a[lo]-x; a[lo] = a[lo]+x */
#define PLAIN_BFLY(lo, hi)                      \
  do {                                          \
    int32_t x_ = a[hi];                         \
    int32_t u_ = a[lo];                         \
    a[hi] = sub_q(u_, x_, q);                   \
    a[lo] = add_q(u_, x_, q);                   \
  } while (0)
/* GS butterfly (ntt.C:408-411): x = a[hi]; a[hi] = (a[lo]-x)*w; a[lo] = a[lo]+x */
#define GS_BFLY(lo, hi, w)                      \
  do {                                          \
    int32_t x_ = a[hi];                         \
    int32_t u_ = a[lo];                         \
    a[hi] = mul_q(sub_q(u_, x_, q), (w), q);    \
    a[lo] = add_q(u_, x_, q);                   \
  } while (0)

/* ntt.C:168-197: stride-t butterflies, twiddle psi^(l*j) read from the psi-power table */
void orc_ntt_ct_rev2std_v1(int32_t *a, uint32_t n, const uint32_t *p, uint32_t q) {
  for (uint32_t t = 1, l = n; t < n; t <<= 1, l >>= 1)
    for (uint32_t j = 0; j < t; j++) {
      if (j == 0) {                                          /* ntt.C:178-183: p[0] is not read */
        for (uint32_t s = 0; s < n; s += 2 * t) PLAIN_BFLY(s, s + t);
        continue;
      }
      uint32_t w = p[j * l];
      for (uint32_t s = j; s < n; s += 2 * t) CT_BFLY(s, s + t, w);
    }
}
/* ntt.C:216-243 (plain) and ntt.C:253-278 (psi-merged): w = p[t+j], j = offset in block */
static void ct_rev2std(int32_t *a, uint32_t n, const uint32_t *p, uint32_t q, int peel0) {
  for (uint32_t t = 1; t < n; t <<= 1)
    for (uint32_t j = 0; j < t; j++) {
      if (peel0 && j == 0) {                                 /* ntt.C:226-231 */
        for (uint32_t s = 0; s < n; s += 2 * t) PLAIN_BFLY(s, s + t);
        continue;
      }
      uint32_t w = p[t + j];
      for (uint32_t s = j; s < n; s += 2 * t) CT_BFLY(s, s + t, w);
    }
}
/* ntt.C:295-329 (plain) and ntt.C:342-371 (psi-merged): w = p[t+j], j = block index,
 * block j spans [2dj, 2dj+2d), d = n/2t */
static void ct_std2rev(int32_t *a, uint32_t n, const uint32_t *p, uint32_t q, int peel0) {
  uint32_t d = n;
  for (uint32_t t = 1; t < n; t <<= 1) {
    d >>= 1;
    for (uint32_t j = 0, u = 0; j < t; j++, u += 2 * d) {
      if (peel0 && j == 0) {                                 /* ntt.C:313-317 */
        for (uint32_t s = 0; s < d; s++) PLAIN_BFLY(s, s + d);
        continue;
      }
      uint32_t w = p[t + j];
      for (uint32_t s = u; s < u + d; s++) CT_BFLY(s, s + d, w);
    }
  }
}
/* ntt.C:387-416 (plain) and ntt.C:428-451 (psi-merged) */
static void gs_rev2std(int32_t *a, uint32_t n, const uint32_t *p, uint32_t q, int peel0) {
  uint32_t t = n;
  for (uint32_t d = 1; d < n; d <<= 1) {
    t >>= 1;
    for (uint32_t j = 0, u = 0; j < t; j++, u += 2 * d) {
      if (peel0 && j == 0) {                                 /* ntt.C:401-405 */
        for (uint32_t s = 0; s < d; s++) PLAIN_BFLY(s, s + d);
        continue;
      }
      uint32_t w = p[t + j];
      for (uint32_t s = u; s < u + d; s++) GS_BFLY(s, s + d, w);
    }
  }
}
/* ntt.C:467-493 (plain) and ntt.C:505-525 (psi-merged) */
static void gs_std2rev(int32_t *a, uint32_t n, const uint32_t *p, uint32_t q, int peel0) {
  for (uint32_t t = n >> 1; t > 0; t >>= 1)
    for (uint32_t j = 0; j < t; j++) {
      if (peel0 && j == 0) {                                 /* ntt.C:477-481 */
        for (uint32_t s = 0; s < n; s += 2 * t) PLAIN_BFLY(s, s + t);
        continue;
      }
      uint32_t w = p[t + j];
      for (uint32_t s = j; s < n; s += 2 * t) GS_BFLY(s, s + t, w);
    }
}

void orc_ntt_ct_rev2std   (int32_t *a, uint32_t n, const uint32_t *p, uint32_t q) { ct_rev2std(a, n, p, q, 1); }
void orc_mulntt_ct_rev2std(int32_t *a, uint32_t n, const uint32_t *p, uint32_t q) { ct_rev2std(a, n, p, q, 0); }
void orc_ntt_ct_std2rev   (int32_t *a, uint32_t n, const uint32_t *p, uint32_t q) { ct_std2rev(a, n, p, q, 1); }
void orc_mulntt_ct_std2rev(int32_t *a, uint32_t n, const uint32_t *p, uint32_t q) { ct_std2rev(a, n, p, q, 0); }
void orc_ntt_gs_rev2std   (int32_t *a, uint32_t n, const uint32_t *p, uint32_t q) { gs_rev2std(a, n, p, q, 1); }
void orc_nttmul_gs_rev2std(int32_t *a, uint32_t n, const uint32_t *p, uint32_t q) { gs_rev2std(a, n, p, q, 0); }
void orc_ntt_gs_std2rev   (int32_t *a, uint32_t n, const uint32_t *p, uint32_t q) { gs_std2rev(a, n, p, q, 1); }
void orc_nttmul_gs_std2rev(int32_t *a, uint32_t n, const uint32_t *p, uint32_t q) { gs_std2rev(a, n, p, q, 0); }

/* ------------------------------------------------------------------------- */
/* plans + products: R/NTT/ntt256.C:5-24                                       */
/* ------------------------------------------------------------------------- */

struct orc_plan {
  uint32_t n, q, psi, omega, ninv;
  uint32_t *tab[ORC_TABLE_COUNT];
  /* cyclic surface (exists whenever n | q-1, even when 2n does not) */
  uint32_t *cyc_fwd_rev, *cyc_inv;   /* omega_powers_rev / inv_omega_powers, smallest omega */
};

orc_plan *orc_plan_create(uint32_t n, uint32_t q, uint32_t psi) {
  if (n < 2 || (n & (n - 1)) || !orc_is_prime(q) || q >= (1u << 31)) return NULL;
  orc_plan *P = (orc_plan *)calloc(1, sizeof *P);
  if (!P) return NULL;
  P->n = n; P->q = q;
  P->ninv = orc_invmod(n % q, q);
  if (((uint64_t)(q - 1) % (2ull * n)) == 0) {
    if (psi == 0) psi = orc_smallest_psi(n, q);
    if (psi == 0 || orc_powmod(psi, n, q) != q - 1) { free(P); return NULL; }
    P->psi = psi;
    P->omega = (uint32_t)((uint64_t)psi * psi % q);
    for (int k = 0; k < ORC_TABLE_COUNT; k++) {
      P->tab[k] = (uint32_t *)malloc(sizeof(uint32_t) * n);
      orc_make_table(k, n, q, psi, P->tab[k]);
    }
  } else if ((q - 1) % n != 0) {
    free(P); return NULL;
  }
  uint32_t om = P->omega ? P->omega : orc_smallest_omega(n, q);
  P->cyc_fwd_rev = (uint32_t *)malloc(sizeof(uint32_t) * n);
  P->cyc_inv = (uint32_t *)malloc(sizeof(uint32_t) * n);
  orc_make_omega_table(ORC_OMEGA_POWERS_REV, n, q, om, P->cyc_fwd_rev);
  orc_make_omega_table(ORC_INV_OMEGA_POWERS, n, q, om, P->cyc_inv);
  return P;
}
void orc_plan_destroy(orc_plan *P) {
  if (!P) return;
  for (int k = 0; k < ORC_TABLE_COUNT; k++) free(P->tab[k]);
  free(P->cyc_fwd_rev); free(P->cyc_inv);
  free(P);
}
uint32_t orc_plan_psi(const orc_plan *P) { return P->psi; }
const uint32_t *orc_plan_table(const orc_plan *P, int kind) {
  return (kind >= 0 && kind < ORC_TABLE_COUNT) ? P->tab[kind] : NULL;
}

static void schoolbook(const orc_plan *P, int32_t *c, const int32_t *a, const int32_t *b) {
  /* colab_programs/schoolbook.py:23-46: res[k] = conv[k] - conv[k+n] (mod q) */
  uint32_t n = P->n, q = P->q;
  for (uint32_t k = 0; k < n; k++) {
    uint64_t lo = 0, hi = 0;
    for (uint32_t i = 0; i <= k; i++) lo = (lo + (uint64_t)(uint32_t)a[i] * (uint32_t)b[k - i]) % q;
    for (uint32_t i = k + 1; i < n; i++) hi = (hi + (uint64_t)(uint32_t)a[i] * (uint32_t)b[n + k - i]) % q;
    c[k] = (int32_t)((lo + q - hi) % q);
  }
}

int orc_product(const orc_plan *P, int variant, int32_t *c, int32_t *a, int32_t *b) {
  uint32_t n = P->n, q = P->q;
  if (variant != ORC_PRODUCT_CYCLIC && !P->psi) return -1;
  switch (variant) {
    case ORC_PRODUCT_CT:                                    /* ntt256.C:5-13 */
      orc_mul_array_tab(a, n, P->tab[ORC_PSI_POWERS], q);
      orc_ntt_ct_std2rev(a, n, P->tab[ORC_OMEGA_POWERS_REV], q);
      orc_mul_array_tab(b, n, P->tab[ORC_PSI_POWERS], q);
      orc_ntt_ct_std2rev(b, n, P->tab[ORC_OMEGA_POWERS_REV], q);
      orc_mul_array(c, n, a, b, q);
      orc_ntt_ct_rev2std(c, n, P->tab[ORC_INV_OMEGA_POWERS], q);
      orc_mul_array_tab(c, n, P->tab[ORC_SCALED_INV_PSI_POWERS], q);
      return 0;
    case ORC_PRODUCT_GS:                                    /* ntt256.C:16-24 */
      orc_mul_array_tab(a, n, P->tab[ORC_PSI_POWERS], q);
      orc_ntt_gs_std2rev(a, n, P->tab[ORC_OMEGA_POWERS], q);
      orc_mul_array_tab(b, n, P->tab[ORC_PSI_POWERS], q);
      orc_ntt_gs_std2rev(b, n, P->tab[ORC_OMEGA_POWERS], q);
      orc_mul_array(c, n, a, b, q);
      orc_ntt_gs_rev2std(c, n, P->tab[ORC_INV_OMEGA_POWERS_REV], q);
      orc_mul_array_tab(c, n, P->tab[ORC_SCALED_INV_PSI_POWERS], q);
      return 0;
    case ORC_PRODUCT_MERGED:        /* ntt256.h:62-69 wrappers composed as in SURVEY facts */
      orc_mulntt_ct_std2rev(a, n, P->tab[ORC_MIXED_POWERS_REV], q);
      orc_mulntt_ct_std2rev(b, n, P->tab[ORC_MIXED_POWERS_REV], q);
      orc_mul_array(c, n, a, b, q);
      orc_nttmul_gs_rev2std(c, n, P->tab[ORC_INV_MIXED_POWERS_REV], q);
      orc_scalar_mul_array(c, n, (int32_t)P->ninv, q);
      return 0;
    case ORC_PRODUCT_CYCLIC:        /* psi-free surface ntt256.h:28,37 + ntt.h:52 */
      orc_ntt_ct_std2rev(a, n, P->cyc_fwd_rev, q);
      orc_ntt_ct_std2rev(b, n, P->cyc_fwd_rev, q);
      orc_mul_array(c, n, a, b, q);
      orc_ntt_ct_rev2std(c, n, P->cyc_inv, q);
      orc_scalar_mul_array(c, n, (int32_t)P->ninv, q);
      return 0;
    case ORC_PRODUCT_SCHOOLBOOK:
      schoolbook(P, c, a, b);
      return 0;
    default:
      return -1;
  }
}

int orc_product_batch(const orc_plan *P, int variant, int32_t *c, const int32_t *a,
                      const int32_t *b, size_t batch) {
  uint32_t n = P->n;
  int32_t *ta = (int32_t *)malloc(sizeof(int32_t) * n * 2);
  if (!ta) return -2;
  int32_t *tb = ta + n;
  int rc = 0;
  for (size_t r = 0; r < batch && rc == 0; r++) {
    memcpy(ta, a + r * n, sizeof(int32_t) * n);
    memcpy(tb, b + r * n, sizeof(int32_t) * n);
    rc = orc_product(P, variant, c + r * n, ta, tb);
  }
  free(ta);
  return rc;
}

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);          /* time_testing256.c:178 uses the same clock */
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

double orc_bench_loop(const orc_plan *P, int variant, const int32_t *a, const int32_t *b,
                      size_t batch, double min_seconds, uint64_t *calls_out) {
  uint32_t n = P->n;
  size_t ch = 65536 / n ? 65536 / n : 1;   /* restore `ch` operand pairs, then time `ch` calls */
  int32_t *buf = (int32_t *)malloc(sizeof(int32_t) * n * (2 * ch + 1));
  if (!buf) return 0.0;
  int32_t *ta = buf, *tb = buf + n * ch, *tc = buf + 2 * n * ch;
  double busy = 0.0, t_begin = now_s();
  uint64_t calls = 0;
  volatile int32_t sink = 0;
  while (now_s() - t_begin < min_seconds) {
    for (size_t r = 0; r < batch; r += ch) {
      size_t m = batch - r < ch ? batch - r : ch;
      memcpy(ta, a + r * n, m * n * sizeof(int32_t));
      memcpy(tb, b + r * n, m * n * sizeof(int32_t));
      double t0 = now_s();
      for (size_t k = 0; k < m; k++) {
        orc_product(P, variant, tc, ta + k * n, tb + k * n);
        sink ^= tc[0];
      }
      busy += now_s() - t0;
      calls += m;
    }
  }
  (void)sink;
  free(buf);
  if (calls_out) *calls_out = calls;
  return busy > 0 ? calls / busy : 0.0;
}

/* ------------------------------------------------------------------------- */
/* q = 12289 Longa-Naehrig lazy-reduction path: R/NTT-RED/ntt_red.c            */
/* ------------------------------------------------------------------------- */
#define RQ 12289

int32_t orc_red(int32_t x) { return 3 * (x & 4095) - (x >> 12); }            /* :34-36 */
int32_t orc_mul_red(int32_t x, int32_t y) {                                   /* :39-46 */
  int64_t z = (int64_t)x * y;
  return (int32_t)(3 * (z & 4095) - (z >> 12));
}
void orc_red_shift_array(int32_t *a, uint32_t n) {                            /* :103-111 */
  for (uint32_t i = 0; i < n; i++) if (a[i] > (RQ - 1) / 2) a[i] -= RQ;
}
void orc_red_reduce_array(int32_t *a, uint32_t n) {                           /* :124-130 */
  for (uint32_t i = 0; i < n; i++) a[i] = orc_red(a[i]);
}
void orc_red_reduce_array_twice(int32_t *a, uint32_t n) {                     /* :138-144 */
  for (uint32_t i = 0; i < n; i++) a[i] = orc_red(orc_red(a[i]));
}
void orc_red_correct(int32_t *a, uint32_t n) {                                /* :150-169 */
  for (uint32_t i = 0; i < n; i++) {
    int32_t x = a[i];
    if (x < 0) x += RQ;          /* x += (x>>16)&Q        */
    x -= RQ;                     /* x -= Q                */
    if (x < 0) x += RQ;          /* x += (x>>16)&Q        */
    a[i] = x;
  }
}
void orc_red_normalize(int32_t *a, uint32_t n) {                              /* :72-82 */
  for (uint32_t i = 0; i < n; i++) { int32_t x = a[i] % RQ; a[i] = x < 0 ? x + RQ : x; }
}
void orc_red_normalize_inv3(int32_t *a, uint32_t n) {                         /* :87-97 */
  for (uint32_t i = 0; i < n; i++) {
    int32_t x = (int32_t)(((int64_t)a[i] * 8193) % RQ);
    a[i] = x < 0 ? x + RQ : x;
  }
}
void orc_red_mul_reduce_array_tab(int32_t *a, uint32_t n, const int32_t *p) { /* :197-203 */
  for (uint32_t i = 0; i < n; i++) a[i] = orc_mul_red(a[i], p[i]);
}
void orc_red_mul_reduce_array(int32_t *c, uint32_t n, const int32_t *a, const int32_t *b) {
  for (uint32_t i = 0; i < n; i++) c[i] = orc_mul_red(a[i], b[i]);            /* :205-211 */
}
void orc_red_scalar_mul_reduce_array(int32_t *a, uint32_t n, int32_t c) {     /* :217-223 */
  for (uint32_t i = 0; i < n; i++) a[i] = orc_mul_red(a[i], c);
}

static int32_t centre(uint32_t x) { return x > (RQ - 1) / 2 ? (int32_t)x - RQ : (int32_t)x; }

/* R/NTT-RED/ntt_red256_tables.c: each table = the NTT/ table times 3^-1, centred;
 * scaled_inv_psi_powers carries 3^-8 (cancels the eight red()s of the product), _var 3^-6 */
int orc_red_make_table(int kind, uint32_t n, uint32_t psi, int32_t *out) {
  uint32_t *tmp = (uint32_t *)malloc(sizeof(uint32_t) * n);
  if (!tmp) return -2;
  uint32_t inv3 = orc_invmod(3, RQ);
  uint32_t scale = inv3;
  int base = kind;
  if (kind == ORC_SCALED_INV_PSI_POWERS) scale = orc_powmod(inv3, 8, RQ);
  if (kind == ORC_RED_SCALED_INV_PSI_POWERS_VAR) { scale = orc_powmod(inv3, 6, RQ); base = ORC_SCALED_INV_PSI_POWERS; }
  if (orc_make_table(base, n, RQ, psi, tmp) != 0) { free(tmp); return -1; }
  int level = (base >= ORC_OMEGA_POWERS && base <= ORC_INV_MIXED_POWERS_REV);
  for (uint32_t i = 0; i < n; i++)
    out[i] = (level && i == 0) ? 0 : centre((uint32_t)((uint64_t)tmp[i] * scale % RQ));
  free(tmp);
  return 0;
}

/* lazy butterflies: no reduction on the add/sub path (ntt_red.c:262-264, 432-435) */
#define RCT_BFLY(lo, hi, w, mul)                                   \
  do {                                                             \
    int32_t x_ = (mul) ? orc_mul_red(a[hi], (w)) : a[hi];          \
    int32_t u_ = a[lo];                                            \
    a[hi] = u_ - x_;                                               \
    a[lo] = u_ + x_;                                               \
  } while (0)
#define RGS_BFLY(lo, hi, w, mul)                                   \
  do {                                                             \
    int32_t x_ = a[hi];                                            \
    int32_t u_ = a[lo];                                            \
    a[hi] = (mul) ? orc_mul_red(u_ - x_, (w)) : u_ - x_;           \
    a[lo] = u_ + x_;                                               \
  } while (0)

/* Unlike the canonical path, skipping the j=0 multiply is observable here (mul_red
 * changes the representative), so `peel` states which entry point peels it. */
static void red_ct_rev2std(int32_t *a, uint32_t n, const int32_t *p, int peel) {
  for (uint32_t t = 1; t < n; t <<= 1)
    for (uint32_t j = 0; j < t; j++)
      for (uint32_t s = j; s < n; s += 2 * t) RCT_BFLY(s, s + t, p[t + j], !(peel && j == 0));
}
static void red_ct_std2rev(int32_t *a, uint32_t n, const int32_t *p, int peel) {
  uint32_t d = n;
  for (uint32_t t = 1; t < n; t <<= 1) {
    d >>= 1;
    for (uint32_t j = 0, u = 0; j < t; j++, u += 2 * d)
      for (uint32_t s = u; s < u + d; s++) RCT_BFLY(s, s + d, p[t + j], !(peel && j == 0));
  }
}
static void red_gs_rev2std(int32_t *a, uint32_t n, const int32_t *p, int peel) {
  uint32_t t = n;
  for (uint32_t d = 1; d < n; d <<= 1) {
    t >>= 1;
    for (uint32_t j = 0, u = 0; j < t; j++, u += 2 * d)
      for (uint32_t s = u; s < u + d; s++) RGS_BFLY(s, s + d, p[t + j], !(peel && j == 0));
  }
}
static void red_gs_std2rev(int32_t *a, uint32_t n, const int32_t *p, int peel) {
  for (uint32_t t = n >> 1; t > 0; t >>= 1)
    for (uint32_t j = 0; j < t; j++)
      for (uint32_t s = j; s < n; s += 2 * t) RGS_BFLY(s, s + t, p[t + j], !(peel && j == 0));
}
void orc_red_ct_rev2std(int32_t *a, uint32_t n, const int32_t *p)        { red_ct_rev2std(a, n, p, 1); }
void orc_red_mulntt_ct_rev2std(int32_t *a, uint32_t n, const int32_t *p) { red_ct_rev2std(a, n, p, 0); }
void orc_red_ct_std2rev(int32_t *a, uint32_t n, const int32_t *p)        { red_ct_std2rev(a, n, p, 1); }
void orc_red_mulntt_ct_std2rev(int32_t *a, uint32_t n, const int32_t *p) { red_ct_std2rev(a, n, p, 0); }
void orc_red_gs_rev2std(int32_t *a, uint32_t n, const int32_t *p)        { red_gs_rev2std(a, n, p, 1); }
void orc_red_nttmul_gs_rev2std(int32_t *a, uint32_t n, const int32_t *p) { red_gs_rev2std(a, n, p, 0); }
void orc_red_gs_std2rev(int32_t *a, uint32_t n, const int32_t *p)        { red_gs_std2rev(a, n, p, 1); }
void orc_red_nttmul_gs_std2rev(int32_t *a, uint32_t n, const int32_t *p) { red_gs_std2rev(a, n, p, 0); }

/* R/NTT-RED/ntt_red256.C:5-27 (variant 1, CT) and :30-52 (variant 4, GS) */
int orc_red_product(uint32_t n, uint32_t psi, int variant, int32_t *c, int32_t *a, int32_t *b) {
  if (variant != 1 && variant != 4) return -1;
  int32_t *tab = (int32_t *)malloc(sizeof(int32_t) * n * 4);
  if (!tab) return -2;
  int32_t *psi_p = tab, *fwd = tab + n, *inv = tab + 2 * n, *fin = tab + 3 * n;
  orc_red_make_table(ORC_PSI_POWERS, n, psi, psi_p);
  orc_red_make_table(variant == 1 ? ORC_OMEGA_POWERS_REV : ORC_OMEGA_POWERS, n, psi, fwd);
  orc_red_make_table(variant == 1 ? ORC_INV_OMEGA_POWERS : ORC_INV_OMEGA_POWERS_REV, n, psi, inv);
  orc_red_make_table(ORC_SCALED_INV_PSI_POWERS, n, psi, fin);
  int32_t *ops[2] = {a, b};
  for (int k = 0; k < 2; k++) {
    orc_red_shift_array(ops[k], n);
    orc_red_mul_reduce_array_tab(ops[k], n, psi_p);
    if (variant == 1) orc_red_ct_std2rev(ops[k], n, fwd); else orc_red_gs_std2rev(ops[k], n, fwd);
    orc_red_reduce_array(ops[k], n);
  }
  orc_red_mul_reduce_array(c, n, a, b);
  orc_red_reduce_array_twice(c, n);
  if (variant == 1) orc_red_ct_rev2std(c, n, inv); else orc_red_gs_rev2std(c, n, inv);
  orc_red_mul_reduce_array_tab(c, n, fin);
  orc_red_reduce_array_twice(c, n);
  orc_red_correct(c, n);
  free(tab);
  return 0;
}

/* ------------------------------------------------------------------------- */
/* synthetic inputs: splitmix64, coefficient = next() % q (the reference uses   */
/* rand() % Q: time_testing256.c:97-98, Generator_Params/generate_coeff.c:47)   */
/* ------------------------------------------------------------------------- */
void orc_fill_random(int32_t *dst, size_t count, uint32_t q, uint64_t seed) {
  uint64_t s = seed;
  for (size_t i = 0; i < count; i++) {
    s += 0x9E3779B97F4A7C15ull;
    uint64_t z = s;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    dst[i] = (int32_t)(z % q);
  }
}
