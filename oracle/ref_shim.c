/*
 * ref_shim.c -- thin export layer linked WITH the unmodified reference sources into
 * oracle/_ref/libntt_ref*.so (see oracle/Makefile).  TEST INFRASTRUCTURE ONLY.
 *
 * The reference's n=256 wrappers are `static inline` in its headers
 * (NTT/ntt256.h:20-69, NTT-RED/ntt_red256.h:21-70), so a shared object built from
 * the reference .C files alone would not export them; this file includes the
 * reference headers where they lie (-I into /root/reference, nothing is copied) and
 * re-exports each wrapper, the four products, the tables and a timing loop that
 * follows the reference's own method (time_testing256.c:175-185).
 */
#define _POSIX_C_SOURCE 199309L
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <time.h>

#include "NTT/ntt256.h"
#include "NTT-RED/ntt_red256.h"

/* products: 1 = ntt256_product1, 4 = ntt256_product4, 101 = ntt_red256_product1,
 * 104 = ntt_red256_product4, 10 = merged CT-fwd/GS-inv pipeline, 110 = its RED twin */
int ref_product(int variant, int32_t *c, int32_t *a, int32_t *b) {
  switch (variant) {
    case 1:   ntt256_product1(c, a, b); return 0;
    case 4:   ntt256_product4(c, a, b); return 0;
    case 101: ntt_red256_product1(c, a, b); return 0;
    case 104: ntt_red256_product4(c, a, b); return 0;
    case 10:
      mulntt256_ct_std2rev(a);
      mulntt256_ct_std2rev(b);
      mul_array(c, 256, a, b);
      inttmul256_gs_rev2std(c);
      scalar_mul_array(c, 256, ntt256_inv_n);
      return 0;
    case 110:
      shift_array(a, 256); mulntt_red256_ct_std2rev(a); reduce_array(a, 256);
      shift_array(b, 256); mulntt_red256_ct_std2rev(b); reduce_array(b, 256);
      mul_reduce_array(c, 256, a, b);
      reduce_array_twice(c, 256);
      inttmul_red256_gs_rev2std(c);
      scalar_mul_reduce_array(c, 256, ntt_red256_rescale8);
      reduce_array_twice(c, 256);
      correct(c, 256);
      return 0;
    default: return -1;
  }
}

int ref_product_batch(int variant, int32_t *c, const int32_t *a, const int32_t *b, size_t batch) {
  int32_t ta[256], tb[256];
  for (size_t r = 0; r < batch; r++) {
    memcpy(ta, a + 256 * r, sizeof ta);
    memcpy(tb, b + 256 * r, sizeof tb);
    if (ref_product(variant, c + 256 * r, ta, tb)) return -1;
  }
  return 0;
}

/* standalone n=256 transforms, by name id */
int ref_transform(int id, int32_t *a) {
  switch (id) {
    case 0:  ntt256_ct_rev2std(a); return 0;
    case 1:  ntt256_gs_rev2std(a); return 0;
    case 2:  ntt256_ct_std2rev(a); return 0;
    case 3:  ntt256_gs_std2rev(a); return 0;
    case 4:  intt256_ct_rev2std(a); return 0;
    case 5:  intt256_gs_rev2std(a); return 0;
    case 6:  intt256_ct_std2rev(a); return 0;
    case 7:  intt256_gs_std2rev(a); return 0;
    case 8:  mulntt256_ct_rev2std(a); return 0;
    case 9:  mulntt256_ct_std2rev(a); return 0;
    case 10: inttmul256_gs_rev2std(a); return 0;
    case 11: inttmul256_gs_std2rev(a); return 0;
    case 12: ntt_ct_rev2std_v1(a, 256, ntt256_psi_powers); return 0;
    /* RED twins (unreduced int32 outputs) */
    case 100: ntt_red256_ct_rev2std(a); return 0;
    case 101: ntt_red256_gs_rev2std(a); return 0;
    case 102: ntt_red256_ct_std2rev(a); return 0;
    case 103: ntt_red256_gs_std2rev(a); return 0;
    case 104: intt_red256_ct_rev2std(a); return 0;
    case 105: intt_red256_gs_rev2std(a); return 0;
    case 106: intt_red256_ct_std2rev(a); return 0;
    case 107: intt_red256_gs_std2rev(a); return 0;
    case 108: mulntt_red256_ct_rev2std(a); return 0;
    case 109: mulntt_red256_ct_std2rev(a); return 0;
    case 110: inttmul_red256_gs_rev2std(a); return 0;
    case 111: inttmul_red256_gs_std2rev(a); return 0;
    default: return -1;
  }
}

/* the generic entry points with the CALLER's table (ntt.h:71-183, ntt_red.h:174-284): what pins
 * the j = 0 peel of the un-merged functions, which never read p[t] (ntt.C:313-317 ...) */
int ref_transform_tab(int id, int32_t *a, uint32_t n, const uint16_t *p) {
  switch (id) {
    case 0: ntt_ct_rev2std_v1(a, n, p); return 0;
    case 1: ntt_ct_rev2std(a, n, p); return 0;
    case 2: mulntt_ct_rev2std(a, n, p); return 0;
    case 3: ntt_ct_std2rev(a, n, p); return 0;
    case 4: mulntt_ct_std2rev(a, n, p); return 0;
    case 5: ntt_gs_rev2std(a, n, p); return 0;
    case 6: nttmul_gs_rev2std(a, n, p); return 0;
    case 7: ntt_gs_std2rev(a, n, p); return 0;
    case 8: nttmul_gs_std2rev(a, n, p); return 0;
    default: return -1;
  }
}
int ref_red_transform_tab(int id, int32_t *a, uint32_t n, const int16_t *p) {
  switch (id) {
    case 1: ntt_red_ct_rev2std(a, n, p); return 0;
    case 2: mulntt_red_ct_rev2std(a, n, p); return 0;
    case 3: ntt_red_ct_std2rev(a, n, p); return 0;
    case 4: mulntt_red_ct_std2rev(a, n, p); return 0;
    case 5: ntt_red_gs_rev2std(a, n, p); return 0;
    case 6: nttmul_red_gs_rev2std(a, n, p); return 0;
    case 7: ntt_red_gs_std2rev(a, n, p); return 0;
    case 8: nttmul_red_gs_std2rev(a, n, p); return 0;
    default: return -1;
  }
}

/* tables, by the oracle's enum order (ntt_oracle.h); red != 0 selects the int16 set */
const void *ref_table(int kind, int red) {
  if (!red) switch (kind) {
    case 0: return ntt256_psi_powers;
    case 1: return ntt256_inv_psi_powers;
    case 2: return ntt256_scaled_inv_psi_powers;
    case 3: return ntt256_omega_powers;
    case 4: return ntt256_omega_powers_rev;
    case 5: return ntt256_inv_omega_powers;
    case 6: return ntt256_inv_omega_powers_rev;
    case 7: return ntt256_mixed_powers;
    case 8: return ntt256_mixed_powers_rev;
    case 9: return ntt256_inv_mixed_powers;
    case 10: return ntt256_inv_mixed_powers_rev;
    default: return 0;
  }
  switch (kind) {
    case 0: return ntt_red256_psi_powers;
    case 1: return ntt_red256_inv_psi_powers;
    case 2: return ntt_red256_scaled_inv_psi_powers;
    case 3: return ntt_red256_omega_powers;
    case 4: return ntt_red256_omega_powers_rev;
    case 5: return ntt_red256_inv_omega_powers;
    case 6: return ntt_red256_inv_omega_powers_rev;
    case 7: return ntt_red256_mixed_powers;
    case 8: return ntt_red256_mixed_powers_rev;
    case 9: return ntt_red256_inv_mixed_powers;
    case 10: return ntt_red256_inv_mixed_powers_rev;
    case 100: return ntt_red256_scaled_inv_psi_powers_var;
    default: return 0;
  }
}

int ref_param(int which) {
  switch (which) {
    case 0: return ntt256_psi;      case 1: return ntt256_omega;
    case 2: return ntt256_inv_psi;  case 3: return ntt256_inv_omega;
    case 4: return ntt256_inv_n;    case 5: return ntt_red256_inv_k;
    case 6: return ntt_red256_rescale8; case 7: return ntt_red256_rescale6;
    default: return -1;
  }
}

/* elementwise / range helpers, re-exported under ref_ names for the table-driven tests */
void ref_mul_array16(int32_t *a, uint32_t n, const uint16_t *p) { mul_array16(a, n, p); }
void ref_mul_array(int32_t *c, uint32_t n, const int32_t *a, const int32_t *b) { mul_array(c, n, a, b); }
void ref_scalar_mul_array(int32_t *a, uint32_t n, int32_t c) { scalar_mul_array(a, n, c); }
void ref_bitrev_shuffle(int32_t *a, uint32_t n) { bitrev_shuffle(a, n); }
void ref_red_helper(int which, int32_t *a, uint32_t n) {
  switch (which) {
    case 0: shift_array(a, n); break;
    case 1: reduce_array(a, n); break;
    case 2: reduce_array_twice(a, n); break;
    case 3: correct(a, n); break;
    case 4: normalize(a, n); break;
    case 5: normalize_inv3(a, n); break;
  }
}

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/* polymul/s of this thread; operand restore excluded like time_testing256.c:175-185 */
double ref_bench_loop(int variant, const int32_t *a, const int32_t *b, size_t batch,
                      double min_seconds, uint64_t *calls_out) {
  enum { CH = 64 };                       /* restore CH operand pairs, then time CH calls */
  static __thread int32_t ta[CH * 256], tb[CH * 256], tc[256];
  double busy = 0.0, t_begin = now_s();
  uint64_t calls = 0;
  volatile int32_t sink = 0;
  while (now_s() - t_begin < min_seconds) {
    for (size_t r = 0; r < batch; r += CH) {
      size_t m = batch - r < CH ? batch - r : CH;
      memcpy(ta, a + 256 * r, m * 256 * sizeof(int32_t));
      memcpy(tb, b + 256 * r, m * 256 * sizeof(int32_t));
      double t0 = now_s();
      for (size_t k = 0; k < m; k++) {
        ref_product(variant, tc, ta + 256 * k, tb + 256 * k);
        sink ^= tc[0];
      }
      busy += now_s() - t0;
      calls += m;
    }
  }
  (void)sink;
  if (calls_out) *calls_out = calls;
  return busy > 0 ? calls / busy : 0.0;
}
