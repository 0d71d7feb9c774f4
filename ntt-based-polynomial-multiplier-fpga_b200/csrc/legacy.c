/*
 * legacy.c -- the reference's C call surface (include/nttb200_legacy.h) as thin host
 * wrappers over the nttb200_* C ABI: batch = 1 on a lazily created default plan
 * (n = 256, q = 12289, psi = 1002 -- the parameters of R/NTT/ntt256_tables.h:20-24).
 */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>

#include "nttb200.h"
#include "nttb200_legacy.h"

#define LEGACY_Q 12289u

static nttb200_plan *g_plan256;
static pthread_once_t g_once = PTHREAD_ONCE_INIT;
static int g_plan_rc;
static int g_clobber;

static void die(const char *what) {
  fprintf(stderr, "nttb200 legacy surface: %s failed: %s\n", what, nttb200_last_error());
  abort();
}

static void make_default_plan(void) { g_plan_rc = nttb200_plan_create(&g_plan256, 256, LEGACY_Q, 1002, 0); }

static nttb200_plan *default_plan(void) {
  pthread_once(&g_once, make_default_plan);
  if (g_plan_rc != 0 || !g_plan256) die("nttb200_plan_create(256, 12289, 1002)");
  return g_plan256;
}

void nttb200_legacy_set_clobber(int on) { g_clobber = on; }

/* Post-state of an operand of the optimized products ("a and b are modified",
 * R/NTT-RED/ntt_red256.h:77-86): shift -> x psi^i (mul_red) -> forward transform -> reduce,
 * R/NTT-RED/ntt_red256.C:5-13 (CT, product1) and :31-39 (GS, product4) -- an UNREDUCED
 * representative, reproduced with the exact Longa-Naehrig pieces below. */
static void red_post_state(int32_t *x, int gs) {
  shift_array(x, 256);
  mul_reduce_array16(x, 256, ntt_red256_psi_powers);
  if (gs) ntt_red_gs_std2rev(x, 256, ntt_red256_omega_powers);
  else ntt_red_ct_std2rev(x, 256, ntt_red256_omega_powers_rev);
  reduce_array(x, 256);
}

/* clobber: 0 none, 1 the canonical post-state of ntt256_product1/4 (psi-twisted NTT, bit-reversed
 * order: R/NTT/ntt256.C:6-9), 2 / 3 the post-state of ntt_red256_product1 / 4 */
static void product(int32_t *c, int32_t *a, int32_t *b, int clobber) {
  nttb200_plan *P = default_plan();
  if (nttb200_polymul_batch(P, c, a, b, 1) != 0) die("nttb200_polymul_batch");
  if (clobber == 1) {
    if (nttb200_ntt_batch(P, NTTB200_MULNTT_STD2REV, a, 1) != 0) die("nttb200_ntt_batch");
    if (nttb200_ntt_batch(P, NTTB200_MULNTT_STD2REV, b, 1) != 0) die("nttb200_ntt_batch");
  } else if (clobber >= 2) {
    red_post_state(a, clobber == 3);
    red_post_state(b, clobber == 3);
  }
}

void ntt256_product1(int32_t *c, int32_t *a, int32_t *b) { product(c, a, b, g_clobber ? 1 : 0); }
void ntt256_product4(int32_t *c, int32_t *a, int32_t *b) { product(c, a, b, g_clobber ? 1 : 0); }
void ntt_red256_product1(int32_t *c, int32_t *a, int32_t *b) { product(c, a, b, g_clobber ? 2 : 0); }
void ntt_red256_product4(int32_t *c, int32_t *a, int32_t *b) { product(c, a, b, g_clobber ? 3 : 0); }

/* caller's 16-bit table -> the ABI's 32-bit table */
static void table_transform(int32_t *a, uint32_t n, const uint16_t *p, int dataflow, int skip_j0) {
  uint32_t *p32 = (uint32_t *)malloc(sizeof(uint32_t) * (n ? n : 1));
  if (!p32) die("malloc");
  for (uint32_t i = 0; i < n; i++) p32[i] = p[i];
  int rc = nttb200_ntt_table_batch(n, LEGACY_Q, dataflow, skip_j0, p32, a, 1);
  free(p32);
  if (rc != 0) die("nttb200_ntt_table_batch");
}

void ntt_ct_rev2std_v1(int32_t *a, uint32_t n, const uint16_t *p) {
  /* v1 reads w_t^j from the psi-power table as p[j * (n/t)] (R/NTT/ntt.C:186): rebuild
   * the level layout p'[t+j] the other entry points use, then run the same dataflow */
  uint32_t *lvl = (uint32_t *)calloc(n ? n : 1, sizeof(uint32_t));
  if (!lvl) die("calloc");
  for (uint32_t t = 1, l = n; t < n; t <<= 1, l >>= 1)
    for (uint32_t j = 1; j < t; j++) lvl[t + j] = p[j * l];   /* j = 0 is peeled: p[0] is never read */
  int rc = nttb200_ntt_table_batch(n, LEGACY_Q, NTTB200_DF_CT_REV2STD, 1, lvl, a, 1);
  free(lvl);
  if (rc != 0) die("nttb200_ntt_table_batch");
}
/* The un-merged entry points peel the j = 0 block: its butterflies run without a multiplication
 * and p[t] is never read (R/NTT/ntt.C:226-231, 313-317, 401-405, 477-481) -- skip_j0 = 1; the
 * psi-merged ones multiply every block by the caller's p[t+j]. */
void ntt_ct_rev2std(int32_t *a, uint32_t n, const uint16_t *p)    { table_transform(a, n, p, NTTB200_DF_CT_REV2STD, 1); }
void mulntt_ct_rev2std(int32_t *a, uint32_t n, const uint16_t *p) { table_transform(a, n, p, NTTB200_DF_CT_REV2STD, 0); }
void ntt_ct_std2rev(int32_t *a, uint32_t n, const uint16_t *p)    { table_transform(a, n, p, NTTB200_DF_CT_STD2REV, 1); }
void mulntt_ct_std2rev(int32_t *a, uint32_t n, const uint16_t *p) { table_transform(a, n, p, NTTB200_DF_CT_STD2REV, 0); }
void ntt_gs_rev2std(int32_t *a, uint32_t n, const uint16_t *p)    { table_transform(a, n, p, NTTB200_DF_GS_REV2STD, 1); }
void nttmul_gs_rev2std(int32_t *a, uint32_t n, const uint16_t *p) { table_transform(a, n, p, NTTB200_DF_GS_REV2STD, 0); }
void ntt_gs_std2rev(int32_t *a, uint32_t n, const uint16_t *p)    { table_transform(a, n, p, NTTB200_DF_GS_STD2REV, 1); }
void nttmul_gs_std2rev(int32_t *a, uint32_t n, const uint16_t *p) { table_transform(a, n, p, NTTB200_DF_GS_STD2REV, 0); }

/* elementwise: a length-n array is one row of an n-coefficient "plan"; these ops do not
 * use the plan's roots, only its modulus, so any n is served through the 256-plan by
 * treating the array as ceil(n/256) rows when n is a multiple of 256, else element-wise
 * through a scratch row. */
static void elementwise(int32_t *c, uint32_t n, const int32_t *a, const int32_t *b) {
  nttb200_plan *P = default_plan();
  uint32_t rows = (n + 255) / 256;
  int32_t *ta = (int32_t *)calloc((size_t)rows * 256 * 3, sizeof(int32_t));
  if (!ta) die("calloc");
  int32_t *tb = ta + (size_t)rows * 256, *tc = tb + (size_t)rows * 256;
  for (uint32_t i = 0; i < n; i++) { ta[i] = a[i]; tb[i] = b[i]; }
  int rc = nttb200_mul_array_batch(P, tc, ta, tb, rows);
  for (uint32_t i = 0; i < n && rc == 0; i++) c[i] = tc[i];
  free(ta);
  if (rc != 0) die("nttb200_mul_array_batch");
}

void mul_array(int32_t *c, uint32_t n, const int32_t *a, const int32_t *b) { elementwise(c, n, a, b); }

void mul_array16(int32_t *a, uint32_t n, const uint16_t *p) {
  int32_t *p32 = (int32_t *)malloc(sizeof(int32_t) * (n ? n : 1));
  if (!p32) die("malloc");
  for (uint32_t i = 0; i < n; i++) p32[i] = p[i];
  elementwise(a, n, a, p32);
  free(p32);
}

void scalar_mul_array(int32_t *a, uint32_t n, int32_t c) {
  int32_t *cv = (int32_t *)malloc(sizeof(int32_t) * (n ? n : 1));
  if (!cv) die("malloc");
  for (uint32_t i = 0; i < n; i++) cv[i] = c;
  elementwise(a, n, a, cv);
  free(cv);
}

/* ---- permutations ------------------------------------------------------------------------ */
void bitrev_shuffle(int32_t *a, uint32_t n) {
  if (nttb200_bitrev_shuffle_batch(a, n, 1) != 0) die("nttb200_bitrev_shuffle_batch");
}
void shuffle_with_table(int32_t *a, const uint16_t p[][2], uint32_t n) {
  /* the reference does not pass the array length: it is the largest index in the table + 1 */
  size_t words = 0;
  for (uint32_t i = 0; i < n; i++) {
    if ((size_t)p[i][0] + 1 > words) words = (size_t)p[i][0] + 1;
    if ((size_t)p[i][1] + 1 > words) words = (size_t)p[i][1] + 1;
  }
  if (words && nttb200_shuffle_with_table(a, words, &p[0][0], n) != 0) die("nttb200_shuffle_with_table");
}

/* ---- Longa-Naehrig surface, exact ------------------------------------------------------------ */
static void red_op(int op, int32_t *c, const int32_t *a, const int32_t *b, int32_t sc, uint32_t n) {
  if (nttb200_red_elementwise_batch(op, c, a, b, sc, n) != 0) die("nttb200_red_elementwise_batch");
}
void normalize(int32_t *a, uint32_t n) { red_op(NTTB200_RED_NORMALIZE, a, a, NULL, 0, n); }
void normalize_inv3(int32_t *a, uint32_t n) { red_op(NTTB200_RED_NORMALIZE_INV3, a, a, NULL, 0, n); }
void shift_array(int32_t *a, uint32_t n) { red_op(NTTB200_RED_SHIFT, a, a, NULL, 0, n); }
void reduce_array(int32_t *a, uint32_t n) { red_op(NTTB200_RED_REDUCE, a, a, NULL, 0, n); }
void reduce_array_twice(int32_t *a, uint32_t n) { red_op(NTTB200_RED_REDUCE_TWICE, a, a, NULL, 0, n); }
void correct(int32_t *a, uint32_t n) { red_op(NTTB200_RED_CORRECT, a, a, NULL, 0, n); }
void mul_reduce_array(int32_t *c, uint32_t n, const int32_t *a, const int32_t *b) {
  red_op(NTTB200_RED_MUL_RED, c, a, b, 0, n);
}
void scalar_mul_reduce_array(int32_t *a, uint32_t n, int32_t c) { red_op(NTTB200_RED_SCALAR_MUL_RED, a, a, NULL, c, n); }
void mul_reduce_array16(int32_t *a, uint32_t n, const int16_t *p) {
  int32_t *p32 = (int32_t *)malloc(sizeof(int32_t) * (n ? n : 1));
  if (!p32) die("malloc");
  for (uint32_t i = 0; i < n; i++) p32[i] = p[i];
  red_op(NTTB200_RED_MUL_RED, a, a, p32, 0, n);
  free(p32);
}
static void red_transform(int32_t *a, uint32_t n, const int16_t *p, int dataflow, int skip_j0) {
  int32_t *p32 = (int32_t *)malloc(sizeof(int32_t) * (n ? n : 1));
  if (!p32) die("malloc");
  for (uint32_t i = 0; i < n; i++) p32[i] = p[i];
  int rc = nttb200_red_ntt_table_batch(n, dataflow, skip_j0, p32, a, 1);
  free(p32);
  if (rc != 0) die("nttb200_red_ntt_table_batch");
}
void ntt_red_ct_rev2std(int32_t *a, uint32_t n, const int16_t *p)    { red_transform(a, n, p, NTTB200_DF_CT_REV2STD, 1); }
void mulntt_red_ct_rev2std(int32_t *a, uint32_t n, const int16_t *p) { red_transform(a, n, p, NTTB200_DF_CT_REV2STD, 0); }
void ntt_red_ct_std2rev(int32_t *a, uint32_t n, const int16_t *p)    { red_transform(a, n, p, NTTB200_DF_CT_STD2REV, 1); }
void mulntt_red_ct_std2rev(int32_t *a, uint32_t n, const int16_t *p) { red_transform(a, n, p, NTTB200_DF_CT_STD2REV, 0); }
void ntt_red_gs_rev2std(int32_t *a, uint32_t n, const int16_t *p)    { red_transform(a, n, p, NTTB200_DF_GS_REV2STD, 1); }
void nttmul_red_gs_rev2std(int32_t *a, uint32_t n, const int16_t *p) { red_transform(a, n, p, NTTB200_DF_GS_REV2STD, 0); }
void ntt_red_gs_std2rev(int32_t *a, uint32_t n, const int16_t *p)    { red_transform(a, n, p, NTTB200_DF_GS_STD2REV, 1); }
void nttmul_red_gs_std2rev(int32_t *a, uint32_t n, const int16_t *p) { red_transform(a, n, p, NTTB200_DF_GS_STD2REV, 0); }
