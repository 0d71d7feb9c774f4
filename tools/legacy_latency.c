/* per-call latency of the reference-named entry points of libnttb200.so (GPU box):
 *   gcc -O2 -Iinclude tools/legacy_latency.c -Lntt-based-polynomial-multiplier-fpga_b200 -lnttb200 -Wl,-rpath,$PWD/ntt-based-polynomial-multiplier-fpga_b200 */
#define _POSIX_C_SOURCE 199309L
#include <stdio.h>
#include <stdlib.h>
#include <time.h>
#include "nttb200.h"
#include "nttb200_legacy.h"
static double now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + t.tv_nsec * 1e-9; }
int main(void) {
  enum { N = 256, Q = 12289, K = 1000 };
  static int32_t a[N], b[N], c[N];
  static uint32_t t32[N];
  static uint16_t t16[N];
  static int16_t r16[N];
  for (int i = 0; i < N; i++) { a[i] = (i * 7919) % Q; b[i] = (i * 104729 + 5) % Q; }
  nttb200_make_table(NTTB200_OMEGA_POWERS_REV, N, Q, 1002, t32);
  for (int i = 0; i < N; i++) { t16[i] = (uint16_t)t32[i]; r16[i] = (int16_t)((t32[i] * 8193u) % Q > 6144 ? (int)((t32[i] * 8193u) % Q) - Q : (int)((t32[i] * 8193u) % Q)); }
  ntt256_product1(c, a, b);
  ntt_ct_std2rev(a, N, t16);
  double t0 = now();
  for (int k = 0; k < K; k++) ntt256_product1(c, a, b);
  double t1 = now();
  for (int k = 0; k < K; k++) { for (int i = 0; i < N; i++) a[i] = (a[i] & 0x1fff) % Q; ntt_ct_std2rev(a, N, t16); }
  double t2 = now();
  for (int k = 0; k < K; k++) { for (int i = 0; i < N; i++) a[i] = (a[i] % 6144); ntt_red_ct_std2rev(a, N, r16); }
  double t3 = now();
  for (int k = 0; k < K; k++) mul_array(c, N, a, b);
  double t4 = now();
  for (int k = 0; k < K; k++) reduce_array(a, N);
  double t5 = now();
  printf("ntt256_product1 %.1f us  ntt_ct_std2rev %.1f us  ntt_red_ct_std2rev %.1f us  mul_array %.1f us  reduce_array %.1f us\n",
         (t1 - t0) / K * 1e6, (t2 - t1) / K * 1e6, (t3 - t2) / K * 1e6, (t4 - t3) / K * 1e6, (t5 - t4) / K * 1e6);
  return 0;
}
