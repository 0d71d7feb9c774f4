/*
 * nttb200_batch_bench.c -- the batch product call from a plain C host program: what a maintainer
 * of the reference would write around nttb200_polymul_batch(), timed
 * end to end with host buffers (int32 in, int32 out), pinned and pageable.
 *
 *   gcc -O2 -Iinclude tools/nttb200_batch_bench.c -Lntt-based-polynomial-multiplier-fpga_b200 -lnttb200 \
 *       -Wl,-rpath,$PWD/ntt-based-polynomial-multiplier-fpga_b200 -o nttb200_batch_bench
 *   ./nttb200_batch_bench [n q log2_batch calls]          (default 256 12289 16 20)
 */
#define _POSIX_C_SOURCE 200809L
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "nttb200.h"

static double now(void) {
  struct timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  return t.tv_sec + 1e-9 * t.tv_nsec;
}
static void die(const char *what) {
  fprintf(stderr, "%s: %s\n", what, nttb200_last_error());
  exit(1);
}

static void run(nttb200_plan *plan, int32_t *c, const int32_t *a, const int32_t *b, size_t batch, int calls,
                const char *what) {
  for (int i = 0; i < 3; i++)
    if (nttb200_polymul_batch(plan, c, a, b, batch)) die("nttb200_polymul_batch");
  const double t0 = now();
  for (int i = 0; i < calls; i++)
    if (nttb200_polymul_batch(plan, c, a, b, batch)) die("nttb200_polymul_batch");
  const double dt = now() - t0;
  unsigned long long r16 = 0, r32 = 0, c16 = 0;
  int threads = 0;
  nttb200_plan_wire_stats(plan, &r16, &r32, &c16, &threads);
  printf("%-9s %8.2f M polymul/s  (%.3f ms per call of %zu; rows on the 16-bit wire %llu, 32-bit %llu, host threads %d)\n",
         what, batch * (double)calls / dt / 1e6, 1e3 * dt / calls, batch, r16, r32, threads);
}

int main(int argc, char **argv) {
  const uint32_t n = argc > 1 ? (uint32_t)atoi(argv[1]) : 256;
  const uint32_t q = argc > 2 ? (uint32_t)strtoul(argv[2], 0, 10) : 12289;
  const size_t batch = (size_t)1 << (argc > 3 ? atoi(argv[3]) : 16);
  const int calls = argc > 4 ? atoi(argv[4]) : 20;
  const size_t bytes = batch * n * sizeof(int32_t);
  nttb200_plan *plan;
  if (nttb200_plan_create(&plan, n, q, 0, NTTB200_PLAN_DEFAULT)) die("nttb200_plan_create");
  printf("%s\n", nttb200_plan_describe(plan));

  int32_t *pin[3], *pag[3];
  for (int i = 0; i < 3; i++) {
    pin[i] = nttb200_host_alloc(bytes);
    pag[i] = malloc(bytes);
    if (!pin[i] || !pag[i]) die("allocation");
  }
  unsigned long long s = 88172645463325252ull;
  for (size_t i = 0; i < batch * n; i++) {
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    pin[0][i] = (int32_t)(s % q);
    pin[1][i] = (int32_t)((s >> 32) % q);
  }
  memcpy(pag[0], pin[0], bytes);
  memcpy(pag[1], pin[1], bytes);
  memset(pag[2], 0xff, bytes);

  run(plan, pin[2], pin[0], pin[1], batch, calls, "pinned");
  run(plan, pag[2], pag[0], pag[1], batch, calls, "pageable");
  if (memcmp(pin[2], pag[2], bytes)) { fprintf(stderr, "pinned and pageable results differ\n"); return 1; }
  /* x^0 * b = b: a row check that needs no reference */
  memset(pin[0], 0, (size_t)n * sizeof(int32_t));
  pin[0][0] = 1;
  if (nttb200_polymul_batch(plan, pin[2], pin[0], pin[1], batch)) die("nttb200_polymul_batch");
  if (memcmp(pin[2], pin[1], (size_t)n * sizeof(int32_t))) { fprintf(stderr, "1 * b != b\n"); return 1; }
  printf("checks ok (pinned == pageable, 1 * b == b)\n");
  for (int i = 0; i < 3; i++) { nttb200_host_free(pin[i]); free(pag[i]); }
  nttb200_plan_destroy(plan);
  return 0;
}
