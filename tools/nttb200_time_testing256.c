/*
 * nttb200_time_testing256 -- command-line twin of the reference's benchmark driver
 *   Multiplier_NTT_Based/NTT_Software/NTT_Software_Evaluations/NTT-256/time_testing256.c:118-245
 * on the GPU library: reads coeficientes_a.txt / coeficientes_b.txt (decimals separated by blanks,
 * time_testing256.c:17-44; written by Generator_Params/generate_coeff.c:35-59), multiplies them
 * 30 times with ntt256_product4 under clock_gettime, prints the mean and C in the reference's
 * layout (time_testing256.c:46-64, 235-236), so that stdout diffs against the reference
 * binary's except for the measured time.
 *
 *   nttb200_time_testing256 [a.txt b.txt] [--variant 1|4|red1|red4] [--batch B]
 * --batch B additionally times B copies through nttb200_polymul_batch (the way the GPU is
 * meant to be used) and reports polymul/s on stderr.
 */
#define _POSIX_C_SOURCE 199309L
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "nttb200.h"
#include "nttb200_gen.h"
#include "nttb200_legacy.h"

#define N 256

static double now_s(void) {
  struct timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  return t.tv_sec + t.tv_nsec / 1e9;
}

int main(int argc, char **argv) {
  const char *fa = "coeficientes_a.txt", *fb = "coeficientes_b.txt", *variant = "4";
  long batch = 0;
  int pos = 0;
  for (int i = 1; i < argc; i++) {
    if (!strcmp(argv[i], "--variant") && i + 1 < argc) variant = argv[++i];
    else if (!strcmp(argv[i], "--batch") && i + 1 < argc) batch = atol(argv[++i]);
    else if (pos == 0) { fa = argv[i]; pos++; }
    else if (pos == 1) { fb = argv[i]; pos++; }
  }
  void (*product)(int32_t *, int32_t *, int32_t *) = ntt256_product4;
  const char *banner = "Executando mult ntt256 GS(C, A, B)...\n\n", *label = "ntt256 gs";
  if (!strcmp(variant, "1")) { product = ntt256_product1; banner = "Executando mult ntt256  CT (C, A, B)...\n\n"; label = "ntt256 ct"; }
  else if (!strcmp(variant, "red1")) { product = ntt_red256_product1; banner = "Executando mult ntt256 CT com Redução(C, A, B)...\n\n"; label = "ntt256 ct com red"; }
  else if (!strcmp(variant, "red4")) { product = ntt_red256_product4; banner = "Executando mult ntt256 GS com Redução(C, A, B)...\n\n"; label = "ntt256 gs com red"; }

  int32_t *a = calloc(N, sizeof *a), *b = calloc(N, sizeof *b), *c = calloc(N, sizeof *c);
  int32_t *a0 = calloc(N, sizeof *a0), *b0 = calloc(N, sizeof *b0);
  if (!a || !b || !c || !a0 || !b0) { fprintf(stderr, "Falha na alocação de memória\n"); return 1; }
  printf("Lendo valores a partir do arquivo txt\n");
  if (nttb200_read_coeff_file(fa, a0, N) < 0) { perror("Erro ao abrir o arquivo para leitura"); return 1; }
  printf("Lendo valores a partir do arquivo txt\n");
  if (nttb200_read_coeff_file(fb, b0, N) < 0) { perror("Erro ao abrir o arquivo para leitura"); return 1; }

  printf("%s", banner);
  /* one untimed call first: it creates the CUDA context, the plan and its tables (the reference
   * has no such one-time cost; without this the mean of 30 calls is that cost / 30) */
  memcpy(a, a0, N * sizeof *a);
  memcpy(b, b0, N * sizeof *b);
  product(c, a, b);
  double sum = 0;
  const int iters = 30;
  for (int it = 0; it < iters; it++) {
    memcpy(a, a0, N * sizeof *a);                 /* reset_polyABC: the reference clobbers a, b */
    memcpy(b, b0, N * sizeof *b);
    memset(c, 0, N * sizeof *c);
    double t0 = now_s();
    product(c, a, b);
    sum += now_s() - t0;
  }
  printf("Tempo total médio %s: %.3f ms\n", label, sum / iters * 1000);
  printf("Polinômio C (Resultado C = A * B):\n");
  nttb200_print_array(stdout, c, N);
  printf("\n");

  if (batch > 0) {
    nttb200_plan *plan = NULL;
    if (nttb200_plan_create(&plan, N, 12289, 1002, 0)) { fprintf(stderr, "%s\n", nttb200_last_error()); return 1; }
    const size_t bytes = (size_t)batch * N * sizeof(int32_t);
    int32_t *ba = nttb200_host_alloc(bytes), *bb = nttb200_host_alloc(bytes), *bc = nttb200_host_alloc(bytes);
    if (!ba || !bb || !bc) { fprintf(stderr, "%s\n", nttb200_last_error()); return 1; }
    for (long r = 0; r < batch; r++) { memcpy(ba + r * N, a0, N * 4); memcpy(bb + r * N, b0, N * 4); }
    nttb200_polymul_batch(plan, bc, ba, bb, (size_t)batch);      /* warm-up */
    double t0 = now_s();
    if (nttb200_polymul_batch(plan, bc, ba, bb, (size_t)batch)) { fprintf(stderr, "%s\n", nttb200_last_error()); return 1; }
    double dt = now_s() - t0;
    int same = !memcmp(bc, c, N * 4) && !memcmp(bc + (batch - 1) * N, c, N * 4);
    fprintf(stderr, "batch %ld through nttb200_polymul_batch: %.3f ms, %.3e polymul/s, rows %s the single result\n",
            batch, dt * 1000, batch / dt, same ? "equal" : "DIFFER FROM");
    nttb200_host_free(ba); nttb200_host_free(bb); nttb200_host_free(bc);
    nttb200_plan_destroy(plan);
    if (!same) return 2;
  }
  free(a); free(b); free(c); free(a0); free(b0);
  return 0;
}
