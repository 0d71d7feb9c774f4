/*
 * nttb200_generate_coeff -- command-line twin of the reference's coefficient generator
 * (Generator_Params/generate_coeff.c:12-59): writes coeficientes_a.txt and coeficientes_b.txt, N random
 * coefficients in [0, Q) each, decimals separated by blanks, 10 per line, with the reference's messages.
 * The reference is compiled for Q = 12289, N = 256 and seeds rand() with the time; here both are the
 * defaults and may be given:   nttb200_generate_coeff [N [Q [seed]]]
 * The files are what time_testing256.c:17-44 (and tools/nttb200_time_testing256.c) read back.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>

#include "nttb200_gen.h"

static int gerar_e_escrever(const char *nome_arquivo, int n, int q) {
  int32_t *a = (int32_t *)malloc((size_t)n * sizeof *a);
  if (!a) return -1;
  for (int i = 0; i < n; i++) a[i] = rand() % q;          /* generate_coeff.c:47 */
  const int rc = nttb200_write_coeff_file(nome_arquivo, a, (size_t)n);
  if (rc) perror("Erro ao abrir/criar o arquivo");
  free(a);
  return rc;
}

int main(int argc, char **argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 256;
  const int q = argc > 2 ? atoi(argv[2]) : 12289;
  if (n < 1 || q < 2) {
    fprintf(stderr, "usage: %s [N [Q [seed]]]\n", argv[0]);
    return 2;
  }
  srand(argc > 3 ? (unsigned)strtoul(argv[3], NULL, 0) : (unsigned)time(NULL));
  printf("Gerando coeficientes aleatorios com Q=%d e N=%d.\n", q, n);
  if (gerar_e_escrever("coeficientes_a.txt", n, q) || gerar_e_escrever("coeficientes_b.txt", n, q)) return 1;
  printf("Arquivos 'coeficientes_a.txt' e 'coeficientes_b.txt' gerados com sucesso!\n");
  return 0;
}
