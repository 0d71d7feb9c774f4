/*
 * nttb200_gen.h -- parameter generator for the FPGA-style datapath parameters and the text
 * formats around the path.  Host-side C, no GPU involved.
 *
 * Replaces  G/ = Multiplier_NTT_Based/NTT_Software/Generator_Params/
 *   generate_params()    G/generate_params.C:12-52   (declared G/generate_params.h:16)
 *   generate_twiddles()  G/generate_params.C:54-73   (declared G/generate_params.h:18)
 *   modexp / miller_rabin / is_prime / generate_large_prime   G/prime_generate.C:9-107
 *   egcd / modinv        G/helper.C:5-35
 *   coefficient files    G/generate_coeff.c:35-59, reader R/time_testing256.c:17-44
 * with runtime (n, K, P, q) instead of the reference's #define N 256 / K 13 / P 8 and its
 * hard-wired q = 12289 (G/generate_params.C:22).
 */
#ifndef NTTB200_GEN_H
#define NTTB200_GEN_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  uint32_t n, q;          /* ring size, modulus                                              */
  uint32_t psi, psi_inv;  /* smallest element of exact order 2n, its inverse                 */
  uint32_t w, w_inv;      /* psi^2 and its inverse                                           */
  uint32_t R;             /* 2^((log2 n + 1) * ceil(K / (log2 n + 1)))  (word-level Montgomery) */
  uint32_t n_inv;         /* n^-1 mod q                                                      */
  uint32_t PE;            /* 2 * P                                                           */
} nttb200_gen_params_t;

/* q = 0 means the reference's fixed 12289.  K = bit width of the datapath (reference: 13),
 * P = number of butterfly units (reference: 8).  Returns 0, or NTTB200_EPARAM (-1) when q is
 * not a prime with 2n | q-1. */
int nttb200_gen_params(uint32_t n, uint32_t K, uint32_t P, uint32_t q, nttb200_gen_params_t *out);

/* number of words generate_twiddles writes per array: sum over j < log2 n of
 * max(1, (n / 2P) >> j) * P   (272 for n = 256, P = 8) */
size_t nttb200_gen_twiddle_count(uint32_t n, uint32_t P);
/* W[idx] = w^e R mod q, W_INV[idx] = w_inv^e R mod q in processing-element order,
 * e = ((P << j) k + (i << j)) mod n/2 */
int nttb200_gen_twiddles(uint32_t *W, uint32_t *W_INV, uint32_t n, uint32_t P, uint32_t w, uint32_t w_inv,
                         uint32_t q, uint32_t R);

/* number theory helpers (deterministic Miller-Rabin for 32-bit inputs) */
uint32_t nttb200_modexp(uint32_t base, uint32_t exp, uint32_t mod);
int32_t nttb200_modinv(int32_t a, int32_t m);               /* -1 when no inverse exists */
int nttb200_miller_rabin(uint32_t p);
/* first prime >= 2^(k-1) found from `seed` with q = 1 (mod 2n); 0 if none below 2^k */
uint32_t nttb200_gen_prime(uint32_t k, uint32_t n, uint64_t seed);

/* ---- the reference's own entry points, same signatures, N=256 K=13 P=8 q=12289 ---- */
void generate_params(int *psi, int *psi_inv, int *w, int *w_inv, int *R, int *n_inv, int *PE, int *q);
void generate_twiddles(uint32_t W[], uint32_t W_INV[], uint32_t w, uint32_t w_inv, uint32_t q, uint32_t R);

/* ---- text formats ------------------------------------------------------------------- */
/* coefficient file: decimals separated by blanks, 10 per line (G/generate_coeff.c:44-57) */
int nttb200_write_coeff_file(const char *path, const int32_t *a, size_t n);
/* reads up to n integers (R/time_testing256.c:17-44); returns how many were read, -1 on open failure */
long nttb200_read_coeff_file(const char *path, int32_t *a, size_t n);
/* result format of R/time_testing256.c:46-64: two blanks, %5d, 16 per line */
void nttb200_print_array(void *FILE_ptr, const int32_t *a, size_t n);
/* one hexadecimal word per line, lower case, no prefix (the HW vector files
 * Hardware_Multiplier/simulation/modelsim/test/ *.txt, colab_programs/dec2Hex.py) */
int nttb200_write_hex_file(const char *path, const uint32_t *a, size_t n);
long nttb200_read_hex_file(const char *path, uint32_t *a, size_t n);

#ifdef __cplusplus
}
#endif
#endif
