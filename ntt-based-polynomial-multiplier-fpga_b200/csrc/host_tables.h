/*
 * host_tables.h -- host-side number theory and twiddle-table generation (plain C).
 *
 * The reference ships its tables as literals for (n=256, q=12289, psi=1002) only
 * (R/NTT/ntt256_tables.C) and has no generator for them; these functions produce
 * the same tables from their closed forms for any (n, q, psi)  (SURVEY 8a-T).
 */
#ifndef NTTB200_HOST_TABLES_H
#define NTTB200_HOST_TABLES_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

uint32_t ht_powmod(uint32_t base, uint64_t exp, uint32_t q);
uint32_t ht_invmod(uint32_t a, uint32_t q);
uint32_t ht_bitrev(uint32_t x, uint32_t bits);
uint32_t ht_log2(uint32_t n);
/* Shoup companion floor(w * 2^32 / q) */
static inline uint32_t ht_shoup(uint32_t w, uint32_t q) {
  return (uint32_t)(((uint64_t)w << 32) / q);
}
/* p[t+j] = lead^(n/2t) * root^((n/2t) * (rev ? bitrev_t(j) : j)),  p[0] = 0 */
void ht_level_table(uint32_t *out, uint32_t n, uint32_t q, uint32_t lead, uint32_t root, int rev);

#ifdef __cplusplus
}
#endif
#endif
