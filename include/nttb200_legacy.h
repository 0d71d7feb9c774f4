/*
 * nttb200_legacy.h -- the reference's own C call surface, exported by libnttb200.so so
 * that a program written against NTT_Software links against the GPU library unchanged
 * (same names, same argument meaning; q = 12289 is hard-wired in the reference,
 * R/NTT/ntt.C:18, R/NTT-RED/ntt_red.c:24).
 *
 * R/ = Multiplier_NTT_Based/NTT_Software/NTT_Software_Evaluations/NTT-256/
 *
 * Error behaviour: the reference functions are `void`; violated preconditions abort()
 * through assert (R/NTT-RED/ntt_red.c:42).  The drop-ins keep `void` and likewise
 * abort() -- after printing nttb200_last_error() to stderr -- when the GPU call fails
 * (including "no CUDA device": there is no CPU fallback).
 *
 * Each call moves ONE polynomial over PCIe and back (15 us per product call on a B200 box: the
 * kernel reads and writes a mapped pinned buffer, no DMA); that is the reference's calling
 * convention, not the fast path.  Batch work belongs on nttb200_polymul_batch().
 */
#ifndef NTTB200_LEGACY_H
#define NTTB200_LEGACY_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- products: c = a*b mod (x^256+1, 12289), inputs and result in [0, 12289) ------- */
void ntt256_product1(int32_t *c, int32_t *a, int32_t *b);      /* R/NTT/ntt256.h:85, ntt256.C:5-13  */
void ntt256_product4(int32_t *c, int32_t *a, int32_t *b);      /* R/NTT/ntt256.h:86, ntt256.C:16-24 */
void ntt_red256_product1(int32_t *c, int32_t *a, int32_t *b);  /* R/NTT-RED/ntt_red256.h:87, ntt_red256.C:5-27  */
void ntt_red256_product4(int32_t *c, int32_t *a, int32_t *b);  /* R/NTT-RED/ntt_red256.h:90, ntt_red256.C:30-52 */
/* The reference documents "arrays a and b are modified" (R/NTT/ntt256.h:80): after
 * ntt256_product1/4 they hold the psi-twisted forward NTT in bit-reversed order.  By
 * default the drop-ins leave a and b untouched; nttb200_legacy_set_clobber(1) reproduces the
 * reference's post-state for all four: the canonical transform for ntt256_product1/4, and for
 * the optimized pair the UNREDUCED representative their pipeline leaves behind (shift -> x psi^i
 * -> forward transform -> reduce_array, R/NTT-RED/ntt_red256.C:5-13, 31-39), bit for bit. */
void nttb200_legacy_set_clobber(int on);

/* ---- generic-n transforms with the caller's table (R/NTT/ntt.h:71-183) ---------------- */
void ntt_ct_rev2std_v1(int32_t *a, uint32_t n, const uint16_t *p);   /* ntt.h:71,  ntt.C:168-197 */
void ntt_ct_rev2std(int32_t *a, uint32_t n, const uint16_t *p);      /* ntt.h:90,  ntt.C:216-243 */
void mulntt_ct_rev2std(int32_t *a, uint32_t n, const uint16_t *p);   /* ntt.h:100, ntt.C:253-278 */
void ntt_ct_std2rev(int32_t *a, uint32_t n, const uint16_t *p);      /* ntt.h:117, ntt.C:295-329 */
void mulntt_ct_std2rev(int32_t *a, uint32_t n, const uint16_t *p);   /* ntt.h:127, ntt.C:342-371 */
void ntt_gs_rev2std(int32_t *a, uint32_t n, const uint16_t *p);      /* ntt.h:143, ntt.C:387-416 */
void nttmul_gs_rev2std(int32_t *a, uint32_t n, const uint16_t *p);   /* ntt.h:155, ntt.C:428-451 */
void ntt_gs_std2rev(int32_t *a, uint32_t n, const uint16_t *p);      /* ntt.h:171, ntt.C:467-493 */
void nttmul_gs_std2rev(int32_t *a, uint32_t n, const uint16_t *p);   /* ntt.h:183, ntt.C:505-525 */

/* ---- elementwise (R/NTT/ntt.h:37-52) ------------------------------------------------- */
void mul_array16(int32_t *a, uint32_t n, const uint16_t *p);                      /* ntt.C:119-125 */
void mul_array(int32_t *c, uint32_t n, const int32_t *a, const int32_t *b);       /* ntt.C:131-137 */
void scalar_mul_array(int32_t *a, uint32_t n, int32_t c);                         /* ntt.C:147-153 */

/* ---- permutations (R/NTT/ntt.h:23-30) ------------------------------------------------------ */
void bitrev_shuffle(int32_t *a, uint32_t n);                                   /* ntt.C:27-44 */
void shuffle_with_table(int32_t *a, const uint16_t p[][2], uint32_t n);          /* ntt.C:50-59 */

/* ---- Longa-Naehrig surface (R/NTT-RED/ntt_red.h:65-284), exact: the transforms return the
 * same UNREDUCED signed values as the reference (R/NTT-RED/ntt_red256.h:18) ------------------ */
void normalize(int32_t *a, uint32_t n);                                        /* ntt_red.c:72-82   */
void normalize_inv3(int32_t *a, uint32_t n);                                   /* ntt_red.c:87-97   */
void shift_array(int32_t *a, uint32_t n);                                      /* ntt_red.c:103-111 */
void reduce_array(int32_t *a, uint32_t n);                                     /* ntt_red.c:124-130 */
void reduce_array_twice(int32_t *a, uint32_t n);                               /* ntt_red.c:138-144 */
void correct(int32_t *a, uint32_t n);                                          /* ntt_red.c:150-169 */
void mul_reduce_array16(int32_t *a, uint32_t n, const int16_t *p);             /* ntt_red.c:197-203 */
void mul_reduce_array(int32_t *c, uint32_t n, const int32_t *a, const int32_t *b);   /* ntt_red.c:205-211 */
void scalar_mul_reduce_array(int32_t *a, uint32_t n, int32_t c);               /* ntt_red.c:217-223 */
void ntt_red_ct_rev2std(int32_t *a, uint32_t n, const int16_t *p);             /* ntt_red.c:244-270 */
void mulntt_red_ct_rev2std(int32_t *a, uint32_t n, const int16_t *p);          /* ntt_red.c:280-305 */
void ntt_red_ct_std2rev(int32_t *a, uint32_t n, const int16_t *p);             /* ntt_red.c:321-355 */
void mulntt_red_ct_std2rev(int32_t *a, uint32_t n, const int16_t *p);          /* ntt_red.c:368-400 */
void ntt_red_gs_rev2std(int32_t *a, uint32_t n, const int16_t *p);             /* ntt_red.c:414-443 */
void nttmul_red_gs_rev2std(int32_t *a, uint32_t n, const int16_t *p);          /* ntt_red.c:456-480 */
void ntt_red_gs_std2rev(int32_t *a, uint32_t n, const int16_t *p);             /* ntt_red.c:495-520 */
void nttmul_red_gs_std2rev(int32_t *a, uint32_t n, const int16_t *p);          /* ntt_red.c:534-554 */

/* ---- the n = 256 tables, under the reference's names (R/NTT/ntt256_tables.h:29-44,
 * R/NTT-RED/ntt_red256_tables.h:34-51): read-only data of libnttb200.so, generated at build time
 * from the closed forms (csrc/gen_legacy_tables.c), so that the reference's header-only
 * wrappers -- ntt256_ct_std2rev(a) = ntt_ct_std2rev(a, 256, ntt256_omega_powers_rev) etc.,
 * ntt256.h:20-69, ntt_red256.h:21-70 -- link against this library alone.  The scalar
 * parameters (psi = 1002 ... rescale8 = 8822) are `static const` in the reference headers. */
extern const uint16_t ntt256_psi_powers[256], ntt256_inv_psi_powers[256], ntt256_scaled_inv_psi_powers[256];
extern const uint16_t ntt256_omega_powers[256], ntt256_omega_powers_rev[256];
extern const uint16_t ntt256_inv_omega_powers[256], ntt256_inv_omega_powers_rev[256];
extern const uint16_t ntt256_mixed_powers[256], ntt256_mixed_powers_rev[256];
extern const uint16_t ntt256_inv_mixed_powers[256], ntt256_inv_mixed_powers_rev[256];
extern const int16_t ntt_red256_psi_powers[256], ntt_red256_inv_psi_powers[256];
extern const int16_t ntt_red256_scaled_inv_psi_powers[256], ntt_red256_scaled_inv_psi_powers_var[256];
extern const int16_t ntt_red256_omega_powers[256], ntt_red256_omega_powers_rev[256];
extern const int16_t ntt_red256_inv_omega_powers[256], ntt_red256_inv_omega_powers_rev[256];
extern const int16_t ntt_red256_mixed_powers[256], ntt_red256_mixed_powers_rev[256];
extern const int16_t ntt_red256_inv_mixed_powers[256], ntt_red256_inv_mixed_powers_rev[256];

#ifdef __cplusplus
}
#endif
#endif /* NTTB200_LEGACY_H */
