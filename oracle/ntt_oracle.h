/*
 * ntt_oracle.h -- CPU oracle for the NTT polynomial-multiplication hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it, and only as the checker / reported baseline.
 *
 * It is a parametrised (runtime n, q, psi; 32-bit tables) restatement of the
 * reference's single-threaded C algorithms.  Paths cited below are relative to
 *   /root/reference/Multiplier_NTT_Based/NTT_Software/NTT_Software_Evaluations/NTT-256/
 * The reference itself is hard-wired to (n=256, q=12289, psi=1002); this oracle is
 * pinned to it there (tests/test_oracle_vs_ref.py diff it against oracle/_ref, the
 * reference compiled unmodified) and to the reference's known-answer vectors and the
 * hardware golden vectors (q=7681) for a second modulus.
 */
#ifndef NTT_ORACLE_H
#define NTT_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- modular helpers ------------------------------------------------------ */
uint32_t orc_powmod(uint32_t b, uint64_t e, uint32_t q);
uint32_t orc_invmod(uint32_t a, uint32_t q);            /* q prime */
int      orc_is_prime(uint32_t q);
/* smallest psi with psi^n == -1 (mod q), i.e. a primitive 2n-th root; 0 if none
 * (rule of Generator_Params/generate_params.C:25-44) */
uint32_t orc_smallest_psi(uint32_t n, uint32_t q);
/* smallest primitive n-th root of unity (for the cyclic, psi-free surface) */
uint32_t orc_smallest_omega(uint32_t n, uint32_t q);

/* ---- tables (layout of NTT/ntt256_tables.C) ------------------ */
enum orc_table {
  ORC_PSI_POWERS = 0,            /* p[i] = psi^i                               */
  ORC_INV_PSI_POWERS,            /* p[i] = psi^-i                              */
  ORC_SCALED_INV_PSI_POWERS,     /* p[i] = n^-1 psi^-i                         */
  ORC_OMEGA_POWERS,              /* p[t+j] = omega^((n/2t) j)                  */
  ORC_OMEGA_POWERS_REV,          /* p[t+j] = omega^((n/2t) rev_t(j))           */
  ORC_INV_OMEGA_POWERS,
  ORC_INV_OMEGA_POWERS_REV,
  ORC_MIXED_POWERS,              /* p[t+j] = psi^(n/2t) omega^((n/2t) j)       */
  ORC_MIXED_POWERS_REV,          /* p[t+j] = psi^(n/2t) omega^((n/2t) rev(j))  */
  ORC_INV_MIXED_POWERS,
  ORC_INV_MIXED_POWERS_REV,
  ORC_INV_PSI_POWERS_REV,        /* p[i] = psi^-rev_n(i) (present in the .C, unused) */
  ORC_TABLE_COUNT
};
/* omega = psi^2.  out has n entries (entry 0 of the level tables is 0, unused). */
int orc_make_table(int kind, uint32_t n, uint32_t q, uint32_t psi, uint32_t *out);
/* same with an explicit omega (psi ignored for the omega tables): cyclic surface */
int orc_make_omega_table(int kind, uint32_t n, uint32_t q, uint32_t omega, uint32_t *out);

/* ---- elementwise ops (NTT/ntt.C:119-153) ---------------------------------- */
void orc_mul_array_tab(int32_t *a, uint32_t n, const uint32_t *p, uint32_t q);   /* mul_array16 */
void orc_mul_array(int32_t *c, uint32_t n, const int32_t *a, const int32_t *b, uint32_t q);
void orc_scalar_mul_array(int32_t *a, uint32_t n, int32_t c, uint32_t q);
void orc_bitrev_shuffle(int32_t *a, uint32_t n);                                 /* ntt.C:27-44 */

/* ---- the nine transforms (NTT/ntt.C:168-525) ------------------------------ */
void orc_ntt_ct_rev2std_v1(int32_t *a, uint32_t n, const uint32_t *p, uint32_t q);
void orc_ntt_ct_rev2std   (int32_t *a, uint32_t n, const uint32_t *p, uint32_t q);
void orc_mulntt_ct_rev2std(int32_t *a, uint32_t n, const uint32_t *p, uint32_t q);
void orc_ntt_ct_std2rev   (int32_t *a, uint32_t n, const uint32_t *p, uint32_t q);
void orc_mulntt_ct_std2rev(int32_t *a, uint32_t n, const uint32_t *p, uint32_t q);
void orc_ntt_gs_rev2std   (int32_t *a, uint32_t n, const uint32_t *p, uint32_t q);
void orc_nttmul_gs_rev2std(int32_t *a, uint32_t n, const uint32_t *p, uint32_t q);
void orc_ntt_gs_std2rev   (int32_t *a, uint32_t n, const uint32_t *p, uint32_t q);
void orc_nttmul_gs_std2rev(int32_t *a, uint32_t n, const uint32_t *p, uint32_t q);

/* ---- products (NTT/ntt256.C:5-24 and the merged CT-fwd/GS-inv pipeline) ---- */
enum orc_variant {
  ORC_PRODUCT_CT = 1,      /* ntt256_product1: psi-twist, CT std2rev, CT rev2std, scaled psi^-1 */
  ORC_PRODUCT_GS = 4,      /* ntt256_product4: same skeleton with the GS loops               */
  ORC_PRODUCT_MERGED = 10, /* mulntt_ct_std2rev x2, mul_array, nttmul_gs_rev2std, * n^-1      */
  ORC_PRODUCT_SCHOOLBOOK = 20, /* O(n^2) definition, colab_programs/schoolbook.py:23-46       */
  ORC_PRODUCT_CYCLIC = 30  /* psi-free surface: ct_std2rev, mul_array, ct_rev2std(inv), *n^-1 */
};
typedef struct orc_plan orc_plan;
orc_plan *orc_plan_create(uint32_t n, uint32_t q, uint32_t psi /*0 = smallest*/);
void      orc_plan_destroy(orc_plan *P);
uint32_t  orc_plan_psi(const orc_plan *P);
const uint32_t *orc_plan_table(const orc_plan *P, int kind);
/* c = a*b in Z_q[x]/(x^n+1); a and b are clobbered exactly like the reference
 * does for variants 1 and 4 (ntt256.h:76-83).  Returns 0, or -1 on bad variant. */
int orc_product(const orc_plan *P, int variant, int32_t *c, int32_t *a, int32_t *b);
/* row-major [batch][n]; a and b are NOT clobbered (copied per row). */
int orc_product_batch(const orc_plan *P, int variant, int32_t *c, const int32_t *a,
                      const int32_t *b, size_t batch);
/* timed loop used by bench.py's cpu_baseline leg: runs the variant over the rows
 * round-robin for at least min_seconds; returns polymul/s of this thread (the
 * per-call operand restore is excluded from the timed region like
 * time_testing256.c:175-185 excludes reset_polyABC). */
double orc_bench_loop(const orc_plan *P, int variant, const int32_t *a, const int32_t *b,
                      size_t batch, double min_seconds, uint64_t *calls_out);

/* ---- q = 12289 Longa-Naehrig path (NTT-RED/ntt_red.c, ntt_red256.C) -------- */
int32_t orc_red(int32_t x);                                   /* ntt_red.c:34-36  */
int32_t orc_mul_red(int32_t x, int32_t y);                    /* ntt_red.c:39-46  */
void orc_red_shift_array(int32_t *a, uint32_t n);             /* ntt_red.c:103-111 */
void orc_red_reduce_array(int32_t *a, uint32_t n);            /* ntt_red.c:124-130 */
void orc_red_reduce_array_twice(int32_t *a, uint32_t n);      /* ntt_red.c:138-144 */
void orc_red_correct(int32_t *a, uint32_t n);                 /* ntt_red.c:150-169 */
void orc_red_normalize(int32_t *a, uint32_t n);               /* ntt_red.c:72-82  */
void orc_red_normalize_inv3(int32_t *a, uint32_t n);          /* ntt_red.c:87-97  */
void orc_red_mul_reduce_array_tab(int32_t *a, uint32_t n, const int32_t *p);
void orc_red_mul_reduce_array(int32_t *c, uint32_t n, const int32_t *a, const int32_t *b);
void orc_red_scalar_mul_reduce_array(int32_t *a, uint32_t n, int32_t c);
/* RED tables: level/psi tables * 3^-1 centred to (-q/2, q/2]; kind as above, plus: */
enum { ORC_RED_SCALED_INV_PSI_POWERS_VAR = 100 };  /* centre(n^-1 3^-6 psi^-i) */
int  orc_red_make_table(int kind, uint32_t n, uint32_t psi, int32_t *out);
void orc_red_ct_rev2std   (int32_t *a, uint32_t n, const int32_t *p);  /* ntt_red.c:244 */
void orc_red_mulntt_ct_rev2std(int32_t *a, uint32_t n, const int32_t *p); /* :280 */
void orc_red_ct_std2rev   (int32_t *a, uint32_t n, const int32_t *p);  /* ntt_red.c:321 */
void orc_red_mulntt_ct_std2rev(int32_t *a, uint32_t n, const int32_t *p); /* :368 */
void orc_red_gs_rev2std   (int32_t *a, uint32_t n, const int32_t *p);  /* ntt_red.c:414 */
void orc_red_nttmul_gs_rev2std(int32_t *a, uint32_t n, const int32_t *p); /* :456 */
void orc_red_gs_std2rev   (int32_t *a, uint32_t n, const int32_t *p);  /* ntt_red.c:495 */
void orc_red_nttmul_gs_std2rev(int32_t *a, uint32_t n, const int32_t *p); /* :534 */
/* ntt_red256_product1 (variant 1) / product4 (variant 4), generic n, q=12289 */
int  orc_red_product(uint32_t n, uint32_t psi, int variant, int32_t *c, int32_t *a, int32_t *b);

/* ---- synthetic inputs: splitmix64, value = next() % q ---------- */
void orc_fill_random(int32_t *dst, size_t count, uint32_t q, uint64_t seed);

#ifdef __cplusplus
}
#endif
#endif
