/*
 * nttb200.h -- C ABI of libnttb200.so: batched NTT-based negacyclic polynomial
 * multiplication  c = INTT(NTT(a) o NTT(b))  in Z_q[x]/(x^n + 1)  on NVIDIA B200.
 *
 * Plain C, plain pointers and sizes; no CUDA or torch types in any signature (a
 * CUDA stream crosses the boundary as `void *`).  This is the layer that replaces
 * the reference's host<->FPGA transfer path
 *   Multiplier_NTT_Based/Software_Hardware_Comunnicator/linux_app/NTT_PCIECommunicationv2.c:109-252
 *   (mode 1/2: DMA polyA/polyB to the board, mode 3: GO, DMA polyC back)
 * with CUDA streams and pinned buffers, and that the reference's product/NTT call
 * surface (nttb200_legacy.h) is implemented on.
 *
 * Reference paths below are relative to
 *   R/ = Multiplier_NTT_Based/NTT_Software/NTT_Software_Evaluations/NTT-256/
 *
 * Data layout everywhere: row-major int32_t [batch][n], one polynomial per row,
 * coefficient i of polynomial r at r*n + i, values in [0, q)  (the reference's
 * `int32_t a[256]`, R/NTT/ntt256.h:76-83, batched).
 *
 * There is NO CPU fallback: every entry point that computes fails with
 * NTTB200_ECUDA when no CUDA device is usable.
 */
#ifndef NTTB200_H
#define NTTB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NTTB200_VERSION 200

/* ---- error codes (0 = OK).  The reference functions are `void` and abort on
 * violated preconditions (assert in R/NTT-RED/ntt_red.c:42,79,94); the nttb200_*
 * functions return a code instead and keep a per-thread message. --------------- */
enum {
  NTTB200_OK = 0,
  NTTB200_EPARAM = -1,   /* bad (n, q, psi), NULL pointer, unsupported size       */
  NTTB200_ECUDA = -2,    /* CUDA runtime/driver error, or no device               */
  NTTB200_ENOMEM = -3,   /* host or device allocation failed                      */
  NTTB200_ERANGE = -4    /* an input coefficient outside [0, q): looked for by plans created
                            with NTTB200_PLAN_CHECK_RANGE only                     */
};
const char *nttb200_last_error(void);
int nttb200_version(void);

/* ---- devices ------------------------------------------------------------------ */
int nttb200_device_count(void);          /* >= 0, or NTTB200_ECUDA                  */
int nttb200_set_device(int device);      /* device for plans created by this thread */
int nttb200_get_device(void);

/* ---- plans: one (n, q, psi) parameter set, its twiddle tables on one GPU ------- */
typedef struct nttb200_plan nttb200_plan;

/* flags */
#define NTTB200_PLAN_DEFAULT 0u
#define NTTB200_PLAN_CYCLIC 1u /* psi-free surface: product mod x^n - 1 (needs n | q-1 only;
                                  R/NTT/ntt256.h:28,37 + R/NTT/ntt.h:52).  `psi` is then read
                                  as omega (0 = smallest primitive n-th root).            */

#define NTTB200_PLAN_NO_PLANTARD 2u /* diagnostics: products of half-word moduli (q <= 12385) use the
                                      generic Shoup/Montgomery kernel instead of the Plantard one */

#define NTTB200_PLAN_CHECK_RANGE 4u /* products and transforms first verify that every input coefficient
                                     lies in [0, q) -- the reference's unchecked precondition
                                     (R/NTT/ntt256.h:82-83) -- and return NTTB200_ERANGE naming the
                                     first offender.  Host arrays are scanned where they lie; for
                                     device arrays the call waits for a checking kernel (it is then
                                     synchronous).  A debugging aid: off on the timed path.           */

/* n: power of two, 8 <= n <= 2^17.  q: odd prime < 2^31 with 2n | q-1 (n | q-1 when
 * CYCLIC).  psi: primitive 2n-th root of unity mod q, or 0 for the smallest one -- the
 * rule of Generator_Params/generate_params.C:25-44.  The reference's own tables use
 * psi = 1002 for (256, 12289) (R/NTT/ntt256_tables.h:20); products do not depend on the
 * choice, standalone transforms do. */
int nttb200_plan_create(nttb200_plan **plan, uint32_t n, uint32_t q, uint32_t psi, uint32_t flags);
void nttb200_plan_destroy(nttb200_plan *plan);
uint32_t nttb200_plan_n(const nttb200_plan *plan);
uint32_t nttb200_plan_q(const nttb200_plan *plan);
uint32_t nttb200_plan_psi(const nttb200_plan *plan);
int nttb200_plan_device(const nttb200_plan *plan);
/* name of the kernel family / arithmetic class the plan dispatches to (for reports) */
const char *nttb200_plan_describe(const nttb200_plan *plan);

/* ---- products ------------------------------------------------------------------
 * Replaces, batched:  ntt256_product1 / ntt256_product4 (R/NTT/ntt256.C:5-24) and
 * ntt_red256_product1 / ntt_red256_product4 (R/NTT-RED/ntt_red256.C:5-52); all four
 * return the same canonical c, so one entry point serves them.  Unlike the reference
 * (R/NTT/ntt256.h:80 "arrays a and b are modified") a and b are left untouched;
 * c must not overlap a or b.
 *
 * Host buffers: pageable or pinned (nttb200_host_alloc); the call stages through
 * pinned memory, overlaps H2D / kernel / D2H across an internal stream ring and
 * returns when c is complete. */
int nttb200_polymul_batch(nttb200_plan *plan, int32_t *c, const int32_t *a, const int32_t *b,
                          size_t batch);
/* The same product, asynchronous: the call queues the job and returns a ticket; a worker thread of
 * the plan runs the queue through the SAME pipeline as the synchronous call, as one stream of jobs --
 * so the first chunks of a product are narrowed and sent while the last chunks of the one before it
 * drain (consecutive synchronous calls leave their fill and drain bare: about 9 % of a 2^16-row call).
 * a, b must stay valid and unchanged, and c untouched, until nttb200_polymul_wait(plan, ticket) has
 * returned (its return value is the product's status; ticket 0 waits for every product queued so far).
 * Products complete in the order they were queued; a ticket is waited for once, by one thread.
 * nttb200_plan_destroy runs the queue to its end. */
int nttb200_polymul_batch_async(nttb200_plan *plan, int32_t *c, const int32_t *a, const int32_t *b,
                                size_t batch, unsigned long long *ticket);
int nttb200_polymul_wait(nttb200_plan *plan, unsigned long long ticket);
/* Wire format of the host-buffer call above, half-word moduli (q <= 12385, n <= 1024), calls of
 * 2^19 words per operand or more: the call is bound by the PCIe link, so chunks of the batch
 * cross it as 16-bit words -- narrowed from / widened into the caller's int32_t rows by a pool of
 * host threads through pinned staging (csrc/hostwire.c) -- as long as the host keeps up, and as
 * the caller's 32-bit words otherwise; a chunk holding a word that does not fit 16 bits always
 * travels as 32-bit words, so results never depend on the wire.  This stands where the
 * reference streams 32-bit FIFO words to the board (COMM/linux_app/NTT_PCIECommunicationv2.c:
 * 166-224).  The result rows of a 16-bit chunk return as 16-bit words that the pool widens into c
 * (NTTB200_WIRE_C32=1: as int32 words written by the kernel and copied straight into a pinned c).
 * Pageable (malloc) buffers of the other plans with n <= 1024 go through the same pipeline with
 * the pool as a parallel 32-bit stager (NTTB200_STAGE_PAGEABLE=0: the driver's pageable path).
 * How much is narrowed follows the size of the pool: 12 threads or more narrow every chunk, 6 to 11
 * mix 32-bit chunks in when the pool lags, fewer leave pinned buffers to the DMA engines alone.
 * Environment: NTTB200_WIRE=auto|16|32, NTTB200_HOST_THREADS (default: the CPUs the process may
 * run on, at most 32).  Statistics of the plan's last host-buffer call: polynomials whose
 * operands crossed the link as 16-bit / as 32-bit words (both 0 for calls below the threshold),
 * and those whose result came back as 16-bit words: */
int nttb200_plan_wire_stats(const nttb200_plan *plan, unsigned long long *rows16,
                            unsigned long long *rows32, unsigned long long *rows_c16,
                            int *host_threads);
/* Device-resident buffers (the timed path: no PCIe in the loop).  Asynchronous on
 * `stream` (a cudaStream_t passed as void*, NULL = the legacy default stream).  Calls on one
 * stream run one after the other; a caller with several independent batches gets 5-12 % more
 * out of the GPU by alternating between two streams (the tail of one launch then overlaps the
 * start of the next; bench.py reports both figures). */
int nttb200_polymul_batch_dev(nttb200_plan *plan, int32_t *c_dev, const int32_t *a_dev,
                              const int32_t *b_dev, size_t batch, void *stream);

/* ---- packed 16-bit I/O: an EXTENSION outside the reference API for plans with
 * q <= 12385 and n <= 1024: coefficients travel as uint16_t, which halves the PCIe and HBM
 * bytes per product (6 n instead of 12 n).  Never used for the roofline figures of bench.py.
 * Operands must be 16-byte aligned. */
int nttb200_polymul_batch_u16(nttb200_plan *plan, uint16_t *c, const uint16_t *a, const uint16_t *b,
                              size_t batch);
int nttb200_polymul_batch_u16_dev(nttb200_plan *plan, uint16_t *c_dev, const uint16_t *a_dev,
                                  const uint16_t *b_dev, size_t batch, void *stream);

/* ---- several GPUs of one box: contiguous batch slices, one plan and one host thread per GPU,
 * no collective (independent products share no data).  One process per GPU with
 * nttb200_set_device() + nttb200_shard_bounds() is the other way (what bench.py does). ------- */
typedef struct nttb200_multi nttb200_multi;
int nttb200_multi_create(nttb200_multi **multi, uint32_t n, uint32_t q, uint32_t psi, uint32_t flags,
                         int ngpus /* 0 = every visible GPU */);
void nttb200_multi_destroy(nttb200_multi *multi);
int nttb200_multi_gpus(const nttb200_multi *multi);
nttb200_plan *nttb200_multi_plan(nttb200_multi *multi, int gpu);
/* host buffers; rows [g B/G, (g+1) B/G) are multiplied on GPU g; returns when c is complete */
int nttb200_multi_polymul_batch(nttb200_multi *multi, int32_t *c, const int32_t *a, const int32_t *b,
                                size_t batch);
void nttb200_shard_bounds(size_t batch, int world, int rank, size_t *lo, size_t *hi);

/* ---- standalone transforms ("the NTT call surface") ---------------------------
 * In place on [batch][n].  Outputs are canonical [0, q) and bit-identical to the
 * reference function of the same dataflow run with the same table.             */
enum nttb200_transform {
  /* forward, standard order in -> bit-reversed order out */
  NTTB200_NTT_STD2REV = 0,      /* ntt_ct_std2rev(omega_powers_rev) == ntt_gs_std2rev(omega_powers)
                                   R/NTT/ntt.C:295-329, 467-493                             */
  NTTB200_MULNTT_STD2REV = 1,   /* mulntt_ct_std2rev(mixed_powers_rev): NTT of a[i]*psi^i
                                   R/NTT/ntt.C:342-371                                      */
  /* inverse (UNSCALED, like the reference: intt(ntt(a)) = n*a, R/NTT/ntt256.h:16-17),
   * bit-reversed order in -> standard order out */
  NTTB200_INTT_REV2STD = 2,     /* ntt_gs_rev2std(inv_omega_powers_rev) == ntt_ct_rev2std(inv_omega_powers)
                                   R/NTT/ntt.C:387-416, 216-243                             */
  NTTB200_INTTMUL_REV2STD = 3,  /* nttmul_gs_rev2std(inv_mixed_powers_rev): result[i]*psi^-i
                                   R/NTT/ntt.C:428-451                                      */
  /* same as 2/3 with the n^-1 scaling folded in (what the products use) */
  NTTB200_INTT_REV2STD_SCALED = 4,
  NTTB200_INTTMUL_REV2STD_SCALED = 5,
  /* forward with the INVERSE root, standard in -> bit-reversed out, and inverse-root
   * twins used by intt256_ct_std2rev / intt256_gs_std2rev (R/NTT/ntt256.h:45-51) */
  NTTB200_INTT_STD2REV = 6,
  /* forward root, bit-reversed in -> standard out: ntt256_ct_rev2std / ntt256_gs_rev2std
   * (R/NTT/ntt256.h:20-26) */
  NTTB200_NTT_REV2STD = 7
};
/* (host array, pinned or pageable; a large pageable array is staged through pinned memory by the
 * host pool, as in nttb200_polymul_batch) */
int nttb200_ntt_batch(nttb200_plan *plan, int transform, int32_t *a, size_t batch);
/* a_dev must be 16-byte aligned (NTTB200_EPARAM otherwise): rows move with 128-bit accesses */
int nttb200_ntt_batch_dev(nttb200_plan *plan, int transform, int32_t *a_dev, size_t batch,
                          void *stream);

/* Table-driven transforms: the exact dataflow of the reference function run with the
 * CALLER's table `p` (n entries, layout p[t+j] of R/NTT/ntt256_tables.C) -- what the
 * legacy drop-ins in nttb200_legacy.h are built on.  `dataflow`: */
enum nttb200_dataflow {
  NTTB200_DF_CT_STD2REV = 0,    /* R/NTT/ntt.C:295-371  (ntt_ct_std2rev, mulntt_ct_std2rev)   */
  NTTB200_DF_GS_REV2STD = 1,    /* R/NTT/ntt.C:387-451  (ntt_gs_rev2std, nttmul_gs_rev2std)   */
  NTTB200_DF_CT_REV2STD = 2,    /* R/NTT/ntt.C:216-278  (ntt_ct_rev2std, mulntt_ct_rev2std)   */
  NTTB200_DF_GS_STD2REV = 3     /* R/NTT/ntt.C:467-525  (ntt_gs_std2rev, nttmul_gs_std2rev)   */
};
/* skip_j0 = 1: the UN-MERGED entry points (ntt_ct_std2rev, ntt_ct_rev2std, ntt_gs_std2rev,
 * ntt_gs_rev2std), which peel the j = 0 block -- its butterflies run without a multiplication
 * and p[t] is never read (R/NTT/ntt.C:226-231, 313-317, 401-405, 477-481); 0: the psi-merged
 * ones (mulntt_ct_*, nttmul_gs_*), which multiply every block by p[t+j].  With the reference's
 * own tables (p[t] = 1) the two coincide. */
int nttb200_ntt_table_batch(uint32_t n, uint32_t q, int dataflow, int skip_j0, const uint32_t *p,
                            int32_t *a, size_t batch);

/* ---- exact emulation of the Longa-Naehrig ("RED") surface, q = 12289 hard-wired as in the
 * reference (R/NTT-RED/ntt_red.c:24): signed, UNREDUCED 32-bit outputs, bit-identical to
 * ntt_red_{ct,gs}_{rev2std,std2rev} / mulntt_red_ct_* / nttmul_red_gs_* (ntt_red.c:244-554)
 * run with the caller's table p (n entries, already multiplied by 1/3 as the reference's
 * tables are).  skip_j0 = 1 for the un-merged entry points, which do the j = 0 butterflies
 * without a multiplication.  `dataflow` as enum nttb200_dataflow. */
int nttb200_red_ntt_table_batch(uint32_t n, int dataflow, int skip_j0, const int32_t *p, int32_t *a,
                                size_t batch);
enum nttb200_red_op {
  NTTB200_RED_NORMALIZE = 0,      /* normalize               ntt_red.c:72-82   */
  NTTB200_RED_NORMALIZE_INV3 = 1, /* normalize_inv3          ntt_red.c:87-97   */
  NTTB200_RED_SHIFT = 2,          /* shift_array             ntt_red.c:103-111 */
  NTTB200_RED_REDUCE = 3,         /* reduce_array            ntt_red.c:124-130 */
  NTTB200_RED_REDUCE_TWICE = 4,   /* reduce_array_twice      ntt_red.c:138-144 */
  NTTB200_RED_CORRECT = 5,        /* correct                 ntt_red.c:150-169 */
  NTTB200_RED_MUL_RED = 6,        /* mul_reduce_array[16]    ntt_red.c:197-211: c[i] = mul_red(a[i], b[i]) */
  NTTB200_RED_SCALAR_MUL_RED = 7  /* scalar_mul_reduce_array ntt_red.c:217-223 */
};
/* c may alias a; b is read by MUL_RED only, scalar by SCALAR_MUL_RED only */
int nttb200_red_elementwise_batch(int op, int32_t *c, const int32_t *a, const int32_t *b, int32_t scalar,
                                  size_t count);
/* in-place bit-reversal permutation of every row (bitrev_shuffle, R/NTT/ntt.C:27-44) and the
 * table-driven variant (shuffle_with_table, ntt.C:50-59: npairs swaps a[p[i][0]] <-> a[p[i][1]]) */
int nttb200_bitrev_shuffle_batch(int32_t *a, uint32_t n, size_t batch);
int nttb200_shuffle_with_table(int32_t *a, size_t words, const uint16_t *pairs, uint32_t npairs);

/* ---- elementwise surface (R/NTT/ntt.C:119-153), batched, host buffers ----------- */
int nttb200_mul_array_batch(nttb200_plan *plan, int32_t *c, const int32_t *a, const int32_t *b,
                            size_t batch);                       /* c[i] = a[i]*b[i] mod q   */
int nttb200_scalar_mul_array_batch(nttb200_plan *plan, int32_t *a, int32_t s, size_t batch);

/* ---- host-side tables, reference layout (R/NTT/ntt256_tables.C) ---- */
enum nttb200_table {
  NTTB200_PSI_POWERS = 0, NTTB200_INV_PSI_POWERS, NTTB200_SCALED_INV_PSI_POWERS,
  NTTB200_OMEGA_POWERS, NTTB200_OMEGA_POWERS_REV, NTTB200_INV_OMEGA_POWERS,
  NTTB200_INV_OMEGA_POWERS_REV, NTTB200_MIXED_POWERS, NTTB200_MIXED_POWERS_REV,
  NTTB200_INV_MIXED_POWERS, NTTB200_INV_MIXED_POWERS_REV, NTTB200_INV_PSI_POWERS_REV,
  NTTB200_TABLE_COUNT
};
/* fills out[0..n) */
int nttb200_make_table(int kind, uint32_t n, uint32_t q, uint32_t psi, uint32_t *out);
/* The Longa-Naehrig table set (R/NTT-RED/ntt_red256_tables.c:16-469), q = 12289: the table of the
 * same `kind` times 3^-1, centred to (-q/2, q/2]; SCALED_INV_PSI_POWERS carries n^-1 3^-8
 * (ntt_red256_tables.h:28, rescale8 = 8822 at n = 256) and the extra kind below n^-1 3^-6
 * (rescale6 = 5664).  fills out[0..n) */
#define NTTB200_RED_SCALED_INV_PSI_POWERS_VAR 100
int nttb200_make_red_table(int kind, uint32_t n, uint32_t psi, int32_t *out);
uint32_t nttb200_find_psi(uint32_t n, uint32_t q);     /* smallest primitive 2n-th root, 0 if none */
uint32_t nttb200_find_omega(uint32_t n, uint32_t q);   /* smallest primitive n-th root, 0 if none  */
int nttb200_is_prime(uint32_t q);

/* ---- pinned host buffers (replace the malloc'd DMA buffers of
 * NTT_PCIECommunicationv2.c:120-131) ------------------------------------------- */
void *nttb200_host_alloc(size_t bytes);
void nttb200_host_free(void *p);
/* device buffers, for callers that keep operands resident between calls */
void *nttb200_dev_alloc(size_t bytes);
void nttb200_dev_free(void *p);
int nttb200_memcpy_h2d(void *dst_dev, const void *src_host, size_t bytes, void *stream);
int nttb200_memcpy_d2h(void *dst_host, const void *src_dev, size_t bytes, void *stream);
int nttb200_stream_sync(void *stream);
/* streams for callers without the CUDA runtime (the Terasic-ABI shim): create / destroy /
 * query (1 = all work done, 0 = still running, < 0 = error) */
void *nttb200_stream_create(void);
void nttb200_stream_destroy(void *stream);
int nttb200_stream_query(void *stream);

/* ---- measurement helpers (used by bench.py; device-side, no host data) ---------- */
/* runs `iters` dependent IMAD / IMAD.HI / IADD3 chains on every SM and returns the
 * measured chip-wide rate in lane-ops per second (the INT32-pipe roofline
 * denominator).  which: 0 = IMAD (mad.lo), 1 = IMAD.HI (mul.hi),
 * 2 = IADD3/LOP3, 3 = the Shoup butterfly mix (3 IMAD + 2 ALU). */
int nttb200_measure_int_peak(int which, double *lane_ops_per_s);
/* number of kernels the last nttb200_polymul_batch* / ntt_batch* call launched */
int nttb200_last_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* NTTB200_H */
