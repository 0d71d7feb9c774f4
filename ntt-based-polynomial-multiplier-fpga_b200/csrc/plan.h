/*
 * plan.h -- internal definition of nttb200_plan shared by the runtime translation units.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <mutex>
#include <vector>

#include "modarith.cuh"
#include "nttb200.h"

struct DevTable {
  uint2 *d = nullptr;            /* device: n entries (w, floor(w 2^32/q)), reference level layout */
  std::vector<uint2> h;          /* host mirror (entries < 64 feed the kernel-parameter block)     */
  uint32_t *d1 = nullptr;        /* device: the same table in Plantard form w~ (half-word moduli)  */
  std::vector<uint32_t> h1;
  uint32_t *d2 = nullptr;        /* device: Plantard form with W centred (signed kernel, ntt_small_splant.cuh) */
  std::vector<uint32_t> h2;
  uint32_t *d3 = nullptr;        /* device: entry by entry -(w^2) 2^32 mod q, centred: the group multiplication
                                    of the incomplete transform (ntt_small_splant.cuh)                         */
  std::vector<uint32_t> h3;
};

/* Plantard form of a constant: ((-w 2^32) mod q) * q^-1 mod 2^32  (ntt_small_plant.cuh) */
static inline uint32_t nttb200_plant_form(uint32_t w, uint32_t q, uint32_t qinv) {
  const uint64_t W = (q - (((uint64_t)(w % q)) << 32) % q) % q;
  return (uint32_t)W * qinv;
}

/* the same with W = (-w 2^32) mod q taken in (-q/2, q/2]: since q q^-1 = 1 (mod 2^32), one less
 * where W was above q/2 */
static inline uint32_t nttb200_plant_form_centred(uint32_t w, uint32_t q, uint32_t qinv) {
  const uint64_t W = (q - (((uint64_t)(w % q)) << 32) % q) % q;
  return (uint32_t)W * qinv - (W > q / 2 ? 1u : 0u);
}

enum PlanKernel { PK_SMALL = 0, PK_LARGE = 1 };
constexpr unsigned NTTB200_SCHED_SLOTS = 4096;

struct HostSlot {
  cudaStream_t stream = nullptr;
  cudaEvent_t done = nullptr;
  uint32_t *d_a = nullptr, *d_b = nullptr, *d_c = nullptr;
};

/* slot of the 16/32-bit wire pipeline of nttb200_polymul_batch (half-word moduli) */
struct WireSlot {
  cudaStream_t stream = nullptr;
  cudaEvent_t done = nullptr;
  uint32_t *d_a = nullptr, *d_b = nullptr, *d_c = nullptr;      /* device: words of either width   */
  uint32_t *h_a = nullptr, *h_b = nullptr, *h_c = nullptr;      /* pinned staging, words of either width */
  int state = 0, kind = 0;
  bool host_out = false;                                        /* the result passes through h_c   */
  struct WireJob *job = nullptr;                                /* whose rows the slot holds       */
  size_t row0 = 0, rows = 0;
  unsigned long long job_a = 0, job_b = 0, job_c = 0;
};

/* one host-buffer product on its way through the wire pipeline (nttb200.cu: wire_stream) */
struct WireJob {
  int32_t *c = nullptr;
  const int32_t *a = nullptr, *b = nullptr;
  size_t batch = 0;
  size_t next = 0;          /* rows handed to slots so far              */
  size_t busy = 0;          /* slots that hold rows of this job         */
  bool pinned = false, c_direct = false;
  bool wire = false;        /* goes through the wire pipeline (decided when the job is queued) */
  unsigned long long ticket = 0;
  int rc = 0;
  bool finished = false;
};
struct AsyncCtx;            /* worker thread + queue of nttb200_polymul_batch_async (nttb200.cu) */

struct LargeLane {
  cudaStream_t stream = nullptr;
  cudaEvent_t done = nullptr;
};

struct nttb200_plan {
  uint32_t n = 0, logn = 0, q = 0, psi = 0, omega = 0, flags = 0;
  uint32_t n_inv = 0;
  int device = 0;
  int arith = ARITH_LAZY;
  int kernel = PK_SMALL;
  bool plant = false;            /* half-word modulus: products run the Plantard kernel             */
  int sm_count = 0;
  ModQ m{};
  char desc[160] = {0};

  /* tables by role.  Product path: fwd = mixed_powers_rev, inv = inv_mixed_powers_rev
   * (omega_powers_rev / inv_omega_powers_rev for CYCLIC plans). */
  DevTable fwd_mixed, inv_mixed;          /* psi-merged                                 */
  DevTable fwd_plain, inv_plain;          /* omega_powers_rev / inv_omega_powers_rev    */
  DevTable fwd_invroot;                   /* inv_omega_powers_rev used as a FORWARD table */
  DevTable inv_fwdroot;                   /* omega_powers_rev used as an INVERSE table    */

  /* host-buffer path: mapped pinned buffer (a | b | c) for small calls, read and written by
   * the kernel over PCIe without any DMA */
  uint32_t *zc_host = nullptr, *zc_dev = nullptr;
  /* host-buffer path: ring of device staging slots */
  std::mutex mu;
  std::vector<HostSlot> slots;
  size_t slot_polys = 0;
  std::vector<WireSlot> wslots;
  size_t wire_polys = 0;           /* rows per wire chunk */
  unsigned long long wire16_chunks = 0, wire32_chunks = 0, wire_c32_rows = 0;      /* rows sent on each wire by the last call */
  std::atomic<AsyncCtx *> async{nullptr};    /* created by the first nttb200_polymul_batch_async */
  std::mutex async_init_mu;

  /* tail scheduler of the Plantard product kernel: ring of (next chunk, warps finished) pairs,
   * zero between launches; every launch that uses it is handed the next pair */
  unsigned long long *sched_ring = nullptr;
  mutable std::atomic<unsigned> sched_seq{0};

  /* large-n: internal stream lanes, each with scratch for scratch_polys polynomials */
  uint32_t *scratch = nullptr;
  size_t scratch_polys = 0;
  unsigned *flow_ctl = nullptr;    /* ticket + per-polynomial completion counters of the dataflow kernel */
  size_t flow_ctl_words = 0;
  int fused_clusters = -1;         /* resident clusters of the fused large-n kernel (-1: not asked yet, 0: not covered) */
  std::vector<LargeLane> lanes;
  cudaEvent_t fork = nullptr;
  cudaEvent_t scratch_done = nullptr;   /* end of the last call that used the scratch */
  std::mutex large_mu;
};

/* error plumbing (nttb200.cu) */
int nttb200_fail(int code, const char *fmt, ...);
#define NTT_CUDA(expr)                                                                    \
  do {                                                                                    \
    cudaError_t e_ = (expr);                                                              \
    if (e_ != cudaSuccess)                                                                \
      return nttb200_fail(NTTB200_ECUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(e_), \
                          __FILE__, __LINE__);                                            \
  } while (0)

void nttb200_count_launch(int k);
/* Is a launch on `st` that reads [a, a+abytes), [b, b+bbytes) and writes [c, c+cbytes) independent of
 * every launch of this library that may still be running on that stream?  Records the launch either
 * way (nttb200.cu). */
void nttb200_launch_forget(cudaStream_t st);
bool nttb200_launch_independent(cudaStream_t st, const void *a, size_t abytes, const void *b, size_t bbytes,
                                const void *c, size_t cbytes);

/* dispatchers implemented in the kernel translation units */
int launch_polymul_small(const nttb200_plan *P, uint32_t *c, const uint32_t *a, const uint32_t *b,
                         size_t batch, cudaStream_t st);
/* dir 0: forward CT std2rev with table `tab`; dir 1: inverse GS rev2std.
 * scale: 0 = none, 1 = multiply by n^-1 (inverse only) */
int launch_ntt_small(const nttb200_plan *P, const DevTable &tab, int dir, int scale, uint32_t *a,
                     size_t batch, cudaStream_t st);
int launch_generic_transform(uint32_t n, uint32_t logn, const ModQ &m, int dataflow, const uint2 *d_tab,
                             uint32_t *a, size_t batch, cudaStream_t st, int skip0 = 0);
int launch_pointwise(uint32_t *c, const uint32_t *a, const uint32_t *b, size_t count, const ModQ &m,
                     uint32_t r2, cudaStream_t st);
int launch_scale(uint32_t *a, const uint2 *d_tab, uint2 sc, uint32_t n, size_t count, const ModQ &m,
                 cudaStream_t st);
int launch_polymul_large(nttb200_plan *P, uint32_t *c, const uint32_t *a, const uint32_t *b,
                         size_t batch, cudaStream_t st);
int launch_ntt_large(nttb200_plan *P, const DevTable &tab, int dir, int scale, uint32_t *a, size_t batch,
                     cudaStream_t st);
int small_kernel_info(const nttb200_plan *P, int *regs, int *smem_bytes, int *blocks_per_sm);
int launch_polymul_small_plant(const nttb200_plan *P, uint32_t *c, const uint32_t *a, const uint32_t *b,
                               size_t batch, cudaStream_t st);
int small_kernel_info_plant(const nttb200_plan *P, int *regs, int *smem_bytes, int *blocks_per_sm);
int small_plant_signed(const nttb200_plan *P);      /* 1: products run polymul_splant_kernel (NTTB200_PLANT_SIGNED) */
int launch_ntt_small_plant(const nttb200_plan *P, const DevTable &tab, int dir, int scale, uint32_t *a,
                           size_t batch, cudaStream_t st);
int launch_polymul_small_plant_u16(const nttb200_plan *P, uint16_t *c, const uint16_t *a, const uint16_t *b,
                                   size_t batch, cudaStream_t st);
int launch_polymul_small_plant_u16in(const nttb200_plan *P, uint32_t *c, const uint16_t *a, const uint16_t *b,
                                     size_t batch, cudaStream_t st);
