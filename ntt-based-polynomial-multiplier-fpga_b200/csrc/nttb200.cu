/*
 * nttb200.cu -- runtime behind include/nttb200.h: plans, device tables, the device-resident
 * and host-buffer (pinned ring, H2D / kernel / D2H overlap) entry points.
 *
 * This is the layer that stands where the reference's FPGA transfer path stood
 * (Software_Hardware_Comunnicator/linux_app/NTT_PCIECommunicationv2.c:109-252): there, one
 * polynomial pair per DMA round trip; here, a batch streamed through CUDA streams.
 * There is no CPU compute path: without a CUDA device every computing call fails.
 */
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <unordered_map>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "host_tables.h"
#include "hostwire.h"
#include "ntt_generic.cuh"
#include "plan.h"

/* per-arithmetic-class dispatchers (small_*.cu) */
#define DECL_SMALL(name)                                                                          \
  int launch_polymul_small_##name(const nttb200_plan *, uint32_t *, const uint32_t *, const uint32_t *, \
                                  size_t, cudaStream_t);                                          \
  int launch_ntt_small_##name(const nttb200_plan *, const DevTable &, int, int, uint32_t *, size_t, \
                              cudaStream_t);                                                      \
  int small_kernel_info_##name(const nttb200_plan *, int *, int *, int *);
DECL_SMALL(lazy)
DECL_SMALL(harvey)
DECL_SMALL(canon)

/* ------------------------------------------------------------------------------------ */
/* errors, launch counter                                                                */
/* ------------------------------------------------------------------------------------ */
static thread_local char g_err[512] = "";
static thread_local int g_launches = 0;
static thread_local int g_device = -1;

int nttb200_fail(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
  return code;
}
void nttb200_count_launch(int k) { g_launches += k; }

extern "C" const char *nttb200_last_error(void) { return g_err; }
extern "C" int nttb200_version(void) { return NTTB200_VERSION; }
extern "C" int nttb200_last_launch_count(void) { return g_launches; }

extern "C" int nttb200_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return nttb200_fail(NTTB200_ECUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
  }
  return n;
}
extern "C" int nttb200_set_device(int device) {
  NTT_CUDA(cudaSetDevice(device));
  g_device = device;
  return 0;
}
extern "C" int nttb200_get_device(void) {
  int d = -1;
  if (cudaGetDevice(&d) != cudaSuccess) {
    cudaGetLastError();
    return NTTB200_ECUDA;
  }
  return d;
}

/* ------------------------------------------------------------------------------------ */
/* plans                                                                                 */
/* ------------------------------------------------------------------------------------ */
static int upload_table(DevTable &t, const std::vector<uint32_t> &w, uint32_t q, uint32_t plant_qinv = 0) {
  t.h.resize(w.size());
  for (size_t i = 0; i < w.size(); i++) t.h[i] = make_uint2(w[i], ht_shoup(w[i], q));
  NTT_CUDA(cudaMalloc(&t.d, t.h.size() * sizeof(uint2)));
  NTT_CUDA(cudaMemcpy(t.d, t.h.data(), t.h.size() * sizeof(uint2), cudaMemcpyHostToDevice));
  if (plant_qinv) {
    t.h1.resize(w.size());
    for (size_t i = 0; i < w.size(); i++) t.h1[i] = nttb200_plant_form(w[i], q, plant_qinv);
    NTT_CUDA(cudaMalloc(&t.d1, t.h1.size() * sizeof(uint32_t)));
    NTT_CUDA(cudaMemcpy(t.d1, t.h1.data(), t.h1.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    t.h2.resize(w.size());
    for (size_t i = 0; i < w.size(); i++) t.h2[i] = nttb200_plant_form_centred(w[i], q, plant_qinv);
    NTT_CUDA(cudaMalloc(&t.d2, t.h2.size() * sizeof(uint32_t)));
    NTT_CUDA(cudaMemcpy(t.d2, t.h2.data(), t.h2.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    /* -(w^2) 2^32 mod q, centred, entry by entry in the table's own layout: the group multiplication of
     * the incomplete transform reads the level it stops at (ntt_small_splant.cuh) */
    t.h3.resize(w.size() ? w.size() : 1);
    for (size_t j = 0; j < w.size(); j++) {
      const uint64_t z = (uint64_t)(w[j] % q) * (w[j] % q) % q;
      const uint64_t Z = (q - (z << 32) % q) % q;
      t.h3[j] = Z > q / 2 ? (uint32_t)Z - q : (uint32_t)Z;
    }
    NTT_CUDA(cudaMalloc(&t.d3, t.h3.size() * sizeof(uint32_t)));
    NTT_CUDA(cudaMemcpy(t.d3, t.h3.data(), t.h3.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
  }
  return 0;
}
static void free_table(DevTable &t) {
  if (t.d2) cudaFree(t.d2);
  t.d2 = nullptr;
  if (t.d3) cudaFree(t.d3);
  t.d3 = nullptr;
  if (t.d) cudaFree(t.d);
  if (t.d1) cudaFree(t.d1);
  t.d = nullptr;
  t.d1 = nullptr;
  t.h.clear();
  t.h1.clear();
}

static int pick_arith(uint32_t q, uint32_t logn) {
  /* LAZY: GS values double per stage (q 2^(logn+1) < 2^32) and the pointwise REDC needs
   * ((2 logn + 1) q)^2 < q 2^32 */
  const unsigned long long gs = (unsigned long long)q << (logn + 1);
  const unsigned long long ct = (unsigned long long)(2 * logn + 1);
  if (gs < (1ull << 32) && ct * ct * q < (1ull << 32)) return ARITH_LAZY;
  if (q < (1u << 30)) return ARITH_HARVEY;
  return ARITH_CANON;
}

extern "C" int nttb200_plan_create(nttb200_plan **out, uint32_t n, uint32_t q, uint32_t psi,
                                   uint32_t flags) {
  if (!out) return nttb200_fail(NTTB200_EPARAM, "plan pointer is NULL");
  *out = nullptr;
  if (n < 8 || (n & (n - 1)) || n > (1u << 17))
    return nttb200_fail(NTTB200_EPARAM, "n=%u: need a power of two in [8, 2^17]", n);
  if (q >= (1u << 31) || q < 3 || !nttb200_is_prime(q))
    return nttb200_fail(NTTB200_EPARAM, "q=%u: need an odd prime below 2^31", q);
  const bool cyclic = (flags & NTTB200_PLAN_CYCLIC) != 0;
  uint32_t omega;
  if (cyclic) {
    if ((q - 1) % n) return nttb200_fail(NTTB200_EPARAM, "q=%u: n=%u does not divide q-1", q, n);
    omega = psi ? psi : nttb200_find_omega(n, q);
    if (!omega || ht_powmod(omega, n / 2, q) != q - 1)
      return nttb200_fail(NTTB200_EPARAM, "omega=%u is not a primitive %u-th root mod %u", omega, n, q);
    psi = 0;
  } else {
    if ((uint64_t)(q - 1) % (2ull * n))
      return nttb200_fail(NTTB200_EPARAM,
                          "q=%u has no primitive %u-th root of unity (2n must divide q-1); "
                          "use NTTB200_PLAN_CYCLIC for the psi-free surface", q, 2 * n);
    if (!psi) psi = nttb200_find_psi(n, q);
    if (!psi || psi >= q || ht_powmod(psi, n, q) != q - 1)
      return nttb200_fail(NTTB200_EPARAM, "psi=%u is not a primitive %u-th root mod %u", psi, 2 * n, q);
    omega = (uint32_t)((uint64_t)psi * psi % q);
  }

  int dev = 0;
  int ndev = nttb200_device_count();
  if (ndev <= 0) return nttb200_fail(NTTB200_ECUDA, "no CUDA device (there is no CPU fallback)");
  if (g_device >= 0) NTT_CUDA(cudaSetDevice(g_device));
  NTT_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  NTT_CUDA(cudaGetDeviceProperties(&prop, dev));

  nttb200_plan *P = new (std::nothrow) nttb200_plan;
  if (!P) return nttb200_fail(NTTB200_ENOMEM, "out of host memory");
  P->n = n; P->logn = ht_log2(n); P->q = q; P->psi = psi; P->omega = omega; P->flags = flags;
  P->n_inv = ht_invmod(n % q, q);
  P->device = dev;
  P->sm_count = prop.multiProcessorCount;
  P->arith = pick_arith(q, P->logn);
  P->kernel = (P->logn <= 10) ? PK_SMALL : PK_LARGE;
  P->m.q = q; P->m.nq = 0u - q; P->m.q2 = 2u * q;
  {
    uint32_t inv = 1;                                   /* Newton: q^-1 mod 2^32 */
    for (int i = 0; i < 5; i++) inv *= 2u - q * inv;
    P->m.qinv = inv;
  }

  /* half-word moduli: the product kernel uses Plantard arithmetic (ntt_small_plant.cuh) */
  P->plant = (P->kernel == PK_SMALL) && q <= 12385u && !(flags & NTTB200_PLAN_NO_PLANTARD);
  const uint32_t pq = P->plant ? P->m.qinv : 0;
  std::vector<uint32_t> w(n);
  const uint32_t iomega = ht_invmod(omega, q);
  int rc = 0;
  ht_level_table(w.data(), n, q, 1, omega, 1);
  rc = rc ? rc : upload_table(P->fwd_plain, w, q, pq);
  rc = rc ? rc : upload_table(P->inv_fwdroot, w, q, pq);
  ht_level_table(w.data(), n, q, 1, iomega, 1);
  rc = rc ? rc : upload_table(P->inv_plain, w, q, pq);
  rc = rc ? rc : upload_table(P->fwd_invroot, w, q, pq);
  if (!cyclic) {
    ht_level_table(w.data(), n, q, psi, omega, 1);
    rc = rc ? rc : upload_table(P->fwd_mixed, w, q, pq);
    ht_level_table(w.data(), n, q, ht_invmod(psi, q), iomega, 1);
    rc = rc ? rc : upload_table(P->inv_mixed, w, q, pq);
  }
  if (rc) { nttb200_plan_destroy(P); return rc; }
  static const char *an[] = {"lazy", "harvey", "canon"};
  if (P->plant) {
    if (cudaMalloc(&P->sched_ring, NTTB200_SCHED_SLOTS * 2 * sizeof(unsigned long long)) != cudaSuccess ||
        cudaMemset(P->sched_ring, 0, NTTB200_SCHED_SLOTS * 2 * sizeof(unsigned long long)) != cudaSuccess) {
      cudaGetLastError();
      if (P->sched_ring) cudaFree(P->sched_ring);
      P->sched_ring = nullptr;                         /* static tile assignment only */
    }
  }
  snprintf(P->desc, sizeof P->desc, "n=%u q=%u %s=%u kernel=%s arith=%s%s device=%d sms=%d", n, q,
           cyclic ? "omega" : "psi", cyclic ? omega : psi,
           P->kernel == PK_SMALL ? "fused-small(regs+smem)" : "large(multi-pass)", an[P->arith],
           P->plant ? (small_plant_signed(P) ? "+plantard-signed(product)" : "+plantard(product)") : "", dev, P->sm_count);
  *out = P;
  return 0;
}

static void async_shutdown(nttb200_plan *P);
extern "C" void nttb200_plan_destroy(nttb200_plan *P) {
  if (!P) return;
  async_shutdown(P);                                   /* runs queued asynchronous products to their end */
  int cur = -1;
  cudaGetDevice(&cur);
  cudaSetDevice(P->device);
  for (auto &s : P->slots) {
    if (s.stream) { cudaStreamSynchronize(s.stream); cudaStreamDestroy(s.stream); }
    if (s.done) cudaEventDestroy(s.done);
    cudaFree(s.d_a); cudaFree(s.d_b); cudaFree(s.d_c);
  }
  for (auto &s : P->wslots) {
    if (s.stream) { cudaStreamSynchronize(s.stream); cudaStreamDestroy(s.stream); }
    if (s.done) cudaEventDestroy(s.done);
    cudaFree(s.d_a); cudaFree(s.d_b); cudaFree(s.d_c);
    if (s.h_a) cudaFreeHost(s.h_a);
  }
  if (P->scratch) cudaFree(P->scratch);
  if (P->flow_ctl) cudaFree(P->flow_ctl);
  if (P->sched_ring) cudaFree(P->sched_ring);
  if (P->zc_host) cudaFreeHost(P->zc_host);
  for (auto &ln : P->lanes) {
    if (ln.stream) { cudaStreamSynchronize(ln.stream); cudaStreamDestroy(ln.stream); }
    if (ln.done) cudaEventDestroy(ln.done);
  }
  if (P->fork) cudaEventDestroy(P->fork);
  if (P->scratch_done) cudaEventDestroy(P->scratch_done);
  free_table(P->fwd_mixed); free_table(P->inv_mixed);
  free_table(P->fwd_plain); free_table(P->inv_plain);
  free_table(P->fwd_invroot); free_table(P->inv_fwdroot);
  if (cur >= 0) cudaSetDevice(cur);
  delete P;
}
extern "C" uint32_t nttb200_plan_n(const nttb200_plan *P) { return P ? P->n : 0; }
extern "C" uint32_t nttb200_plan_q(const nttb200_plan *P) { return P ? P->q : 0; }
extern "C" uint32_t nttb200_plan_psi(const nttb200_plan *P) {
  return P ? ((P->flags & NTTB200_PLAN_CYCLIC) ? P->omega : P->psi) : 0;
}
extern "C" int nttb200_plan_device(const nttb200_plan *P) { return P ? P->device : -1; }
extern "C" const char *nttb200_plan_describe(const nttb200_plan *P) { return P ? P->desc : ""; }

/* ------------------------------------------------------------------------------------ */
/* Independent launches on one stream.                                                     */
/* The fused product kernel is launched with programmatic stream serialization: its CTAs  */
/* may start while the previous kernel of the stream drains.  Whether they may also READ   */
/* AND WRITE before that kernel has finished depends on the data: the library remembers,   */
/* per stream, the address ranges of its launches since the last one that waited; a new    */
/* launch is independent when it reads nothing they write and writes nothing they read or   */
/* write.  An independent launch skips the wait at its start (and waits at its end, so the  */
/* stream still completes in order); any other launch waits as before and becomes the new   */
/* base of the history.  Foreign work on the stream (other kernels, copies) never triggers  */
/* early, so it is fully ordered against both kinds.  NTTB200_PDL_NOWAIT=0 turns it off.    */
/* ------------------------------------------------------------------------------------ */
namespace {
struct RangeSet { uintptr_t rlo[2], rhi[2], wlo, whi; };
struct StreamHistory {
  cudaStream_t st = nullptr;
  bool used = false;
  int n = 0;
  unsigned long long stamp = 0;
  RangeSet e[8];
};
std::mutex g_hist_mu;
StreamHistory g_hist[16];
unsigned long long g_hist_clock = 0;
inline bool overlap(uintptr_t alo, uintptr_t ahi, uintptr_t blo, uintptr_t bhi) { return alo < bhi && blo < ahi; }
}  // namespace

bool nttb200_launch_independent(cudaStream_t st, const void *a, size_t abytes, const void *b, size_t bbytes,
                                const void *c, size_t cbytes) {
  static const bool enabled = [] { const char *e = getenv("NTTB200_PDL_NOWAIT"); return !e || atoi(e) != 0; }();
  RangeSet r;
  r.rlo[0] = (uintptr_t)a; r.rhi[0] = r.rlo[0] + abytes;
  r.rlo[1] = (uintptr_t)b; r.rhi[1] = r.rlo[1] + bbytes;
  r.wlo = (uintptr_t)c; r.whi = r.wlo + cbytes;
  std::lock_guard<std::mutex> lock(g_hist_mu);
  StreamHistory *h = nullptr, *spare = nullptr;
  for (auto &x : g_hist) {
    if (x.used && x.st == st) { h = &x; break; }
    if (!spare || (spare->used && (!x.used || x.stamp < spare->stamp))) spare = &x;   /* unused first, else oldest */
  }
  if (!h) { h = spare; h->st = st; h->used = true; h->n = 0; }
  h->stamp = ++g_hist_clock;
  bool indep = enabled && h->n > 0 && h->n < (int)(sizeof h->e / sizeof h->e[0]);
  for (int i = 0; indep && i < h->n; i++) {
    const RangeSet &p = h->e[i];
    for (int k = 0; k < 2; k++) {
      if (overlap(r.rlo[k], r.rhi[k], p.wlo, p.whi)) indep = false;          /* read after write  */
      if (overlap(r.wlo, r.whi, p.rlo[k], p.rhi[k])) indep = false;          /* write after read  */
    }
    if (overlap(r.wlo, r.whi, p.wlo, p.whi)) indep = false;                  /* write after write */
  }
  if (!indep) h->n = 0;                                 /* this launch waits for everything before it */
  h->e[h->n++] = r;
  return indep;
}

/* Kernels of this library that trigger their dependents early but are NOT recorded above (the large-n
 * pipeline writes its result from several launches and lanes) forget the stream's history instead: the
 * next recorded launch then waits, whatever it reads. */
void nttb200_launch_forget(cudaStream_t st) {
  std::lock_guard<std::mutex> lock(g_hist_mu);
  for (auto &x : g_hist)
    if (x.used && x.st == st) x.n = 0;
}

/* ------------------------------------------------------------------------------------ */
/* dispatch                                                                              */
/* ------------------------------------------------------------------------------------ */
int launch_polymul_small(const nttb200_plan *P, uint32_t *c, const uint32_t *a, const uint32_t *b,
                         size_t batch, cudaStream_t st) {
  /* the Plantard kernel prefetches with 16-byte cp.async: operands must be 16-byte aligned */
  if (P->plant && (((uintptr_t)a | (uintptr_t)b | (uintptr_t)c) & 15u) == 0)
    return launch_polymul_small_plant(P, c, a, b, batch, st);
  switch (P->arith) {
    case ARITH_LAZY: return launch_polymul_small_lazy(P, c, a, b, batch, st);
    case ARITH_HARVEY: return launch_polymul_small_harvey(P, c, a, b, batch, st);
    default: return launch_polymul_small_canon(P, c, a, b, batch, st);
  }
}
int launch_ntt_small(const nttb200_plan *P, const DevTable &tab, int dir, int scale, uint32_t *a,
                     size_t batch, cudaStream_t st) {
  /* (the forward Plantard kernel prefetches with 16-byte cp.async: rows must be 16-byte aligned) */
  if (P->plant && tab.d1 && ((uintptr_t)a & 15u) == 0 && !getenv("NTTB200_NTT_SHOUP"))
    return launch_ntt_small_plant(P, tab, dir, scale, a, batch, st);
  switch (P->arith) {
    case ARITH_LAZY: return launch_ntt_small_lazy(P, tab, dir, scale, a, batch, st);
    case ARITH_HARVEY: return launch_ntt_small_harvey(P, tab, dir, scale, a, batch, st);
    default: return launch_ntt_small_canon(P, tab, dir, scale, a, batch, st);
  }
}
int small_kernel_info(const nttb200_plan *P, int *regs, int *smem_bytes, int *blocks_per_sm) {
  if (P->plant) return small_kernel_info_plant(P, regs, smem_bytes, blocks_per_sm);
  switch (P->arith) {
    case ARITH_LAZY: return small_kernel_info_lazy(P, regs, smem_bytes, blocks_per_sm);
    case ARITH_HARVEY: return small_kernel_info_harvey(P, regs, smem_bytes, blocks_per_sm);
    default: return small_kernel_info_canon(P, regs, smem_bytes, blocks_per_sm);
  }
}

static int grid_1d(unsigned long long work, int threads, int sm_count) {
  unsigned long long blocks = (work + threads - 1) / threads;
  unsigned long long cap = (unsigned long long)sm_count * 16;
  return (int)std::max<unsigned long long>(1, std::min(blocks, cap));
}
static int current_sms() {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess)
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}

int launch_generic_transform(uint32_t n, uint32_t logn, const ModQ &m, int dataflow, const uint2 *d_tab,
                             uint32_t *a, size_t batch, cudaStream_t st, int skip0) {
  using namespace nttb200;
  const unsigned long long pairs = (unsigned long long)batch * (n / 2);
  const int grid = grid_1d(pairs, 256, current_sms());
  const bool descending = (dataflow == DF_CT_STD2REV || dataflow == DF_GS_STD2REV);
  for (uint32_t s = 0; s < logn; s++) {
    const uint32_t lh = descending ? (logn - 1 - s) : s;
    const uint32_t half = 1u << lh;
    switch (dataflow) {
      case DF_CT_STD2REV: generic_stage_kernel<DF_CT_STD2REV><<<grid, 256, 0, st>>>(a, d_tab, n, logn, half, lh, pairs, m, skip0); break;
      case DF_GS_REV2STD: generic_stage_kernel<DF_GS_REV2STD><<<grid, 256, 0, st>>>(a, d_tab, n, logn, half, lh, pairs, m, skip0); break;
      case DF_CT_REV2STD: generic_stage_kernel<DF_CT_REV2STD><<<grid, 256, 0, st>>>(a, d_tab, n, logn, half, lh, pairs, m, skip0); break;
      case DF_GS_STD2REV: generic_stage_kernel<DF_GS_STD2REV><<<grid, 256, 0, st>>>(a, d_tab, n, logn, half, lh, pairs, m, skip0); break;
      default: return nttb200_fail(NTTB200_EPARAM, "unknown dataflow %d", dataflow);
    }
    nttb200_count_launch(1);
  }
  NTT_CUDA(cudaGetLastError());
  return 0;
}
int launch_pointwise(uint32_t *c, const uint32_t *a, const uint32_t *b, size_t count, const ModQ &m,
                     uint32_t r2, cudaStream_t st) {
  nttb200::pointwise_kernel<<<grid_1d(count, 256, current_sms()), 256, 0, st>>>(c, a, b, count, m, r2);
  nttb200_count_launch(1);
  NTT_CUDA(cudaGetLastError());
  return 0;
}
int launch_scale(uint32_t *a, const uint2 *d_tab, uint2 sc, uint32_t n, size_t count, const ModQ &m,
                 cudaStream_t st) {
  nttb200::scale_kernel<<<grid_1d(count, 256, current_sms()), 256, 0, st>>>(a, d_tab, sc, n, count, m);
  nttb200_count_launch(1);
  NTT_CUDA(cudaGetLastError());
  return 0;
}

static uint2 pair_of(uint64_t w, uint32_t q) {
  w %= q;
  return make_uint2((uint32_t)w, ht_shoup((uint32_t)w, q));
}

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

/* NTTB200_PLAN_CHECK_RANGE: the reference leaves out-of-range inputs undefined (R/NTT/ntt256.h:82-83
 * "must contain elements in the range [0 .. Q-1]"); a plan created with the flag looks first.
 * Host arrays are scanned where they lie; device arrays by a kernel whose verdict the call waits
 * for (which makes the _dev call synchronous -- it is a debugging aid, not the timed path). */
static int check_range_host(const nttb200_plan *P, const int32_t *x, size_t count, const char *what) {
  const uint32_t q = P->q;
  uint32_t bad = 0;
  const uint32_t *u = (const uint32_t *)x;
  for (size_t i = 0; i < count; i++) bad |= (uint32_t)(u[i] >= q);
  if (!bad) return 0;
  size_t i = 0;
  while (u[i] < q) i++;
  return nttb200_fail(NTTB200_ERANGE, "%s[%zu] = %d is outside [0, %u)", what, i, x[i], q);
}
__global__ void __launch_bounds__(256)
range_check_kernel(const uint32_t *a, const uint32_t *b, unsigned long long count, uint32_t q,
                   unsigned long long *first_bad) {
  unsigned long long gid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned long long gstride = (unsigned long long)gridDim.x * blockDim.x;
  for (; gid < count; gid += gstride) {
    if (a[gid] >= q) atomicMin(first_bad, gid);
    if (b && b[gid] >= q) atomicMin(first_bad, count + gid);
  }
}
static int check_range_dev(const nttb200_plan *P, const uint32_t *a, const uint32_t *b, size_t count,
                           cudaStream_t st) {
  unsigned long long *flag = nullptr, h = ~0ull;
  NTT_CUDA(cudaMalloc(&flag, sizeof h));
  cudaError_t e = cudaMemcpyAsync(flag, &h, sizeof h, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) {
    range_check_kernel<<<grid_1d(count, 256, P->sm_count), 256, 0, st>>>(a, b, count, P->q, flag);
    nttb200_count_launch(1);
    e = cudaMemcpyAsync(&h, flag, sizeof h, cudaMemcpyDeviceToHost, st);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(flag);
  if (e != cudaSuccess) return nttb200_fail(NTTB200_ECUDA, "range check: %s", cudaGetErrorString(e));
  if (h == ~0ull) return 0;
  return nttb200_fail(NTTB200_ERANGE, "%s[%llu] is outside [0, %u)", h >= count ? "b" : "a",
                      h >= count ? h - count : h, P->q);
}

static int polymul_dev(nttb200_plan *P, uint32_t *c, const uint32_t *a, const uint32_t *b, size_t batch,
                       cudaStream_t st) {
  if (batch == 0) return 0;
  if (P->kernel == PK_SMALL) return launch_polymul_small(P, c, a, b, batch, st);
  return launch_polymul_large(P, c, a, b, batch, st);
}

static int transform_dev(nttb200_plan *P, int transform, uint32_t *a, size_t batch, cudaStream_t st) {
  if (batch == 0) return 0;
  const bool cyclic = (P->flags & NTTB200_PLAN_CYCLIC) != 0;
  const DevTable *tab = nullptr;
  int dir = 0, scale = 0;
  switch (transform) {
    case NTTB200_NTT_STD2REV: tab = &P->fwd_plain; dir = 0; break;
    case NTTB200_MULNTT_STD2REV: tab = &P->fwd_mixed; dir = 0; break;
    case NTTB200_INTT_REV2STD: tab = &P->inv_plain; dir = 1; break;
    case NTTB200_INTTMUL_REV2STD: tab = &P->inv_mixed; dir = 1; break;
    case NTTB200_INTT_REV2STD_SCALED: tab = &P->inv_plain; dir = 1; scale = 1; break;
    case NTTB200_INTTMUL_REV2STD_SCALED: tab = &P->inv_mixed; dir = 1; scale = 1; break;
    case NTTB200_INTT_STD2REV: tab = &P->fwd_invroot; dir = 0; break;
    case NTTB200_NTT_REV2STD: tab = &P->inv_fwdroot; dir = 1; break;
    default: return nttb200_fail(NTTB200_EPARAM, "unknown transform %d", transform);
  }
  if (cyclic && (tab == &P->fwd_mixed || tab == &P->inv_mixed))
    return nttb200_fail(NTTB200_EPARAM, "psi-merged transforms need a negacyclic plan");
  if (P->kernel == PK_SMALL) return launch_ntt_small(P, *tab, dir, scale, a, batch, st);
  return launch_ntt_large(P, *tab, dir, scale, a, batch, st);
}

/* ------------------------------------------------------------------------------------ */
/* large n: multi-pass kernels (ntt_large.cuh), batch processed in chunks whose scratch     */
/* (a', b' between the passes) stays L2-resident                                          */
/* ------------------------------------------------------------------------------------ */
static int env_int(const char *name, int dflt, int lo, int hi);
#define DECL_LARGE(name)                                                                          \
  int launch_polymul_large_chunk_##name(const nttb200_plan *, uint32_t *, const uint32_t *,       \
                                        const uint32_t *, uint32_t *, uint32_t *, size_t, cudaStream_t, int); \
  int launch_ntt_large_##name(const nttb200_plan *, const DevTable &, int, int, uint32_t *, size_t, \
                              cudaStream_t);                                                        \
  int large_fused_clusters_##name(const nttb200_plan *, int *);                                     \
  int large_flow_slots_##name(void);                                                                \
  int launch_polymul_large_flow_##name(const nttb200_plan *, uint32_t *, const uint32_t *,          \
                                       const uint32_t *, uint32_t *, unsigned *, size_t, cudaStream_t); \
  int launch_polymul_large_fused_##name(const nttb200_plan *, uint32_t *, const uint32_t *,         \
                                        const uint32_t *, uint32_t *, unsigned, size_t, cudaStream_t);
DECL_LARGE(lazy)
DECL_LARGE(harvey)
DECL_LARGE(canon)

/* The batch is cut into chunks; chunk i runs its three kernels on internal stream i % LANES
 * with its own scratch (a', b'; c' over a'), so the column pass of one chunk overlaps the row
 * pass of the previous one (tails and launch gaps are filled) while each chunk's scratch is
 * small enough to stay in the 126 MB L2 between its passes.
 * NTTB200_LARGE_SCRATCH_MB = scratch bytes per lane (default 32), NTTB200_LARGE_LANES (default 3). */
static size_t large_scratch_budget() {
  static size_t v = (size_t)env_int("NTTB200_LARGE_SCRATCH_MB", 32, 1, 65536) << 20;
  return v;
}
/* 0: three launches per batch chunk; 1: one persistent cluster kernel (n = 2^15, 2^16);
 * 2: one persistent dataflow kernel (n = 2^16).  Read per call: tests switch it. */
static int large_mode() { return env_int("NTTB200_LARGE_FUSED", 0, 0, 2); }
static int large_lanes() { static int v = env_int("NTTB200_LARGE_LANES", 3, 1, 8); return v; }

static bool ranges_overlap(const uint32_t *x, const uint32_t *y, size_t words) {
  const uintptr_t a = (uintptr_t)x, b = (uintptr_t)y, n = words * sizeof(uint32_t);
  return a < b + n && b < a + n;
}

static int ensure_scratch(nttb200_plan *P, size_t polys, int lanes) {
  if (P->scratch_polys >= polys && (int)P->lanes.size() >= lanes) return 0;
  if (P->scratch) cudaFree(P->scratch);
  P->scratch = nullptr; P->scratch_polys = 0;
  NTT_CUDA(cudaMalloc(&P->scratch, (size_t)lanes * polys * P->n * 2 * sizeof(uint32_t)));
  P->scratch_polys = polys;
  while ((int)P->lanes.size() < lanes) {
    LargeLane ln;
    NTT_CUDA(cudaStreamCreateWithFlags(&ln.stream, cudaStreamNonBlocking));
    NTT_CUDA(cudaEventCreateWithFlags(&ln.done, cudaEventDisableTiming));
    P->lanes.push_back(ln);
  }
  if (!P->fork) NTT_CUDA(cudaEventCreateWithFlags(&P->fork, cudaEventDisableTiming));
  return 0;
}

int launch_polymul_large(nttb200_plan *P, uint32_t *c, const uint32_t *a, const uint32_t *b, size_t batch,
                         cudaStream_t st) {
  /* The scratch and the lanes belong to the plan: calls from several host threads / on several
   * streams are serialised on the device through `scratch_done` (recorded at the end of every
   * call, awaited at the start of the next), and on the host by large_mu. */
  std::lock_guard<std::mutex> lock(P->large_mu);
  nttb200_launch_forget(st);
  if (!P->scratch_done) NTT_CUDA(cudaEventCreateWithFlags(&P->scratch_done, cudaEventDisableTiming));
  else NTT_CUDA(cudaStreamWaitEvent(st, P->scratch_done, 0));
  /* n = 2^15, 2^16: one persistent cluster kernel for the whole batch (ntt_large_fused.cuh); its
   * scratch is (resident clusters) x 2n words and lives in L2.  NTTB200_LARGE_FUSED=0 keeps the
   * three-launch pipeline below (the comparison point, and the path of every other n). */
  const int mode = large_mode();
  if (mode == 2 && P->logn == 16) {                       /* the ticket dataflow kernel */
    const int slots = large_flow_slots_canon();
    int rcf = ensure_scratch(P, (size_t)slots, 1);
    if (rcf) return rcf;
    const size_t words = 1 + 3 * batch;
    if (P->flow_ctl_words < words) {
      if (P->flow_ctl) cudaFree(P->flow_ctl);
      P->flow_ctl = nullptr; P->flow_ctl_words = 0;
      NTT_CUDA(cudaMalloc(&P->flow_ctl, words * sizeof(unsigned)));
      P->flow_ctl_words = words;
    }
    NTT_CUDA(cudaMemsetAsync(P->flow_ctl, 0, words * sizeof(unsigned), st));
    switch (P->arith) {
      case ARITH_LAZY: rcf = launch_polymul_large_flow_lazy(P, c, a, b, P->scratch, P->flow_ctl, batch, st); break;
      case ARITH_HARVEY: rcf = launch_polymul_large_flow_harvey(P, c, a, b, P->scratch, P->flow_ctl, batch, st); break;
      default: rcf = launch_polymul_large_flow_canon(P, c, a, b, P->scratch, P->flow_ctl, batch, st); break;
    }
    if (rcf) return rcf;
    NTT_CUDA(cudaEventRecord(P->scratch_done, st));
    return 0;
  }
  if (mode >= 1) {
    if (P->fused_clusters < 0) {
      int nc = 0, rcq;
      switch (P->arith) {
        case ARITH_LAZY: rcq = large_fused_clusters_lazy(P, &nc); break;
        case ARITH_HARVEY: rcq = large_fused_clusters_harvey(P, &nc); break;
        default: rcq = large_fused_clusters_canon(P, &nc); break;
      }
      if (rcq) { cudaGetLastError(); nc = 0; }
      P->fused_clusters = nc;
    }
    if (P->fused_clusters > 0) {
      const unsigned clusters = (unsigned)std::min<size_t>(batch, (size_t)P->fused_clusters);
      int rcf = ensure_scratch(P, (size_t)P->fused_clusters, 1);
      if (rcf) return rcf;
      switch (P->arith) {
        case ARITH_LAZY: rcf = launch_polymul_large_fused_lazy(P, c, a, b, P->scratch, clusters, batch, st); break;
        case ARITH_HARVEY: rcf = launch_polymul_large_fused_harvey(P, c, a, b, P->scratch, clusters, batch, st); break;
        default: rcf = launch_polymul_large_fused_canon(P, c, a, b, P->scratch, clusters, batch, st); break;
      }
      if (rcf) return rcf;
      NTT_CUDA(cudaEventRecord(P->scratch_done, st));
      return 0;
    }
  }
  const size_t chunk = std::max<size_t>(1, std::min<size_t>(batch, large_scratch_budget() / (P->n * 8ull)));
  const size_t nchunks = (batch + chunk - 1) / chunk;
  const int lanes = (int)std::min<size_t>((size_t)large_lanes(), nchunks);
  int rc = ensure_scratch(P, chunk, lanes);
  if (rc) return rc;
  if (lanes > 1) {                      /* fork: the lanes start after everything queued on st */
    NTT_CUDA(cudaEventRecord(P->fork, st));
    for (int l = 0; l < lanes; l++) NTT_CUDA(cudaStreamWaitEvent(P->lanes[l].stream, P->fork, 0));
  }
  size_t k = 0;
  for (size_t done = 0; done < batch; done += chunk, k++) {
    const size_t nb = std::min(chunk, batch - done);
    const size_t o = done * P->n;
    const int l = (int)(k % lanes);
    cudaStream_t ls = lanes > 1 ? P->lanes[l].stream : st;
    uint32_t *ta = P->scratch + (size_t)l * chunk * P->n * 2, *tb = ta + chunk * P->n;
    /* From a lane's second chunk on, the kernel before this chunk's column pass is the same call's
     * own last pass of an EARLIER chunk: unless the result array overlaps an operand, this chunk's
     * operands cannot be its output, so they may be read while it still runs */
    const int early = (k >= (size_t)lanes && !ranges_overlap(c, a, batch * P->n) && !ranges_overlap(c, b, batch * P->n)) ? 1 : 0;
    switch (P->arith) {
      case ARITH_LAZY: rc = launch_polymul_large_chunk_lazy(P, c + o, a + o, b + o, ta, tb, nb, ls, early); break;
      case ARITH_HARVEY: rc = launch_polymul_large_chunk_harvey(P, c + o, a + o, b + o, ta, tb, nb, ls, early); break;
      default: rc = launch_polymul_large_chunk_canon(P, c + o, a + o, b + o, ta, tb, nb, ls, early); break;
    }
    if (rc) return rc;
  }
  if (lanes > 1) {                      /* join */
    for (int l = 0; l < lanes; l++) {
      NTT_CUDA(cudaEventRecord(P->lanes[l].done, P->lanes[l].stream));
      NTT_CUDA(cudaStreamWaitEvent(st, P->lanes[l].done, 0));
    }
  }
  NTT_CUDA(cudaEventRecord(P->scratch_done, st));
  return 0;
}

int launch_ntt_large(nttb200_plan *P, const DevTable &tab, int dir, int scale, uint32_t *a, size_t batch,
                     cudaStream_t st) {
  nttb200_launch_forget(st);
  switch (P->arith) {
    case ARITH_LAZY: return launch_ntt_large_lazy(P, tab, dir, scale, a, batch, st);
    case ARITH_HARVEY: return launch_ntt_large_harvey(P, tab, dir, scale, a, batch, st);
    default: return launch_ntt_large_canon(P, tab, dir, scale, a, batch, st);
  }
}

/* ------------------------------------------------------------------------------------ */
/* public: device-resident                                                               */
/* ------------------------------------------------------------------------------------ */
extern "C" int nttb200_polymul_batch_dev(nttb200_plan *P, int32_t *c, const int32_t *a, const int32_t *b,
                                         size_t batch, void *stream) {
  if (!P || !c || !a || !b) return nttb200_fail(NTTB200_EPARAM, "NULL argument");
  g_launches = 0;
  DeviceGuard guard(P->device);
  if ((P->flags & NTTB200_PLAN_CHECK_RANGE) && batch) {
    const int rc = check_range_dev(P, (const uint32_t *)a, (const uint32_t *)b, batch * P->n, (cudaStream_t)stream);
    if (rc) return rc;
  }
  return polymul_dev(P, (uint32_t *)c, (const uint32_t *)a, (const uint32_t *)b, batch,
                     (cudaStream_t)stream);
}

extern "C" int nttb200_ntt_batch_dev(nttb200_plan *P, int transform, int32_t *a, size_t batch,
                                     void *stream) {
  if (!P || !a) return nttb200_fail(NTTB200_EPARAM, "NULL argument");
  /* the transform kernels move rows with 128-bit accesses */
  if (((uintptr_t)a & 15u) != 0) return nttb200_fail(NTTB200_EPARAM, "a_dev must be 16-byte aligned");
  g_launches = 0;
  DeviceGuard guard(P->device);
  if ((P->flags & NTTB200_PLAN_CHECK_RANGE) && batch) {
    const int rc = check_range_dev(P, (const uint32_t *)a, nullptr, batch * P->n, (cudaStream_t)stream);
    if (rc) return rc;
  }
  return transform_dev(P, transform, (uint32_t *)a, batch, (cudaStream_t)stream);
}

/* ------------------------------------------------------------------------------------ */
/* public: host buffers.  Ring of NSLOT device staging slots, each with its own stream:    */
/* H2D(a,b) -> kernel -> D2H(c) of slot k overlaps the copies of slots k+-1.  cudaMemcpyAsync */
/* from pinned memory (nttb200_host_alloc) is a true async DMA; pageable memory also works  */
/* (the driver stages it).                                                                */
/* ------------------------------------------------------------------------------------ */
static int env_int(const char *name, int dflt, int lo, int hi) {
  const char *e = getenv(name);
  if (!e) return dflt;
  int v = atoi(e);
  return v < lo ? lo : (v > hi ? hi : v);
}
/* ring depth and bytes per operand per slot (NTTB200_NSLOT / NTTB200_SLOT_MB: tuning knobs) */
static int nslot() { static int v = env_int("NTTB200_NSLOT", 3, 1, 8); return v; }
#define NSLOT nslot()

static const size_t ZC_BYTES = 64u << 10;    /* per operand, zero-copy path of small calls */
static const size_t WIRE_MIN_WORDS = 1u << 19;   /* per operand: smaller calls keep the plain DMA ring */

static int ensure_slots(nttb200_plan *P, bool need_b) {
  if (!P->slots.empty()) return 0;
  if (!P->zc_host && P->kernel == PK_SMALL) {
    void *h = nullptr, *d = nullptr;
    if (cudaHostAlloc(&h, 3 * ZC_BYTES, cudaHostAllocMapped | cudaHostAllocPortable) == cudaSuccess &&
        cudaHostGetDevicePointer(&d, h, 0) == cudaSuccess) {
      P->zc_host = (uint32_t *)h;
      P->zc_dev = (uint32_t *)d;
    } else {
      if (h) cudaFreeHost(h);
      cudaGetLastError();                          /* the DMA ring below serves small calls too */
    }
  }
  const size_t target_bytes = (size_t)env_int("NTTB200_SLOT_MB", 16, 1, 256) << 20;   /* per operand per slot */
  const size_t polys = std::max<size_t>(1, target_bytes / (P->n * sizeof(uint32_t)));
  /* built aside and moved into the plan only when every allocation succeeded: a half-built ring
   * must not look initialised to the next call */
  std::vector<HostSlot> ring(NSLOT);
  cudaError_t e = cudaSuccess;
  for (auto &s : ring) {
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaMalloc(&s.d_a, polys * P->n * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc(&s.d_b, polys * P->n * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc(&s.d_c, polys * P->n * sizeof(uint32_t));
  }
  if (e != cudaSuccess) {
    for (auto &s : ring) {
      if (s.stream) cudaStreamDestroy(s.stream);
      if (s.done) cudaEventDestroy(s.done);
      cudaFree(s.d_a); cudaFree(s.d_b); cudaFree(s.d_c);
    }
    cudaGetLastError();
    return nttb200_fail(e == cudaErrorMemoryAllocation ? NTTB200_ENOMEM : NTTB200_ECUDA,
                        "staging ring: %s", cudaGetErrorString(e));
  }
  P->slot_polys = polys;
  P->slots = std::move(ring);
  (void)need_b;
  return 0;
}

/* ------------------------------------------------------------------------------------ */
/* Host buffers, half-word moduli: the WIRE pipeline.                                      */
/* The host-buffer path is bound by the PCIe link (2n words in, n out per product), so    */
/* chunks travel as 16-bit words where the host can narrow them fast enough (hostwire.c:  */
/* worker threads, pinned staging) and as the caller's 32-bit words otherwise:            */
/*   wire16 chunk: narrow a,b (CPU pool) -> H2D 2x2 B/word -> u16 kernel -> D2H 2 B/word   */
/*                 -> widen into c (CPU pool)                                             */
/*   wire32 chunk: H2D straight from the caller's buffers -> kernel -> D2H straight into c */
/* Measured on the GPU box (16 host cores, c2): 32-bit wire 22.2 M         */
/* polymul/s, 16-bit wire 29.5 M -- the host's memory system (narrowing reads 2n words per     */
/* product at ~90 GB/s) is then what binds, not the link.  Two variants are kept behind knobs  */
/* because they measured slower on that box: NTTB200_WIRE_AHEAD=k sends a chunk as 32-bit      */
/* words when the pool is already k chunks behind (27.4 M at k=2); NTTB200_WIRE_C32=1 lets the */
/* kernel write int32 result rows that one D2H copies straight into a pinned c, sparing the    */
/* pool the widening (26.8 M: the extra upstream traffic slows the H2D reads).  Pageable       */
/* caller buffers always go wire16: the pool is a faster stager than the driver's pageable     */
/* path.  A chunk in which some word does not fit 16 bits is re-sent wire32, so results never  */
/* depend on the wire.  NTTB200_WIRE=32|16|auto, NTTB200_WIRE_KWORDS (chunk),                  */
/* NTTB200_WIRE_SLOTS are tuning knobs.                                                        */
/* ------------------------------------------------------------------------------------ */
enum { WS_FREE = 0, WS_STAGING = 1, WS_INFLIGHT = 2, WS_WIDENING = 3 };
/* host-buffer calls running side by side in this process (nttb200_multi_polymul_batch: one per GPU) */
static std::atomic<int> g_wire_share{1};

static int wire_mode() {                               /* read per call: tests switch it */
  const char *e = getenv("NTTB200_WIRE");
  return (e && !strcmp(e, "32")) ? 32 : (e && !strcmp(e, "16")) ? 16 : 0;
}

static int ensure_wire_slots(nttb200_plan *P) {
  if (!P->wslots.empty()) return 0;
  const size_t words = (size_t)env_int("NTTB200_WIRE_KWORDS", 1024, 16, 65536) << 10;   /* per operand */
  const size_t polys = std::max<size_t>(1, words / P->n);
  const size_t w = polys * P->n;
  /* as ensure_slots: nothing is published until every slot is complete (a worker of the host
   * pool handed a slot without staging memory would write through NULL) */
  std::vector<WireSlot> ring(env_int("NTTB200_WIRE_SLOTS", 6, 2, 16));
  cudaError_t e = cudaSuccess;
  for (auto &s : ring) {
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaMalloc(&s.d_a, w * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc(&s.d_b, w * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc(&s.d_c, w * sizeof(uint32_t));
    void *h = nullptr;                                 /* staging holds words of either width */
    if (e == cudaSuccess) e = cudaHostAlloc(&h, 3 * w * sizeof(uint32_t), cudaHostAllocPortable);
    if (e == cudaSuccess) {
      s.h_a = (uint32_t *)h;
      s.h_b = s.h_a + w;
      s.h_c = s.h_b + w;
    }
  }
  if (e != cudaSuccess) {
    for (auto &s : ring) {
      if (s.stream) cudaStreamDestroy(s.stream);
      if (s.done) cudaEventDestroy(s.done);
      cudaFree(s.d_a); cudaFree(s.d_b); cudaFree(s.d_c);
      if (s.h_a) cudaFreeHost(s.h_a);
    }
    cudaGetLastError();
    return nttb200_fail(e == cudaErrorMemoryAllocation ? NTTB200_ENOMEM : NTTB200_ECUDA,
                        "wire pipeline slots: %s", cudaGetErrorString(e));
  }
  P->wire_polys = polys;
  P->wslots = std::move(ring);
  return 0;
}

static bool is_pinned(const void *p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged;
}

/* how a chunk crosses the link */
enum { WK_16 = 0,        /* narrowed by the pool, 16-bit words both ways (or int32 result rows: c_direct) */
       WK_32_DIRECT = 1, /* DMA straight from / into the caller's pinned buffers                        */
       WK_32_STAGED = 2  /* 32-bit words copied by the pool through pinned staging (pageable callers)   */ };

/* 32-bit words: from the caller's pinned buffers, or from staging the pool has filled */
static int wire_send32(nttb200_plan *P, WireSlot &s, bool staged) {
  const size_t n = P->n, bytes = s.rows * n * sizeof(uint32_t);
  int32_t *c = s.job->c;
  const int32_t *a = s.job->a, *b = s.job->b;
  const void *sa = staged ? (const void *)s.h_a : (const void *)(a + s.row0 * n);
  const void *sb = staged ? (const void *)s.h_b : (const void *)(b + s.row0 * n);
  void *dc = staged ? (void *)s.h_c : (void *)(c + s.row0 * n);
  NTT_CUDA(cudaMemcpyAsync(s.d_a, sa, bytes, cudaMemcpyHostToDevice, s.stream));
  NTT_CUDA(cudaMemcpyAsync(s.d_b, sb, bytes, cudaMemcpyHostToDevice, s.stream));
  int rc = launch_polymul_small(P, s.d_c, s.d_a, s.d_b, s.rows, s.stream);
  if (rc) return rc;
  NTT_CUDA(cudaMemcpyAsync(dc, s.d_c, bytes, cudaMemcpyDeviceToHost, s.stream));
  NTT_CUDA(cudaEventRecord(s.done, s.stream));
  s.kind = staged ? WK_32_STAGED : WK_32_DIRECT;
  s.host_out = staged;
  s.state = WS_INFLIGHT;
  P->wire32_chunks += s.rows;
  return 0;
}
/* 16-bit operands; the result comes back as 16-bit rows into pinned staging that the pool widens,
 * or (c_direct) as the caller's int32 rows written by the kernel's own 32-bit stores and copied
 * straight into a pinned c */
static int wire_send16(nttb200_plan *P, WireSlot &s) {
  const size_t n = P->n, bytes = s.rows * n * sizeof(uint16_t);
  int32_t *c = s.job->c;
  const bool c_direct = s.job->c_direct;
  NTT_CUDA(cudaMemcpyAsync(s.d_a, s.h_a, bytes, cudaMemcpyHostToDevice, s.stream));
  NTT_CUDA(cudaMemcpyAsync(s.d_b, s.h_b, bytes, cudaMemcpyHostToDevice, s.stream));
  int rc;
  if (c_direct) {
    rc = launch_polymul_small_plant_u16in(P, s.d_c, (const uint16_t *)s.d_a, (const uint16_t *)s.d_b, s.rows,
                                          s.stream);
    if (rc) return rc;
    NTT_CUDA(cudaMemcpyAsync(c + s.row0 * n, s.d_c, 2 * bytes, cudaMemcpyDeviceToHost, s.stream));
  } else {
    rc = launch_polymul_small_plant_u16(P, (uint16_t *)s.d_c, (const uint16_t *)s.d_a, (const uint16_t *)s.d_b,
                                        s.rows, s.stream);
    if (rc) return rc;
    NTT_CUDA(cudaMemcpyAsync(s.h_c, s.d_c, bytes, cudaMemcpyDeviceToHost, s.stream));
  }
  NTT_CUDA(cudaEventRecord(s.done, s.stream));
  s.kind = WK_16;
  s.host_out = !c_direct;
  s.state = WS_INFLIGHT;
  P->wire16_chunks += s.rows;
  if (c_direct) P->wire_c32_rows += s.rows;
  return 0;
}
/* queue the pool jobs that fill the slot's staging with the chunk's operands */
static void wire_stage_in(nttb200_plan *P, WireSlot &s, int kind, uint32_t *mask) {
  const size_t n = P->n, words = s.rows * n;
  const int32_t *a = s.job->a, *b = s.job->b;
  if (kind == WK_16) {
    *mask = 0;
    s.job_a = nttb200_wire_post_narrow((uint16_t *)s.h_a, a + s.row0 * n, words, mask);
    s.job_b = nttb200_wire_post_narrow((uint16_t *)s.h_b, b + s.row0 * n, words, mask);
  } else {
    s.job_a = nttb200_wire_post_copy((int32_t *)s.h_a, a + s.row0 * n, words, 0);
    s.job_b = nttb200_wire_post_copy((int32_t *)s.h_b, b + s.row0 * n, words, 0);
  }
  s.kind = kind;
  s.state = WS_STAGING;
}

/* The pipeline runs over a STREAM of jobs: `feed(peek)` hands out the next job (nullptr: none right
 * now; peek = true only asks whether one is waiting), `done(job)` is told when the last row of a job has
 * reached its c.  While the last chunks of one job drain, the first chunks of the next one are already
 * being narrowed and sent -- what consecutive synchronous calls cannot do (their fill and drain, about 9 %
 * of a 2^16-row call, lie bare).  The synchronous call is the stream of one job. */
template <typename Feed, typename Done>
static int wire_stream(nttb200_plan *P, Feed &&feed, Done &&done) {
  int rc = ensure_wire_slots(P);
  if (rc) return rc;
  const size_t n = P->n;
  const int mode = wire_mode();
  const bool narrow_ok = P->plant && mode != 32;
  /* How much the pool can take.  A dozen threads or more keep up with the link (1 GPU, 16 host
   * cores: 29.5 M polymul/s all-narrowed against 27.4 M mixed and 22.3 M all-32-bit).  Eight ranks
   * sharing the 32 cores of one box get 4 threads each, and the box's memory system is what binds
   * all of them together: 30.5 M all-narrowed, 47.5 M mixed (a chunk goes out as 32-bit words when
   * the pool is already two chunks behind), 50.8 M all-32-bit.  Hence: >= 12 threads narrow
   * everything, 6..11 mix, fewer leave pinned buffers to the DMA engines. */
  const int pool = nttb200_wire_threads() / std::max(1, g_wire_share.load());
  const int ahead = env_int("NTTB200_WIRE_AHEAD", pool >= 12 ? 99 : (pool >= 6 ? 2 : 0), 0, 99);
  P->wire16_chunks = P->wire32_chunks = P->wire_c32_rows = 0;
  uint32_t mask[16] = {0};                             /* per slot: OR of the words with high bits */
  const size_t nsl = P->wslots.size();
  size_t busy = 0, chunk_no = 0;
  WireJob *cur = nullptr;
  const bool ramp = env_int("NTTB200_WIRE_RAMP", 1, 0, 1) != 0;
  auto retire = [&](WireSlot &s) {                     /* the slot's rows are in c */
    WireJob *jb = s.job;
    s.state = WS_FREE;
    s.job = nullptr;
    busy--;
    if (--jb->busy == 0 && jb->next >= jb->batch) { jb->finished = true; done(jb); }
  };
  rc = 0;
  nttb200_wire_begin();
  for (;;) {
    bool progress = false;
    int staging = 0;
    for (size_t i = 0; i < nsl && !rc; i++) {
      WireSlot &s = P->wslots[i];
      if (s.state == WS_STAGING) {
        if (nttb200_wire_done(s.job_a) && nttb200_wire_done(s.job_b)) {
          if (s.kind == WK_32_STAGED) rc = wire_send32(P, s, true);
          else if (!(mask[i] & 0xffff0000u)) rc = wire_send16(P, s);
          else if (s.job->pinned) rc = wire_send32(P, s, false);         /* a word does not fit 16 bits */
          else wire_stage_in(P, s, WK_32_STAGED, &mask[i]);
          progress = true;
        } else {
          staging++;
        }
      } else if (s.state == WS_INFLIGHT) {
        const cudaError_t e = cudaEventQuery(s.done);
        if (e == cudaSuccess) {
          if (!s.host_out) {
            retire(s);
          } else {
            int32_t *c = s.job->c;
            s.job_c = (s.kind == WK_16)
                          ? nttb200_wire_post_widen(c + s.row0 * n, (const uint16_t *)s.h_c, s.rows * n)
                          : nttb200_wire_post_copy(c + s.row0 * n, (const int32_t *)s.h_c, s.rows * n, 1);
            s.state = WS_WIDENING;
          }
          progress = true;
        } else if (e != cudaErrorNotReady) {
          rc = nttb200_fail(NTTB200_ECUDA, "wire pipeline: %s", cudaGetErrorString(e));
        }
      } else if (s.state == WS_WIDENING) {
        if (nttb200_wire_done(s.job_c)) {
          retire(s);
          progress = true;
        }
      }
    }
    if (rc) break;
    if (!cur) {
      if (busy == 0) chunk_no = 0;                     /* the pipeline ran dry: ramp up again */
      cur = feed(false);
      if (cur && cur->batch == 0) { cur->finished = true; done(cur); cur = nullptr; progress = true; }
    }
    if (cur) {
      for (size_t i = 0; i < nsl; i++) {
        WireSlot &s = P->wslots[i];
        if (s.state != WS_FREE) continue;
        /* taper the last chunks so that the drain (one kernel + D2H + widen) is short -- unless another
         * job is waiting, whose first chunks will cover it */
        const size_t left = cur->batch - cur->next;
        size_t nb = std::min(P->wire_polys, left);
        /* ... and ramp the first ones up (1/8, 1/4, 1/2 of a slot) so that the link starts early */
        if (ramp && chunk_no < 3) nb = std::min(nb, std::max<size_t>(P->wire_polys >> (3 - chunk_no), 64));
        if (left <= P->wire_polys && left > 64 && !feed(true)) nb = std::min(nb, std::max<size_t>(left / 2, 64));
        chunk_no++;
        s.job = cur;
        s.row0 = cur->next;
        s.rows = nb;
        cur->next += nb;
        cur->busy++;
        busy++;
        if (cur->pinned && (!narrow_ok || (mode == 0 && staging >= ahead))) rc = wire_send32(P, s, false);
        else wire_stage_in(P, s, narrow_ok ? WK_16 : WK_32_STAGED, &mask[i]);
        if (cur->next >= cur->batch) cur = nullptr;    /* all its rows are on their way */
        progress = true;
        break;
      }
      if (rc) break;
    }
    if (!cur && busy == 0 && !feed(true)) break;       /* nothing in flight, nothing waiting */
    if (!progress) nttb200_wire_help();
  }
  if (rc) {                                            /* leave no job or copy behind */
    for (auto &s : P->wslots) {
      if (s.state == WS_STAGING) { nttb200_wire_wait(s.job_a); nttb200_wire_wait(s.job_b); }
      if (s.state == WS_WIDENING) nttb200_wire_wait(s.job_c);
      cudaStreamSynchronize(s.stream);
      if (s.state != WS_FREE && s.job) {
        s.job->rc = rc;
        if (--s.job->busy == 0 && !s.job->finished && s.job != cur) { s.job->finished = true; done(s.job); }
      }
      s.state = WS_FREE;
      s.job = nullptr;
    }
    if (cur) { cur->rc = rc; cur->finished = true; done(cur); }
  }
  nttb200_wire_end();
  return rc;
}

static void wire_job_init(WireJob &jb, int32_t *c, const int32_t *a, const int32_t *b, size_t batch) {
  jb.c = c; jb.a = a; jb.b = b; jb.batch = batch;
  const bool c_pinned = is_pinned(c);
  jb.pinned = is_pinned(a) && is_pinned(b) && c_pinned;
  jb.c_direct = c_pinned && env_int("NTTB200_WIRE_C32", 0, 0, 1) == 1;
}

static int polymul_batch_wire(nttb200_plan *P, int32_t *c, const int32_t *a, const int32_t *b, size_t batch) {
  WireJob jb;
  wire_job_init(jb, c, a, b, batch);
  bool handed = false;
  int rc = wire_stream(P,
                       [&](bool peek) -> WireJob * { if (handed) return nullptr; if (!peek) handed = true; return &jb; },
                       [](WireJob *) {});
  return rc ? rc : jb.rc;
}

extern "C" int nttb200_plan_wire_stats(const nttb200_plan *P, unsigned long long *rows16,
                                       unsigned long long *rows32, unsigned long long *rows_c16,
                                       int *host_threads) {
  if (!P) return nttb200_fail(NTTB200_EPARAM, "NULL argument");
  if (rows16) *rows16 = P->wire16_chunks;
  if (rows32) *rows32 = P->wire32_chunks;
  if (rows_c16) *rows_c16 = P->wire16_chunks - P->wire_c32_rows;
  if (host_threads) *host_threads = P->wslots.empty() ? 0 : nttb200_wire_threads();
  return 0;
}

static bool wire_eligible(const nttb200_plan *P, const int32_t *c, const int32_t *a, const int32_t *b, size_t batch) {
  /* large batches: half-word moduli cross the link as 16-bit words; pageable buffers of any small-n
   * plan are staged by the host pool instead of the driver (wire_stream) */
  return P->kernel == PK_SMALL && batch * P->n >= WIRE_MIN_WORDS &&
         ((P->plant && wire_mode() != 32) ||
          (env_int("NTTB200_STAGE_PAGEABLE", 1, 0, 1) && !(is_pinned(a) && is_pinned(b) && is_pinned(c))));
}

/* the host-buffer product with the plan's mutex held and its device current */
static int polymul_batch_locked(nttb200_plan *P, int32_t *c, const int32_t *a, const int32_t *b, size_t batch) {
  int rc = ensure_slots(P, true);
  if (rc) return rc;
  const size_t n = P->n;
  P->wire16_chunks = P->wire32_chunks = P->wire_c32_rows = 0;
  /* Small calls (the reference's one-polynomial-per-call convention, nttb200_legacy.h): no DMA
   * at all -- operands are copied into a mapped pinned buffer that the kernel reads over PCIe
   * directly, and it writes c straight back into host memory: one launch + one sync. */
  if (batch * n * sizeof(uint32_t) <= ZC_BYTES && P->zc_host) {
    HostSlot &s = P->slots[0];
    const size_t bytes = batch * n * sizeof(uint32_t);
    uint32_t *ha = P->zc_host, *hb = ha + ZC_BYTES / 4, *hc = hb + ZC_BYTES / 4;
    uint32_t *da = P->zc_dev, *db = da + ZC_BYTES / 4, *dc = db + ZC_BYTES / 4;
    memcpy(ha, a, bytes);
    memcpy(hb, b, bytes);
    if ((rc = polymul_dev(P, dc, da, db, batch, s.stream))) return rc;
    NTT_CUDA(cudaStreamSynchronize(s.stream));
    memcpy(c, hc, bytes);
    return 0;
  }
  if (wire_eligible(P, c, a, b, batch)) return polymul_batch_wire(P, c, a, b, batch);
  size_t k = 0;
  for (size_t done = 0, nb = 0; done < batch; done += nb, k++) {
    HostSlot &s = P->slots[k % NSLOT];
    /* The call is H2D-bound and ends with one kernel + one D2H that nothing overlaps: taper the
     * last chunks (1/2, 1/4, 1/8 ... of a slot) so that this drain is short. */
    const size_t left = batch - done;
    nb = std::min(P->slot_polys, left);
    if (left <= P->slot_polys && left > 64) nb = std::max<size_t>(left / 2, 64);
    const size_t bytes = nb * n * sizeof(uint32_t);
    /* the slot's previous D2H is ordered before these copies by stream order */
    NTT_CUDA(cudaMemcpyAsync(s.d_a, a + done * n, bytes, cudaMemcpyHostToDevice, s.stream));
    NTT_CUDA(cudaMemcpyAsync(s.d_b, b + done * n, bytes, cudaMemcpyHostToDevice, s.stream));
    if ((rc = polymul_dev(P, s.d_c, s.d_a, s.d_b, nb, s.stream))) return rc;
    NTT_CUDA(cudaMemcpyAsync(c + done * n, s.d_c, bytes, cudaMemcpyDeviceToHost, s.stream));
  }
  for (auto &s : P->slots) NTT_CUDA(cudaStreamSynchronize(s.stream));
  return 0;
}

extern "C" int nttb200_polymul_batch(nttb200_plan *P, int32_t *c, const int32_t *a, const int32_t *b,
                                     size_t batch) {
  if (!P || !c || !a || !b) return nttb200_fail(NTTB200_EPARAM, "NULL argument");
  g_launches = 0;
  if (P->flags & NTTB200_PLAN_CHECK_RANGE) {
    int rcr = check_range_host(P, a, batch * P->n, "a");
    if (!rcr) rcr = check_range_host(P, b, batch * P->n, "b");
    if (rcr) return rcr;
  }
  std::lock_guard<std::mutex> lock(P->mu);
  DeviceGuard guard(P->device);
  return polymul_batch_locked(P, c, a, b, batch);
}

/* ---- asynchronous host-buffer products ------------------------------------------------------
 * nttb200_polymul_batch_async queues the product and returns a ticket; a worker thread of the plan
 * runs the queue through the SAME pipeline as the synchronous call -- as one stream of jobs, so the
 * first chunks of a product are narrowed and sent while the last chunks of the one before it drain.
 * nttb200_polymul_wait blocks until the ticket's c is complete (ticket 0: everything queued so far). */
struct AsyncJob {
  WireJob w;                  /* first member: the pipeline hands WireJob pointers back */
  bool done = false;
};
struct AsyncCtx {
  std::thread th;
  std::mutex mu;
  std::condition_variable cv_work, cv_done;
  std::deque<AsyncJob *> queue;                                  /* submitted, not yet taken */
  std::unordered_map<unsigned long long, AsyncJob *> jobs;       /* until waited for         */
  unsigned long long next_ticket = 1;
  bool stop = false;
};

static void async_worker(nttb200_plan *P) {
  AsyncCtx *A = P->async.load();
  cudaSetDevice(P->device);
  auto finish = [A](WireJob *w) {
    std::lock_guard<std::mutex> lk(A->mu);
    reinterpret_cast<AsyncJob *>(w)->done = true;
    A->cv_done.notify_all();
  };
  for (;;) {
    AsyncJob *head = nullptr;
    {
      std::unique_lock<std::mutex> lk(A->mu);
      A->cv_work.wait(lk, [&] { return A->stop || !A->queue.empty(); });
      if (A->queue.empty()) return;                              /* stop, and nothing left to do */
      head = A->queue.front();
    }
    std::lock_guard<std::mutex> plan_lock(P->mu);
    if (head->w.wire) {
      auto feed = [&](bool peek) -> WireJob * {
        std::lock_guard<std::mutex> lk(A->mu);
        if (A->queue.empty()) return nullptr;
        AsyncJob *j = A->queue.front();
        if (!j->w.wire) return nullptr;                          /* a small product: after the stream has drained */
        if (!peek) A->queue.pop_front();
        return &j->w;
      };
      int rc = wire_stream(P, feed, finish);
      if (rc) {                                                  /* could not even start: fail the head */
        std::unique_lock<std::mutex> lk(A->mu);
        if (!A->queue.empty() && A->queue.front() == head && !head->done) {
          A->queue.pop_front();
          head->w.rc = rc;
          head->done = true;
          A->cv_done.notify_all();
        }
      }
    } else {
      {
        std::lock_guard<std::mutex> lk(A->mu);
        A->queue.pop_front();
      }
      head->w.rc = polymul_batch_locked(P, head->w.c, head->w.a, head->w.b, head->w.batch);
      finish(&head->w);
    }
  }
}

extern "C" int nttb200_polymul_batch_async(nttb200_plan *P, int32_t *c, const int32_t *a, const int32_t *b,
                                           size_t batch, unsigned long long *ticket) {
  if (!P || !c || !a || !b || !ticket) return nttb200_fail(NTTB200_EPARAM, "NULL argument");
  if (P->flags & NTTB200_PLAN_CHECK_RANGE) {
    int rcr = check_range_host(P, a, batch * P->n, "a");
    if (!rcr) rcr = check_range_host(P, b, batch * P->n, "b");
    if (rcr) return rcr;
  }
  /* (the plan's own mutex is the worker's while it streams: submitting must not wait for it) */
  AsyncCtx *A = P->async.load();
  if (!A) {
    std::lock_guard<std::mutex> init(P->async_init_mu);
    if (!(A = P->async.load())) {
      {
        std::lock_guard<std::mutex> lock(P->mu);
        DeviceGuard guard(P->device);
        int rc = ensure_slots(P, true);
        if (rc) return rc;
      }
      A = new AsyncCtx();
      P->async.store(A);
      A->th = std::thread(async_worker, P);
    }
  }
  AsyncJob *j = new AsyncJob();
  {
    DeviceGuard guard(P->device);
    wire_job_init(j->w, c, a, b, batch);
    j->w.wire = wire_eligible(P, c, a, b, batch);
  }
  std::lock_guard<std::mutex> lk(A->mu);
  j->w.ticket = *ticket = A->next_ticket++;
  A->jobs[j->w.ticket] = j;
  A->queue.push_back(j);
  A->cv_work.notify_one();
  return 0;
}

extern "C" int nttb200_polymul_wait(nttb200_plan *P, unsigned long long ticket) {
  if (!P) return nttb200_fail(NTTB200_EPARAM, "NULL argument");
  AsyncCtx *A = P->async.load();
  if (!A) return ticket ? nttb200_fail(NTTB200_EPARAM, "unknown ticket %llu", ticket) : 0;
  std::unique_lock<std::mutex> lk(A->mu);
  int rc = 0;
  if (ticket == 0) {                                             /* everything submitted so far */
    std::vector<unsigned long long> all;
    for (auto &kv : A->jobs) all.push_back(kv.first);
    for (unsigned long long t : all) {
      AsyncJob *j = A->jobs[t];
      A->cv_done.wait(lk, [&] { return j->done; });
      if (j->w.rc && !rc) rc = j->w.rc;
      A->jobs.erase(t);
      delete j;
    }
    return rc ? nttb200_fail(rc, "an asynchronous product failed (%d)", rc) : 0;
  }
  auto it = A->jobs.find(ticket);
  if (it == A->jobs.end()) return nttb200_fail(NTTB200_EPARAM, "unknown ticket %llu (already waited for?)", ticket);
  AsyncJob *j = it->second;
  A->cv_done.wait(lk, [&] { return j->done; });
  rc = j->w.rc;
  A->jobs.erase(ticket);
  delete j;
  return rc ? nttb200_fail(rc, "asynchronous product %llu failed (%d)", ticket, rc) : 0;
}

static void async_shutdown(nttb200_plan *P) {
  AsyncCtx *A = P->async.load();
  if (!A) return;
  {
    std::lock_guard<std::mutex> lk(A->mu);
    A->stop = true;
    A->cv_work.notify_all();
  }
  if (A->th.joinable()) A->th.join();                            /* runs the queue dry first */
  for (auto &kv : A->jobs) delete kv.second;
  delete A;
  P->async.store(nullptr);
}

/* ---- packed 16-bit extension (outside the reference API; half-word moduli only) ---------- */
extern "C" int nttb200_polymul_batch_u16_dev(nttb200_plan *P, uint16_t *c, const uint16_t *a, const uint16_t *b,
                                             size_t batch, void *stream) {
  if (!P || !c || !a || !b) return nttb200_fail(NTTB200_EPARAM, "NULL argument");
  if (!P->plant)
    return nttb200_fail(NTTB200_EPARAM, "16-bit I/O needs a half-word modulus plan (q <= 12385, n <= 1024)");
  if ((((uintptr_t)a | (uintptr_t)b | (uintptr_t)c) & 15u) != 0)
    return nttb200_fail(NTTB200_EPARAM, "16-bit operands and result must be 16-byte aligned");
  g_launches = 0;
  if (batch == 0) return 0;
  DeviceGuard guard(P->device);
  return launch_polymul_small_plant_u16(P, c, a, b, batch, (cudaStream_t)stream);
}

extern "C" int nttb200_polymul_batch_u16(nttb200_plan *P, uint16_t *c, const uint16_t *a, const uint16_t *b,
                                         size_t batch) {
  if (!P || !c || !a || !b) return nttb200_fail(NTTB200_EPARAM, "NULL argument");
  if (!P->plant)
    return nttb200_fail(NTTB200_EPARAM, "16-bit I/O needs a half-word modulus plan (q <= 12385, n <= 1024)");
  g_launches = 0;
  std::lock_guard<std::mutex> lock(P->mu);
  DeviceGuard guard(P->device);
  int rc = ensure_slots(P, true);
  if (rc) return rc;
  const size_t n = P->n;
  const size_t slot = P->slot_polys * 2;             /* the slots are sized for 32-bit rows */
  size_t k = 0;
  for (size_t done = 0, nb = 0; done < batch; done += nb, k++) {
    HostSlot &s = P->slots[k % NSLOT];
    const size_t left = batch - done;
    nb = std::min(slot, left);
    if (left <= slot && left > 64) nb = std::max<size_t>(left / 2, 64);
    const size_t bytes = nb * n * sizeof(uint16_t);
    NTT_CUDA(cudaMemcpyAsync(s.d_a, a + done * n, bytes, cudaMemcpyHostToDevice, s.stream));
    NTT_CUDA(cudaMemcpyAsync(s.d_b, b + done * n, bytes, cudaMemcpyHostToDevice, s.stream));
    if ((rc = launch_polymul_small_plant_u16(P, (uint16_t *)s.d_c, (const uint16_t *)s.d_a,
                                             (const uint16_t *)s.d_b, nb, s.stream))) return rc;
    NTT_CUDA(cudaMemcpyAsync(c + done * n, s.d_c, bytes, cudaMemcpyDeviceToHost, s.stream));
  }
  for (auto &s : P->slots) NTT_CUDA(cudaStreamSynchronize(s.stream));
  return 0;
}

/* Standalone transforms of a large batch held in pageable memory: the wire slots and the host
 * pool as a parallel stager (copy in -> H2D -> transform in place -> D2H -> copy out), the same
 * state machine as polymul_batch_wire with one array. */
static int ntt_batch_staged(nttb200_plan *P, int transform, int32_t *a, size_t batch) {
  int rc = ensure_wire_slots(P);
  if (rc) return rc;
  const size_t n = P->n, nsl = P->wslots.size();
  size_t next = 0, busy = 0;
  nttb200_wire_begin();
  while (next < batch || busy > 0) {
    bool progress = false;
    for (size_t i = 0; i < nsl && !rc; i++) {
      WireSlot &s = P->wslots[i];
      const size_t bytes = s.rows * n * sizeof(uint32_t);
      if (s.state == WS_STAGING && nttb200_wire_done(s.job_a)) {
        cudaError_t e = cudaMemcpyAsync(s.d_a, s.h_a, bytes, cudaMemcpyHostToDevice, s.stream);
        if (e == cudaSuccess) {
          rc = transform_dev(P, transform, s.d_a, s.rows, s.stream);
          if (!rc) e = cudaMemcpyAsync(s.h_c, s.d_a, bytes, cudaMemcpyDeviceToHost, s.stream);
          if (!rc && e == cudaSuccess) e = cudaEventRecord(s.done, s.stream);
        }
        if (!rc && e != cudaSuccess) rc = nttb200_fail(NTTB200_ECUDA, "staged transform: %s", cudaGetErrorString(e));
        s.state = WS_INFLIGHT;
        progress = true;
      } else if (s.state == WS_INFLIGHT) {
        const cudaError_t e = cudaEventQuery(s.done);
        if (e == cudaSuccess) {
          s.job_c = nttb200_wire_post_copy(a + s.row0 * n, (const int32_t *)s.h_c, s.rows * n, 1);
          s.state = WS_WIDENING;
          progress = true;
        } else if (e != cudaErrorNotReady) {
          rc = nttb200_fail(NTTB200_ECUDA, "staged transform: %s", cudaGetErrorString(e));
        }
      } else if (s.state == WS_WIDENING && nttb200_wire_done(s.job_c)) {
        s.state = WS_FREE;
        busy--;
        progress = true;
      }
    }
    if (rc) break;
    if (next < batch) {
      for (size_t i = 0; i < nsl; i++) {
        WireSlot &s = P->wslots[i];
        if (s.state != WS_FREE) continue;
        s.row0 = next;
        s.rows = std::min(P->wire_polys, batch - next);
        next += s.rows;
        busy++;
        s.job_a = nttb200_wire_post_copy((int32_t *)s.h_a, a + s.row0 * n, s.rows * n, 0);
        s.state = WS_STAGING;
        progress = true;
        break;
      }
    }
    if (!progress) nttb200_wire_help();
  }
  if (rc) {
    for (auto &s : P->wslots) {
      if (s.state == WS_STAGING) nttb200_wire_wait(s.job_a);
      if (s.state == WS_WIDENING) nttb200_wire_wait(s.job_c);
      cudaStreamSynchronize(s.stream);
      s.state = WS_FREE;
    }
  }
  nttb200_wire_end();
  return rc;
}

extern "C" int nttb200_ntt_batch(nttb200_plan *P, int transform, int32_t *a, size_t batch) {
  if (!P || !a) return nttb200_fail(NTTB200_EPARAM, "NULL argument");
  g_launches = 0;
  if (P->flags & NTTB200_PLAN_CHECK_RANGE) {
    const int rcr = check_range_host(P, a, batch * P->n, "a");
    if (rcr) return rcr;
  }
  std::lock_guard<std::mutex> lock(P->mu);
  DeviceGuard guard(P->device);
  int rc = ensure_slots(P, false);
  if (rc) return rc;
  const size_t n = P->n;
  if (batch * n >= WIRE_MIN_WORDS && env_int("NTTB200_STAGE_PAGEABLE", 1, 0, 1) && !is_pinned(a))
    return ntt_batch_staged(P, transform, a, batch);
  size_t k = 0;
  for (size_t done = 0; done < batch; done += P->slot_polys, k++) {
    HostSlot &s = P->slots[k % NSLOT];
    const size_t nb = std::min(P->slot_polys, batch - done);
    const size_t bytes = nb * n * sizeof(uint32_t);
    NTT_CUDA(cudaMemcpyAsync(s.d_a, a + done * n, bytes, cudaMemcpyHostToDevice, s.stream));
    if ((rc = transform_dev(P, transform, s.d_a, nb, s.stream))) return rc;
    NTT_CUDA(cudaMemcpyAsync(a + done * n, s.d_a, bytes, cudaMemcpyDeviceToHost, s.stream));
  }
  for (auto &s : P->slots) NTT_CUDA(cudaStreamSynchronize(s.stream));
  return 0;
}

extern "C" int nttb200_mul_array_batch(nttb200_plan *P, int32_t *c, const int32_t *a, const int32_t *b,
                                       size_t batch) {
  if (!P || !c || !a || !b) return nttb200_fail(NTTB200_EPARAM, "NULL argument");
  g_launches = 0;
  std::lock_guard<std::mutex> lock(P->mu);
  DeviceGuard guard(P->device);
  int rc = ensure_slots(P, true);
  if (rc) return rc;
  const uint64_t r1 = (1ull << 32) % P->q;
  const uint32_t r2 = (uint32_t)(r1 * r1 % P->q);
  const size_t n = P->n;
  size_t k = 0;
  for (size_t done = 0; done < batch; done += P->slot_polys, k++) {
    HostSlot &s = P->slots[k % NSLOT];
    const size_t nb = std::min(P->slot_polys, batch - done);
    const size_t bytes = nb * n * sizeof(uint32_t);
    NTT_CUDA(cudaMemcpyAsync(s.d_a, a + done * n, bytes, cudaMemcpyHostToDevice, s.stream));
    NTT_CUDA(cudaMemcpyAsync(s.d_b, b + done * n, bytes, cudaMemcpyHostToDevice, s.stream));
    if ((rc = launch_pointwise(s.d_c, s.d_a, s.d_b, nb * n, P->m, r2, s.stream))) return rc;
    NTT_CUDA(cudaMemcpyAsync(c + done * n, s.d_c, bytes, cudaMemcpyDeviceToHost, s.stream));
  }
  for (auto &s : P->slots) NTT_CUDA(cudaStreamSynchronize(s.stream));
  return 0;
}

extern "C" int nttb200_scalar_mul_array_batch(nttb200_plan *P, int32_t *a, int32_t sc, size_t batch) {
  if (!P || !a) return nttb200_fail(NTTB200_EPARAM, "NULL argument");
  if (sc < 0 || (uint32_t)sc >= P->q) return nttb200_fail(NTTB200_ERANGE, "scalar outside [0, q)");
  g_launches = 0;
  std::lock_guard<std::mutex> lock(P->mu);
  DeviceGuard guard(P->device);
  int rc = ensure_slots(P, false);
  if (rc) return rc;
  const size_t n = P->n;
  size_t k = 0;
  for (size_t done = 0; done < batch; done += P->slot_polys, k++) {
    HostSlot &s = P->slots[k % NSLOT];
    const size_t nb = std::min(P->slot_polys, batch - done);
    const size_t bytes = nb * n * sizeof(uint32_t);
    NTT_CUDA(cudaMemcpyAsync(s.d_a, a + done * n, bytes, cudaMemcpyHostToDevice, s.stream));
    if ((rc = launch_scale(s.d_a, nullptr, pair_of((uint32_t)sc, P->q), P->n, nb * n, P->m, s.stream)))
      return rc;
    NTT_CUDA(cudaMemcpyAsync(a + done * n, s.d_a, bytes, cudaMemcpyDeviceToHost, s.stream));
  }
  for (auto &s : P->slots) NTT_CUDA(cudaStreamSynchronize(s.stream));
  return 0;
}

/* Process-wide mapped pinned scratch for the one-polynomial-per-call compatibility functions:
 * [ table (n x 8 B) | data ], read and written by the kernel over PCIe; one launch + one sync. */
struct ZcCtx {
  std::mutex mu;
  unsigned char *h = nullptr, *d = nullptr;
  size_t bytes = 0;
  cudaStream_t st = nullptr;
  int dev = -1;
};
static ZcCtx g_zc;
static const size_t ZC_CTX_BYTES = 1u << 20;
static const uint32_t LITERAL_CTA_MAX_N = 4096;

/* returns 0 and the locked context, or non-zero when the fast path is not available */
static int zc_acquire(std::unique_lock<std::mutex> &lk) {
  lk = std::unique_lock<std::mutex>(g_zc.mu);
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 1; }
  if (g_zc.h && g_zc.dev == dev) return 0;
  if (g_zc.h) return 1;                              /* created on another device: use the slow path */
  void *h = nullptr, *d = nullptr;
  if (cudaHostAlloc(&h, ZC_CTX_BYTES, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess ||
      cudaHostGetDevicePointer(&d, h, 0) != cudaSuccess ||
      cudaStreamCreateWithFlags(&g_zc.st, cudaStreamNonBlocking) != cudaSuccess) {
    if (h) cudaFreeHost(h);
    cudaGetLastError();
    return 1;
  }
  g_zc.h = (unsigned char *)h; g_zc.d = (unsigned char *)d; g_zc.bytes = ZC_CTX_BYTES; g_zc.dev = dev;
  return 0;
}

template <bool RED>
static int literal_cta_launch(int dataflow, uint32_t *d_a, const uint2 *d_tab, const int32_t *d_rtab, uint32_t n,
                              size_t batch, const ModQ &m, int skip0, cudaStream_t st) {
  using namespace nttb200;
  const uint32_t logn = ht_log2(n);
  const int threads = (int)std::min<uint32_t>(512, std::max<uint32_t>(32, n / 2));
  const int grid = (int)std::min<size_t>(batch, (size_t)current_sms() * 4);
  const size_t smem = n * sizeof(uint32_t) + n * (RED ? sizeof(int32_t) : sizeof(uint2));   /* <= 48 KiB */
  switch (dataflow) {
    case DF_CT_STD2REV: literal_cta_kernel<DF_CT_STD2REV, RED><<<grid, threads, smem, st>>>(d_a, d_tab, d_rtab, n, logn, batch, m, skip0); break;
    case DF_GS_REV2STD: literal_cta_kernel<DF_GS_REV2STD, RED><<<grid, threads, smem, st>>>(d_a, d_tab, d_rtab, n, logn, batch, m, skip0); break;
    case DF_CT_REV2STD: literal_cta_kernel<DF_CT_REV2STD, RED><<<grid, threads, smem, st>>>(d_a, d_tab, d_rtab, n, logn, batch, m, skip0); break;
    default: literal_cta_kernel<DF_GS_STD2REV, RED><<<grid, threads, smem, st>>>(d_a, d_tab, d_rtab, n, logn, batch, m, skip0); break;
  }
  nttb200_count_launch(1);
  NTT_CUDA(cudaGetLastError());
  return 0;
}

/* Table-driven transform: the literal reference dataflow with the caller's table. */
extern "C" int nttb200_ntt_table_batch(uint32_t n, uint32_t q, int dataflow, int skip_j0, const uint32_t *p,
                                       int32_t *a, size_t batch) {
  if (!p || !a) return nttb200_fail(NTTB200_EPARAM, "NULL argument");
  if (n < 2 || (n & (n - 1)) || q < 3 || q >= (1u << 31) || !(q & 1))
    return nttb200_fail(NTTB200_EPARAM, "bad (n=%u, q=%u)", n, q);
  if (dataflow < 0 || dataflow > 3) return nttb200_fail(NTTB200_EPARAM, "unknown dataflow %d", dataflow);
  if (nttb200_device_count() <= 0) return nttb200_fail(NTTB200_ECUDA, "no CUDA device (there is no CPU fallback)");
  g_launches = 0;
  if (batch == 0) return 0;
  std::vector<uint2> h(n);
  for (uint32_t i = 0; i < n; i++) {
    const uint32_t w = p[i] % q;
    h[i] = make_uint2(w, ht_shoup(w, q));
  }
  ModQ m;
  m.q = q; m.nq = 0u - q; m.q2 = 2u * q;
  uint32_t inv = 1;
  for (int i = 0; i < 5; i++) inv *= 2u - q * inv;
  m.qinv = inv;
  const size_t bytes = batch * n * sizeof(uint32_t);
  if (n <= LITERAL_CTA_MAX_N && n * sizeof(uint2) + bytes <= ZC_CTX_BYTES) {   /* one launch, no DMA */
    std::unique_lock<std::mutex> lk;
    if (zc_acquire(lk) == 0) {
      uint2 *h_tab = (uint2 *)g_zc.h;
      uint32_t *h_a = (uint32_t *)(g_zc.h + n * sizeof(uint2));
      memcpy(h_tab, h.data(), n * sizeof(uint2));
      memcpy(h_a, a, bytes);
      int rc0 = literal_cta_launch<false>(dataflow, (uint32_t *)(g_zc.d + n * sizeof(uint2)), (const uint2 *)g_zc.d,
                                          nullptr, n, batch, m, skip_j0 ? 1 : 0, g_zc.st);
      if (rc0) return rc0;
      NTT_CUDA(cudaStreamSynchronize(g_zc.st));
      memcpy(a, h_a, bytes);
      return 0;
    }
  }
  uint2 *d_tab = nullptr;
  uint32_t *d_a = nullptr;
  int rc = 0;
  cudaError_t e = cudaMalloc(&d_tab, n * sizeof(uint2));
  if (e == cudaSuccess) e = cudaMalloc(&d_a, bytes);
  if (e == cudaSuccess) e = cudaMemcpy(d_tab, h.data(), n * sizeof(uint2), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d_a, a, bytes, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) rc = launch_generic_transform(n, ht_log2(n), m, dataflow, d_tab, d_a, batch, 0, skip_j0 ? 1 : 0);
  if (e == cudaSuccess && rc == 0) e = cudaMemcpy(a, d_a, bytes, cudaMemcpyDeviceToHost);
  cudaFree(d_tab);
  cudaFree(d_a);
  if (e != cudaSuccess) return nttb200_fail(NTTB200_ECUDA, "table transform: %s", cudaGetErrorString(e));
  return rc;
}

/* ------------------------------------------------------------------------------------ */
/* multi-GPU: contiguous batch slices, one plan + one host thread per GPU, no collective   */
/* (independent products share nothing but the read-only tables -- SURVEY 8e)              */
/* ------------------------------------------------------------------------------------ */
struct nttb200_multi {
  std::vector<nttb200_plan *> plans;
  uint32_t n = 0;
};

extern "C" int nttb200_multi_create(nttb200_multi **out, uint32_t n, uint32_t q, uint32_t psi, uint32_t flags,
                                    int ngpus) {
  if (!out) return nttb200_fail(NTTB200_EPARAM, "multi pointer is NULL");
  *out = nullptr;
  const int have = nttb200_device_count();
  if (have <= 0) return nttb200_fail(NTTB200_ECUDA, "no CUDA device (there is no CPU fallback)");
  if (ngpus <= 0) ngpus = have;
  if (ngpus > have) return nttb200_fail(NTTB200_EPARAM, "%d GPUs requested, %d visible", ngpus, have);
  nttb200_multi *M = new (std::nothrow) nttb200_multi;
  if (!M) return nttb200_fail(NTTB200_ENOMEM, "out of host memory");
  M->n = n;
  const int saved = g_device;
  int rc = 0;
  for (int g = 0; g < ngpus && rc == 0; g++) {
    nttb200_plan *P = nullptr;
    rc = nttb200_set_device(g);
    if (rc == 0) rc = nttb200_plan_create(&P, n, q, psi, flags);
    if (rc == 0) M->plans.push_back(P);
  }
  g_device = saved;
  if (saved >= 0) cudaSetDevice(saved);
  if (rc) {
    for (auto *P : M->plans) nttb200_plan_destroy(P);
    delete M;
    return rc;
  }
  *out = M;
  return 0;
}
extern "C" void nttb200_multi_destroy(nttb200_multi *M) {
  if (!M) return;
  for (auto *P : M->plans) nttb200_plan_destroy(P);
  delete M;
}
extern "C" int nttb200_multi_gpus(const nttb200_multi *M) { return M ? (int)M->plans.size() : 0; }
extern "C" nttb200_plan *nttb200_multi_plan(nttb200_multi *M, int gpu) {
  return (M && gpu >= 0 && gpu < (int)M->plans.size()) ? M->plans[gpu] : nullptr;
}
/* rows [g B/G, (g+1) B/G) go to GPU g (the partition of SURVEY 8e) */
extern "C" void nttb200_shard_bounds(size_t batch, int world, int rank, size_t *lo, size_t *hi) {
  if (world < 1) world = 1;
  if (lo) *lo = (size_t)(((unsigned __int128)batch * (unsigned)rank) / (unsigned)world);
  if (hi) *hi = (size_t)(((unsigned __int128)batch * (unsigned)(rank + 1)) / (unsigned)world);
}
extern "C" int nttb200_multi_polymul_batch(nttb200_multi *M, int32_t *c, const int32_t *a, const int32_t *b,
                                           size_t batch) {
  if (!M || !c || !a || !b) return nttb200_fail(NTTB200_EPARAM, "NULL argument");
  const int G = (int)M->plans.size();
  std::vector<int> rcs(G, 0);
  std::vector<std::string> errs(G);
  std::vector<int> launches(G, 0);
  std::vector<std::thread> th;
  g_wire_share.store(G);                              /* the G calls below share one host pool */
  for (int g = 0; g < G; g++) {
    th.emplace_back([&, g]() {
      size_t lo, hi;
      nttb200_shard_bounds(batch, G, g, &lo, &hi);
      const size_t o = lo * M->n;
      rcs[g] = nttb200_polymul_batch(M->plans[g], c + o, a + o, b + o, hi - lo);
      if (rcs[g]) errs[g] = nttb200_last_error();
      launches[g] = nttb200_last_launch_count();
    });
  }
  for (auto &t : th) t.join();
  g_wire_share.store(1);
  g_launches = 0;
  for (int g = 0; g < G; g++) {
    g_launches += launches[g];
    if (rcs[g]) return nttb200_fail(rcs[g], "GPU %d: %s", g, errs[g].c_str());
  }
  return 0;
}

/* ------------------------------------------------------------------------------------ */
/* exact emulation of the reference's RED surface and the permutation helpers (q = 12289)   */
/* ------------------------------------------------------------------------------------ */
template <typename F>
static int with_device_copy(void *host, size_t bytes, F &&body) {
  if (nttb200_device_count() <= 0) return nttb200_fail(NTTB200_ECUDA, "no CUDA device (there is no CPU fallback)");
  void *d = nullptr;
  cudaError_t e = cudaMalloc(&d, bytes ? bytes : 1);
  if (e == cudaSuccess) e = cudaMemcpy(d, host, bytes, cudaMemcpyHostToDevice);
  int rc = 0;
  if (e == cudaSuccess) rc = body(d);
  if (e == cudaSuccess && rc == 0) e = cudaGetLastError();
  if (e == cudaSuccess && rc == 0) e = cudaMemcpy(host, d, bytes, cudaMemcpyDeviceToHost);
  cudaFree(d);
  if (e != cudaSuccess) return nttb200_fail(NTTB200_ECUDA, "%s", cudaGetErrorString(e));
  return rc;
}

extern "C" int nttb200_red_ntt_table_batch(uint32_t n, int dataflow, int skip_j0, const int32_t *p, int32_t *a,
                                           size_t batch) {
  using namespace nttb200;
  if (!p || !a) return nttb200_fail(NTTB200_EPARAM, "NULL argument");
  if (n < 2 || (n & (n - 1))) return nttb200_fail(NTTB200_EPARAM, "n=%u is not a power of two", n);
  if (dataflow < 0 || dataflow > 3) return nttb200_fail(NTTB200_EPARAM, "unknown dataflow %d", dataflow);
  g_launches = 0;
  if (batch == 0) return 0;
  const uint32_t logn = ht_log2(n);
  if (n <= LITERAL_CTA_MAX_N && n * sizeof(int32_t) + batch * n * sizeof(int32_t) <= ZC_CTX_BYTES &&
      nttb200_device_count() > 0) {                                               /* one launch, no DMA */
    std::unique_lock<std::mutex> lk;
    if (zc_acquire(lk) == 0) {
      const size_t bytes = batch * n * sizeof(int32_t);
      int32_t *h_tab = (int32_t *)g_zc.h;
      int32_t *h_a = (int32_t *)(g_zc.h + n * sizeof(int32_t));
      memcpy(h_tab, p, n * sizeof(int32_t));
      memcpy(h_a, a, bytes);
      ModQ m{};
      int rc0 = literal_cta_launch<true>(dataflow, (uint32_t *)(g_zc.d + n * sizeof(int32_t)), nullptr,
                                         (const int32_t *)g_zc.d, n, batch, m, skip_j0, g_zc.st);
      if (rc0) return rc0;
      NTT_CUDA(cudaStreamSynchronize(g_zc.st));
      memcpy(a, h_a, bytes);
      return 0;
    }
  }
  return with_device_copy(a, batch * n * sizeof(int32_t), [&](void *d_a) -> int {
    int32_t *d_tab = nullptr;
    NTT_CUDA(cudaMalloc(&d_tab, n * sizeof(int32_t)));
    cudaError_t e = cudaMemcpy(d_tab, p, n * sizeof(int32_t), cudaMemcpyHostToDevice);
    const unsigned long long pairs = (unsigned long long)batch * (n / 2);
    const int grid = grid_1d(pairs, 256, current_sms());
    const bool descending = (dataflow == DF_CT_STD2REV || dataflow == DF_GS_STD2REV);
    for (uint32_t s = 0; s < logn && e == cudaSuccess; s++) {
      const uint32_t lh = descending ? (logn - 1 - s) : s;
      const uint32_t half = 1u << lh;
      int32_t *x = (int32_t *)d_a;
      switch (dataflow) {
        case DF_CT_STD2REV: generic_red_stage_kernel<DF_CT_STD2REV><<<grid, 256>>>(x, d_tab, n, logn, half, lh, pairs, skip_j0); break;
        case DF_GS_REV2STD: generic_red_stage_kernel<DF_GS_REV2STD><<<grid, 256>>>(x, d_tab, n, logn, half, lh, pairs, skip_j0); break;
        case DF_CT_REV2STD: generic_red_stage_kernel<DF_CT_REV2STD><<<grid, 256>>>(x, d_tab, n, logn, half, lh, pairs, skip_j0); break;
        default: generic_red_stage_kernel<DF_GS_STD2REV><<<grid, 256>>>(x, d_tab, n, logn, half, lh, pairs, skip_j0); break;
      }
      nttb200_count_launch(1);
    }
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaFree(d_tab);
    if (e != cudaSuccess) return nttb200_fail(NTTB200_ECUDA, "RED transform: %s", cudaGetErrorString(e));
    return 0;
  });
}

extern "C" int nttb200_red_elementwise_batch(int op, int32_t *c, const int32_t *a, const int32_t *b, int32_t scalar,
                                             size_t count) {
  using namespace nttb200;
  if (!c || !a || (op == RED_OP_MUL_RED && !b)) return nttb200_fail(NTTB200_EPARAM, "NULL argument");
  if (op < 0 || op > RED_OP_SCALAR_MUL_RED) return nttb200_fail(NTTB200_EPARAM, "unknown RED op %d", op);
  if (nttb200_device_count() <= 0) return nttb200_fail(NTTB200_ECUDA, "no CUDA device (there is no CPU fallback)");
  g_launches = 0;
  if (count == 0) return 0;
  const size_t bytes = count * sizeof(int32_t);
  if (2 * bytes <= ZC_CTX_BYTES) {                                                /* one launch, no DMA */
    std::unique_lock<std::mutex> lk;
    if (zc_acquire(lk) == 0) {
      int32_t *h_a = (int32_t *)g_zc.h, *h_b = (int32_t *)(g_zc.h + bytes);
      memcpy(h_a, a, bytes);
      if (op == RED_OP_MUL_RED) memcpy(h_b, b, bytes);
      red_elementwise_kernel<<<grid_1d(count, 256, current_sms()), 256, 0, g_zc.st>>>(
          op, (int32_t *)g_zc.d, (const int32_t *)g_zc.d, (const int32_t *)(g_zc.d + bytes), scalar, count);
      nttb200_count_launch(1);
      NTT_CUDA(cudaGetLastError());
      NTT_CUDA(cudaStreamSynchronize(g_zc.st));
      memcpy(c, h_a, bytes);
      return 0;
    }
  }
  int32_t *d_a = nullptr, *d_b = nullptr;
  cudaError_t e = cudaMalloc(&d_a, bytes);
  if (e == cudaSuccess) e = cudaMemcpy(d_a, a, bytes, cudaMemcpyHostToDevice);
  if (e == cudaSuccess && op == RED_OP_MUL_RED) {
    e = cudaMalloc(&d_b, bytes);
    if (e == cudaSuccess) e = cudaMemcpy(d_b, b, bytes, cudaMemcpyHostToDevice);
  }
  if (e == cudaSuccess) {
    red_elementwise_kernel<<<grid_1d(count, 256, current_sms()), 256>>>(op, d_a, d_a, d_b, scalar, count);
    nttb200_count_launch(1);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy(c, d_a, bytes, cudaMemcpyDeviceToHost);
  cudaFree(d_a);
  cudaFree(d_b);
  if (e != cudaSuccess) return nttb200_fail(NTTB200_ECUDA, "RED elementwise: %s", cudaGetErrorString(e));
  return 0;
}

extern "C" int nttb200_bitrev_shuffle_batch(int32_t *a, uint32_t n, size_t batch) {
  if (!a) return nttb200_fail(NTTB200_EPARAM, "NULL argument");
  if (n == 0 || (n & (n - 1))) return nttb200_fail(NTTB200_EPARAM, "n=%u is not a power of two", n);
  g_launches = 0;
  if (batch == 0 || n < 4) return nttb200_device_count() > 0 ? 0 : nttb200_fail(NTTB200_ECUDA, "no CUDA device (there is no CPU fallback)");
  const unsigned long long total = (unsigned long long)batch * n;
  return with_device_copy(a, total * sizeof(int32_t), [&](void *d) -> int {
    nttb200::bitrev_shuffle_kernel<<<grid_1d(total, 256, current_sms()), 256>>>((uint32_t *)d, ht_log2(n), total);
    nttb200_count_launch(1);
    return 0;
  });
}

extern "C" int nttb200_shuffle_with_table(int32_t *a, size_t words, const uint16_t *pairs, uint32_t npairs) {
  if (!a || (!pairs && npairs)) return nttb200_fail(NTTB200_EPARAM, "NULL argument");
  for (uint32_t i = 0; i < 2 * npairs; i++)
    if (pairs[i] >= words) return nttb200_fail(NTTB200_EPARAM, "swap index %u outside the array", pairs[i]);
  g_launches = 0;
  return with_device_copy(a, words * sizeof(int32_t), [&](void *d) -> int {
    uint16_t *d_p = nullptr;
    NTT_CUDA(cudaMalloc(&d_p, (npairs ? npairs : 1) * 2 * sizeof(uint16_t)));
    cudaError_t e = cudaMemcpy(d_p, pairs, npairs * 2 * sizeof(uint16_t), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
      nttb200::shuffle_table_kernel<<<1, 32>>>((uint32_t *)d, d_p, npairs);
      nttb200_count_launch(1);
      e = cudaDeviceSynchronize();
    }
    cudaFree(d_p);
    if (e != cudaSuccess) return nttb200_fail(NTTB200_ECUDA, "shuffle: %s", cudaGetErrorString(e));
    return 0;
  });
}

/* ------------------------------------------------------------------------------------ */
/* memory helpers                                                                        */
/* ------------------------------------------------------------------------------------ */
extern "C" void *nttb200_host_alloc(size_t bytes) {
  void *p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) {
    nttb200_fail(NTTB200_ENOMEM, "cudaHostAlloc(%zu) failed: %s", bytes, cudaGetErrorString(cudaGetLastError()));
    return nullptr;
  }
  return p;
}
extern "C" void nttb200_host_free(void *p) { if (p) cudaFreeHost(p); }
extern "C" void *nttb200_dev_alloc(size_t bytes) {
  void *p = nullptr;
  if (cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) {
    nttb200_fail(NTTB200_ENOMEM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(cudaGetLastError()));
    return nullptr;
  }
  return p;
}
extern "C" void nttb200_dev_free(void *p) { if (p) cudaFree(p); }
extern "C" int nttb200_memcpy_h2d(void *dst, const void *src, size_t bytes, void *stream) {
  NTT_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
  return 0;
}
extern "C" int nttb200_memcpy_d2h(void *dst, const void *src, size_t bytes, void *stream) {
  NTT_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  return 0;
}
extern "C" void *nttb200_stream_create(void) {
  cudaStream_t s = nullptr;
  if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) {
    nttb200_fail(NTTB200_ECUDA, "cudaStreamCreate failed: %s", cudaGetErrorString(cudaGetLastError()));
    return nullptr;
  }
  return (void *)s;
}
extern "C" void nttb200_stream_destroy(void *stream) { if (stream) cudaStreamDestroy((cudaStream_t)stream); }
extern "C" int nttb200_stream_query(void *stream) {
  cudaError_t e = cudaStreamQuery((cudaStream_t)stream);
  if (e == cudaSuccess) return 1;
  if (e == cudaErrorNotReady) { return 0; }
  return nttb200_fail(NTTB200_ECUDA, "cudaStreamQuery: %s", cudaGetErrorString(e));
}
extern "C" int nttb200_stream_sync(void *stream) {
  NTT_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return 0;
}
