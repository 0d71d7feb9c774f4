/*
 * large_dispatch.inl -- instantiates the multi-pass kernels of ntt_large.cuh (n = 2^11 ..
 * 2^17, rows of 256 coefficients) for ONE arithmetic class (LARGE_ARITH) and defines its
 * dispatch functions.  Included by large_lazy.cu, large_harvey.cu and large_canon.cu.
 */
#include <cuda_runtime.h>

#include <algorithm>

#include "ntt_large.cuh"
#include "ntt_large_fused.cuh"
#include "plan.h"

namespace {
using namespace nttb200;

constexpr int LR = LARGE_LR;    /* row length 2^LR */
constexpr int ROW_WARPS = 8;

inline uint2 lshoup_pair(uint64_t w, uint32_t q) {
  w %= q;
  return make_uint2((uint32_t)w, (uint32_t)((w << 32) / q));
}

void fill_common(LargeParams &p, const nttb200_plan *P, const DevTable *fwd, const DevTable *inv,
                 size_t batch) {
  p.tab = fwd ? fwd->d : nullptr;
  p.tab_inv = inv ? inv->d : nullptr;
  p.batch = batch;
  p.logn = P->logn;
  p.nops = 1;
  p.m = P->m;
  p.one = lshoup_pair(1, P->q);
  p.last_x = p.one;
  p.last_y = inv ? inv->h[1] : p.one;
}
/* the uniform twiddles of a column launch: forward pass = forward table, inverse pass = inverse table */
void fill_utw(LargeParams &p, const DevTable *t) {
  const size_t lim = std::min<size_t>(32, t->h.size());
  for (size_t i = 1; i < lim; i++) p.utw[i] = t->h[i];
}

/* every kernel of the pipeline is launched with programmatic stream serialization: its CTAs are
 * scheduled while the previous kernel of the lane drains, and wait (griddepcontrol.wait) where they
 * first need its results */
template <typename K>
int launch_pdl_large(K kernel, unsigned grid, int threads, int smem, cudaStream_t st, const LargeParams &p) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3((unsigned)threads);
  cfg.dynamicSmemBytes = (size_t)smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  NTT_CUDA(cudaLaunchKernelEx(&cfg, kernel, p));
  nttb200_count_launch(1);
  NTT_CUDA(cudaGetLastError());
  return 0;
}

/* two columns per lane for the classes with cheap butterflies, when the rows are 8-byte aligned */
constexpr int LARGE_CPL = (LARGE_ARITH == ARITH_CANON) ? 1 : 2;
inline bool cols_aligned8(const LargeParams &p) {
  uintptr_t v = 0;
  for (int i = 0; i < 2; i++) v |= (uintptr_t)p.src[i] | (uintptr_t)p.dst[i];
  return (v & 7u) == 0;
}

template <int K1, int CPL>
int cols_fwd_cpl(const nttb200_plan *P, LargeParams &p, cudaStream_t st) {
  using G = ColGeom<K1, CPL>;
  auto kernel = large_cols_fwd_kernel<K1, LARGE_ARITH, CPL>;
  if (G::SMEM_BYTES > 48 * 1024)
    NTT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM_BYTES));
  const unsigned long long grid = p.batch * p.nops * (1ull << (P->logn - K1 - G::LOG_TILE));
  if (grid > 0x7fffffffull) return nttb200_fail(NTTB200_EPARAM, "batch too large for one launch");
  return launch_pdl_large(kernel, (unsigned)grid, G::WARPS * 32, G::SMEM_BYTES, st, p);
}
template <int K1>
int cols_fwd(const nttb200_plan *P, LargeParams &p, cudaStream_t st) {
  if (LARGE_CPL == 2 && cols_aligned8(p)) return cols_fwd_cpl<K1, LARGE_CPL>(P, p, st);
  return cols_fwd_cpl<K1, 1>(P, p, st);
}

template <int K1, int CPL>
int cols_inv_cpl(const nttb200_plan *P, LargeParams &p, cudaStream_t st) {
  using G = ColGeom<K1, CPL>;
  auto kernel = large_cols_inv_kernel<K1, LARGE_ARITH, CPL>;
  if (G::SMEM_BYTES > 48 * 1024)
    NTT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM_BYTES));
  const unsigned long long grid = p.batch * (1ull << (P->logn - K1 - G::LOG_TILE));
  if (grid > 0x7fffffffull) return nttb200_fail(NTTB200_EPARAM, "batch too large for one launch");
  return launch_pdl_large(kernel, (unsigned)grid, G::WARPS * 32, G::SMEM_BYTES, st, p);
}
template <int K1>
int cols_inv(const nttb200_plan *P, LargeParams &p, cudaStream_t st) {
  if (LARGE_CPL == 2 && cols_aligned8(p)) return cols_inv_cpl<K1, LARGE_CPL>(P, p, st);
  return cols_inv_cpl<K1, 1>(P, p, st);
}

#define LARGE_K1_SWITCH(fn, ...)                                                   \
  switch (P->logn - LR) {                                                          \
    case 3: return fn<3>(__VA_ARGS__);                                             \
    case 4: return fn<4>(__VA_ARGS__);                                             \
    case 5: return fn<5>(__VA_ARGS__);                                             \
    case 6: return fn<6>(__VA_ARGS__);                                             \
    case 7: return fn<7>(__VA_ARGS__);                                             \
    case 8: return fn<8>(__VA_ARGS__);                                             \
    case 9: return fn<9>(__VA_ARGS__);                                             \
    default: return nttb200_fail(NTTB200_EPARAM, "multi-pass kernels cover 2^11 <= n <= 2^17"); \
  }

int cols_fwd_any(const nttb200_plan *P, LargeParams &p, cudaStream_t st) { LARGE_K1_SWITCH(cols_fwd, P, p, st) }
int cols_inv_any(const nttb200_plan *P, LargeParams &p, cudaStream_t st) { LARGE_K1_SWITCH(cols_inv, P, p, st) }

unsigned long long row_grid(const nttb200_plan *P, size_t batch) {
  using Gm = SmallGeom<LR>;
  const unsigned long long groups = (batch + ROW_WARPS * Gm::PPW - 1) / (ROW_WARPS * Gm::PPW);
  return groups << (P->logn - LR);
}

int rows_polymul(const nttb200_plan *P, LargeParams &p, cudaStream_t st) {
  using Gm = SmallGeom<LR>;
  auto kernel = large_rows_polymul_kernel<LR, LARGE_ARITH, ROW_WARPS>;
  const int smem = ROW_WARPS * 2 * Gm::PPW * Gm::STRIDE * (int)sizeof(uint32_t);
  NTT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const unsigned long long grid = row_grid(P, p.batch);
  if (grid > 0x7fffffffull) return nttb200_fail(NTTB200_EPARAM, "batch too large for one launch");
  return launch_pdl_large(kernel, (unsigned)grid, ROW_WARPS * 32, smem, st, p);
}

template <int DIR>
int rows_ntt(const nttb200_plan *P, LargeParams &p, cudaStream_t st) {
  using Gm = SmallGeom<LR>;
  auto kernel = large_rows_ntt_kernel<LR, LARGE_ARITH, ROW_WARPS, DIR>;
  const int smem = ROW_WARPS * Gm::PPW * Gm::STRIDE * (int)sizeof(uint32_t);
  const unsigned long long grid = row_grid(P, p.batch);
  if (grid > 0x7fffffffull) return nttb200_fail(NTTB200_EPARAM, "batch too large for one launch");
  kernel<<<(unsigned)grid, ROW_WARPS * 32, smem, st>>>(p);
  nttb200_count_launch(1);
  NTT_CUDA(cudaGetLastError());
  return 0;
}
}  // namespace

#define LARGE_CAT2(a, b) a##b
#define LARGE_CAT(a, b) LARGE_CAT2(a, b)

/* one batch chunk of the product: ta/tb = scratch for a', b' (c' reuses ta) */
int LARGE_CAT(launch_polymul_large_chunk_, LARGE_NAME)(const nttb200_plan *P, uint32_t *c, const uint32_t *a,
                                                       const uint32_t *b, uint32_t *ta, uint32_t *tb,
                                                       size_t batch, cudaStream_t st, int early_loads) {
  const bool cyclic = (P->flags & NTTB200_PLAN_CYCLIC) != 0;
  const DevTable &fwd = cyclic ? P->fwd_plain : P->fwd_mixed;
  const DevTable &inv = cyclic ? P->inv_plain : P->inv_mixed;
  LargeParams p{};
  fill_common(p, P, &fwd, &inv, batch);
  int rc;
  p.src[0] = a; p.src[1] = b; p.dst[0] = ta; p.dst[1] = tb; p.nops = 2;
  fill_utw(p, &fwd);
  p.pdl = early_loads ? 1u : 0u;
  if ((rc = cols_fwd_any(P, p, st))) return rc;
  p.pdl = 0;
  p.src[0] = ta; p.src[1] = tb; p.dst[0] = ta; p.dst[1] = nullptr; p.nops = 1;
  if ((rc = rows_polymul(P, p, st))) return rc;
  /* n^-1 * 2^32: the 2^32 cancels the Montgomery 2^-32 of the pointwise product */
  const uint64_t fs = (uint64_t)P->n_inv * ((1ull << 32) % P->q) % P->q;
  p.last_x = lshoup_pair(fs, P->q);
  p.last_y = lshoup_pair(fs * inv.h[1].x, P->q);
  p.src[0] = ta; p.dst[0] = c;
  fill_utw(p, &inv);
  return cols_inv_any(P, p, st);
}

/* ---- the fused persistent cluster kernel (ntt_large_fused.cuh), n = 2^15 and 2^16 ---------------- */
namespace {
template <int K1>
int fused_config(cudaLaunchConfig_t &cfg, cudaLaunchAttribute *attr, unsigned grid, cudaStream_t st, int *smem_out) {
  using G = FusedGeom<K1>;
  using Gm = SmallGeom<LR>;
  const int smem = (int)sizeof(uint32_t) * std::max(G::COL_SMEM_WORDS, FUSED_WARPS * 2 * Gm::PPW * Gm::STRIDE);
  auto kernel = large_fused_polymul_kernel<K1, LARGE_ARITH>;
  if (smem > 48 * 1024)
    NTT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = FUSED_CLUSTER;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg = cudaLaunchConfig_t{};
  cfg.gridDim = dim3(grid, 1, 1);
  cfg.blockDim = dim3(FUSED_WARPS * 32, 1, 1);
  cfg.dynamicSmemBytes = (size_t)smem;
  cfg.stream = st;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  *smem_out = smem;
  return 0;
}
template <int K1>
int fused_clusters(int *out) {
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  int smem = 0, n = 0;
  int rc = fused_config<K1>(cfg, attr, FUSED_CLUSTER, nullptr, &smem);
  if (rc) return rc;
  NTT_CUDA(cudaOccupancyMaxActiveClusters(&n, large_fused_polymul_kernel<K1, LARGE_ARITH>, &cfg));
  *out = n;
  return 0;
}
template <int K1>
int fused_launch(const FusedParams &p, unsigned clusters, cudaStream_t st) {
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  int smem = 0;
  int rc = fused_config<K1>(cfg, attr, clusters * FUSED_CLUSTER, st, &smem);
  if (rc) return rc;
  NTT_CUDA(cudaLaunchKernelEx(&cfg, large_fused_polymul_kernel<K1, LARGE_ARITH>, p));
  nttb200_count_launch(1);
  return 0;
}
}  // namespace

namespace {
template <int K1>
int flow_launch(const FlowParams &q, int sm_count, cudaStream_t st) {
  using G = FusedGeom<K1>;
  using Gm = SmallGeom<LR>;
  const int smem = (int)sizeof(uint32_t) * std::max(G::COL_SMEM_WORDS, FUSED_WARPS * 2 * Gm::PPW * Gm::STRIDE);
  auto kernel = large_flow_polymul_kernel<K1, LARGE_ARITH>;
  static int per_sm = 0;
  if (!per_sm) {
    if (smem > 48 * 1024)
      NTT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    NTT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, FUSED_WARPS * 32, smem));
    if (per_sm < 1) per_sm = 1;
  }
  const unsigned long long tasks = (q.f.batch + 2 * FLOW_DELAY) * FLOW_TASKS;
  const unsigned grid = (unsigned)std::min<unsigned long long>(tasks, (unsigned long long)per_sm * sm_count);
  kernel<<<grid, FUSED_WARPS * 32, smem, st>>>(q);
  nttb200_count_launch(1);
  NTT_CUDA(cudaGetLastError());
  return 0;
}
}  // namespace

/* the dataflow version: ctl = 1 + 3 * batch zeroed words, scratch = FLOW_SLOTS x 2n words */
int LARGE_CAT(launch_polymul_large_flow_, LARGE_NAME)(const nttb200_plan *P, uint32_t *c, const uint32_t *a,
                                                      const uint32_t *b, uint32_t *scratch, unsigned *ctl,
                                                      size_t batch, cudaStream_t st) {
  const bool cyclic = (P->flags & NTTB200_PLAN_CYCLIC) != 0;
  const DevTable &fwd = cyclic ? P->fwd_plain : P->fwd_mixed;
  const DevTable &inv = cyclic ? P->inv_plain : P->inv_mixed;
  FlowParams q{};
  FusedParams &p = q.f;
  p.a = a; p.b = b; p.c = c; p.scratch = scratch;
  p.tab = fwd.d; p.tab_inv = inv.d;
  p.batch = batch;
  p.m = P->m;
  p.one = lshoup_pair(1, P->q);
  const uint64_t fs = (uint64_t)P->n_inv * ((1ull << 32) % P->q) % P->q;
  p.last_x = lshoup_pair(fs, P->q);
  p.last_y = lshoup_pair(fs * inv.h[1].x, P->q);
  for (int i = 1; i < 32; i++) { p.ufwd[i] = fwd.h[i]; p.uinv[i] = inv.h[i]; }
  q.ctl = ctl;
  switch (P->logn - LR) {
    case 8: return flow_launch<8>(q, P->sm_count, st);
    default: return nttb200_fail(NTTB200_EPARAM, "the dataflow large-n kernel covers n = 2^16");
  }
}
int LARGE_CAT(large_flow_slots_, LARGE_NAME)(void) { return FLOW_SLOTS; }

/* how many clusters of the fused kernel the device holds at once (0: this n is not covered) */
int LARGE_CAT(large_fused_clusters_, LARGE_NAME)(const nttb200_plan *P, int *clusters) {
  *clusters = 0;
  switch (P->logn - LR) {
    case 7: return fused_clusters<7>(clusters);
    case 8: return fused_clusters<8>(clusters);
    default: return 0;
  }
}
/* the whole batch in one launch; scratch = clusters x 2n words */
int LARGE_CAT(launch_polymul_large_fused_, LARGE_NAME)(const nttb200_plan *P, uint32_t *c, const uint32_t *a,
                                                       const uint32_t *b, uint32_t *scratch, unsigned clusters,
                                                       size_t batch, cudaStream_t st) {
  const bool cyclic = (P->flags & NTTB200_PLAN_CYCLIC) != 0;
  const DevTable &fwd = cyclic ? P->fwd_plain : P->fwd_mixed;
  const DevTable &inv = cyclic ? P->inv_plain : P->inv_mixed;
  FusedParams p{};
  p.a = a; p.b = b; p.c = c; p.scratch = scratch;
  p.tab = fwd.d; p.tab_inv = inv.d;
  p.batch = batch;
  p.m = P->m;
  p.one = lshoup_pair(1, P->q);
  /* n^-1 * 2^32: the 2^32 cancels the Montgomery 2^-32 of the pointwise product */
  const uint64_t fs = (uint64_t)P->n_inv * ((1ull << 32) % P->q) % P->q;
  p.last_x = lshoup_pair(fs, P->q);
  p.last_y = lshoup_pair(fs * inv.h[1].x, P->q);
  for (int i = 1; i < 32; i++) { p.ufwd[i] = fwd.h[i]; p.uinv[i] = inv.h[i]; }
  switch (P->logn - LR) {
    case 7: return fused_launch<7>(p, clusters, st);
    case 8: return fused_launch<8>(p, clusters, st);
    default: return nttb200_fail(NTTB200_EPARAM, "the fused large-n kernel covers n = 2^15, 2^16");
  }
}

/* standalone transform in place on a: dir 0 = forward CT std->rev with `tab`, dir 1 = inverse
 * GS rev->std; scale = multiply by n^-1 (inverse only).  Canonical output. */
int LARGE_CAT(launch_ntt_large_, LARGE_NAME)(const nttb200_plan *P, const DevTable &tab, int dir, int scale,
                                             uint32_t *a, size_t batch, cudaStream_t st) {
  LargeParams p{};
  fill_common(p, P, dir == 0 ? &tab : nullptr, dir == 1 ? &tab : nullptr, batch);
  p.src[0] = a; p.dst[0] = a;
  fill_utw(p, &tab);
  int rc;
  if (dir == 0) {
    if ((rc = cols_fwd_any(P, p, st))) return rc;
    return rows_ntt<0>(P, p, st);
  }
  if ((rc = rows_ntt<1>(P, p, st))) return rc;
  if (scale) {
    p.last_x = lshoup_pair(P->n_inv, P->q);
    p.last_y = lshoup_pair((uint64_t)P->n_inv * tab.h[1].x, P->q);
  }
  return cols_inv_any(P, p, st);
}
