/*
 * small_dispatch.inl -- instantiates the n <= 1024 kernels for ONE arithmetic class
 * (SMALL_ARITH) and defines its two dispatch functions.  Included by small_lazy.cu,
 * small_harvey.cu and small_canon.cu so the three classes compile in parallel.
 */
#include <cuda_runtime.h>

#include <algorithm>
#include <mutex>

#include "ntt_small.cuh"
#include "plan.h"

namespace {
using namespace nttb200;

template <int L> struct SmallCfg {           /* launch shape per size */
  static constexpr int WARPS = (L >= 9) ? 4 : 8;
  static constexpr int MINB = 2;
  static constexpr bool TWREG = (L <= 8);    /* lane twiddles live in registers across tiles */
};

template <int R>
void fill_params(SmallParams<R> &p, const nttb200_plan *P, const DevTable *fwd, const DevTable *inv) {
  p.m = P->m;
  p.tw_fwd = fwd ? fwd->d : nullptr;
  p.tw_inv = inv ? inv->d : nullptr;
  for (int i = 0; i < (1 << R); i++) {
    p.u.fwd[i] = (fwd && (size_t)i < fwd->h.size()) ? fwd->h[i] : make_uint2(0, 0);
    p.u.inv[i] = (inv && (size_t)i < inv->h.size()) ? inv->h[i] : make_uint2(0, 0);
  }
}

inline uint2 shoup_pair(uint64_t w, uint32_t q) {
  w %= q;
  return make_uint2((uint32_t)w, (uint32_t)((w << 32) / q));
}

template <typename K>
int grid_for(K kernel, int threads, int smem, int sm_count, unsigned long long tiles_per_block_unit,
             int *grid) {
  /* asked once per kernel and device, not per launch (kernels of one signature share K: key by address) */
  struct Seen { const void *k; int dev; int v; };
  static Seen seen[64];
  static int nseen = 0;
  static std::mutex mu;
  int dev = 0, per_sm = 0;
  cudaGetDevice(&dev);
  {
    std::lock_guard<std::mutex> lock(mu);
    for (int i = 0; i < nseen; i++)
      if (seen[i].k == (const void *)kernel && seen[i].dev == dev) per_sm = seen[i].v;
  }
  if (!per_sm) {
    NTT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    NTT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem));
    if (per_sm < 1) return nttb200_fail(NTTB200_ECUDA, "kernel does not fit on an SM");
    std::lock_guard<std::mutex> lock(mu);
    if (nseen < 64) seen[nseen++] = Seen{(const void *)kernel, dev, per_sm};
  }
  unsigned long long want = tiles_per_block_unit;
  unsigned long long cap = (unsigned long long)sm_count * per_sm;
  /* up to 4 times the resident CTAs while a warp still gets 6 tiles or more: the hardware hands
   * CTAs to the SMs as they free up, which evens out the SM-to-SM spread (small_plant.cu) */
  cap *= std::min<unsigned long long>(4, std::max<unsigned long long>(1, want / (cap * 6)));
  *grid = (int)(want < cap ? (want ? want : 1) : cap);
  return 0;
}

/* launch with programmatic stream serialization; `flags` carries SMALL_FLAG_NOWAIT when the launch is
 * independent of what is in flight on the stream (plan.h: nttb200_launch_independent) */
template <typename K, typename PT>
int launch_pdl(K kernel, int grid, int threads, int smem, cudaStream_t st, const PT &p) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
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

template <int L>
int run_polymul(const nttb200_plan *P, uint32_t *c, const uint32_t *a, const uint32_t *b, size_t batch,
                cudaStream_t st) {
  using Gm = SmallGeom<L>;
  using Cfg = SmallCfg<L>;
  SmallParams<Gm::R> p{};
  const DevTable &fwd = (P->flags & NTTB200_PLAN_CYCLIC) ? P->fwd_plain : P->fwd_mixed;
  const DevTable &inv = (P->flags & NTTB200_PLAN_CYCLIC) ? P->inv_plain : P->inv_mixed;
  fill_params<Gm::R>(p, P, &fwd, &inv);
  p.a = a; p.b = b; p.c = c; p.batch = batch;
  /* n^-1 * 2^32: the 2^32 cancels the Montgomery 2^-32 of the pointwise product */
  const uint64_t fs = (uint64_t)P->n_inv * ((1ull << 32) % P->q) % P->q;
  p.last_x = shoup_pair(fs, P->q);
  p.last_y = shoup_pair(fs * inv.h[1].x, P->q);
  p.flags = SMALL_FLAG_SCALE_LAST;
  auto kernel = polymul_small_kernel<L, SMALL_ARITH, Cfg::WARPS, Cfg::MINB, Cfg::TWREG>;
  const int smem = Cfg::WARPS * 2 * Gm::PPW * Gm::STRIDE * (int)sizeof(uint32_t);
  const unsigned long long tiles = (batch + Gm::PPW - 1) / Gm::PPW;
  int grid = 0;
  int rc = grid_for(kernel, Cfg::WARPS * 32, smem, P->sm_count, (tiles + Cfg::WARPS - 1) / Cfg::WARPS, &grid);
  if (rc) return rc;
  const size_t bytes = batch * Gm::N * sizeof(uint32_t);
  if (nttb200_launch_independent(st, a, bytes, b, bytes, c, bytes)) p.flags |= SMALL_FLAG_NOWAIT;
  return launch_pdl(kernel, grid, Cfg::WARPS * 32, smem, st, p);
}

template <int L, int DIR>
int run_ntt(const nttb200_plan *P, const DevTable &tab, int scale, uint32_t *a, size_t batch,
            cudaStream_t st) {
  using Gm = SmallGeom<L>;
  constexpr int WARPS = SmallCfg<L>::WARPS;
  SmallParams<Gm::R> p{};
  fill_params<Gm::R>(p, P, DIR == 0 ? &tab : nullptr, DIR == 1 ? &tab : nullptr);
  p.a = nullptr; p.b = nullptr; p.c = a; p.batch = batch;
  if (DIR == 1) {
    if (scale) {
      p.last_x = shoup_pair(P->n_inv, P->q);
      p.last_y = shoup_pair((uint64_t)P->n_inv * tab.h[1].x, P->q);
      p.flags = SMALL_FLAG_SCALE_LAST;
    } else {
      p.last_x = make_uint2(0, 0);
      p.last_y = tab.h[1];
    }
  }
  auto kernel = ntt_small_kernel<L, SMALL_ARITH, WARPS, DIR>;
  const int smem = WARPS * Gm::PPW * Gm::STRIDE * (int)sizeof(uint32_t);
  const unsigned long long tiles = (batch + Gm::PPW - 1) / Gm::PPW;
  int grid = 0;
  int rc = grid_for(kernel, WARPS * 32, smem, P->sm_count, (tiles + WARPS - 1) / WARPS, &grid);
  if (rc) return rc;
  /* in place: the array is both read and written */
  if (nttb200_launch_independent(st, a, batch * Gm::N * sizeof(uint32_t), nullptr, 0, a, batch * Gm::N * sizeof(uint32_t)))
    p.flags |= SMALL_FLAG_NOWAIT;
  return launch_pdl(kernel, grid, WARPS * 32, smem, st, p);
}

template <int L>
int info(int *regs, int *smem_bytes, int *blocks_per_sm) {
  using Gm = SmallGeom<L>;
  using Cfg = SmallCfg<L>;
  auto kernel = polymul_small_kernel<L, SMALL_ARITH, Cfg::WARPS, Cfg::MINB, Cfg::TWREG>;
  cudaFuncAttributes at;
  NTT_CUDA(cudaFuncGetAttributes(&at, kernel));
  const int smem = Cfg::WARPS * 2 * Gm::PPW * Gm::STRIDE * (int)sizeof(uint32_t);
  NTT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  int per_sm = 0;
  NTT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, Cfg::WARPS * 32, smem));
  *regs = at.numRegs; *smem_bytes = smem; *blocks_per_sm = per_sm;
  return 0;
}
}  // namespace

#define SMALL_CAT2(a, b) a##b
#define SMALL_CAT(a, b) SMALL_CAT2(a, b)

#define SMALL_SWITCH(expr_of_L)                      \
  switch (P->logn) {                                 \
    case 3: { constexpr int L = 3; expr_of_L; }      \
    case 4: { constexpr int L = 4; expr_of_L; }      \
    case 5: { constexpr int L = 5; expr_of_L; }      \
    case 6: { constexpr int L = 6; expr_of_L; }      \
    case 7: { constexpr int L = 7; expr_of_L; }      \
    case 8: { constexpr int L = 8; expr_of_L; }      \
    case 9: { constexpr int L = 9; expr_of_L; }      \
    case 10: { constexpr int L = 10; expr_of_L; }    \
    default: return nttb200_fail(NTTB200_EPARAM, "small kernels cover 8 <= n <= 1024"); \
  }

int SMALL_CAT(launch_polymul_small_, SMALL_NAME)(const nttb200_plan *P, uint32_t *c, const uint32_t *a,
                                                 const uint32_t *b, size_t batch, cudaStream_t st) {
  SMALL_SWITCH(return run_polymul<L>(P, c, a, b, batch, st))
}

int SMALL_CAT(launch_ntt_small_, SMALL_NAME)(const nttb200_plan *P, const DevTable &tab, int dir, int scale,
                                             uint32_t *a, size_t batch, cudaStream_t st) {
  if (dir == 0) { SMALL_SWITCH(return (run_ntt<L, 0>(P, tab, scale, a, batch, st))) }
  SMALL_SWITCH(return (run_ntt<L, 1>(P, tab, scale, a, batch, st)))
}

int SMALL_CAT(small_kernel_info_, SMALL_NAME)(const nttb200_plan *P, int *regs, int *smem_bytes,
                                              int *blocks_per_sm) {
  SMALL_SWITCH(return info<L>(regs, smem_bytes, blocks_per_sm))
}
