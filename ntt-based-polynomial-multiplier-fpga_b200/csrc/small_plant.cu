/* n <= 1024 fused product kernel for half-word moduli (Plantard arithmetic, ntt_small_plant.cuh) */
#include <cuda_runtime.h>
#include <stdlib.h>

#include <algorithm>

#include "ntt_small_plant.cuh"
#include "ntt_small_splant.cuh"
#include "ntt_splant_wide.cuh"
#include "plan.h"

namespace {
using namespace nttb200;

#ifndef PLANT_WARPS
#define PLANT_WARPS 8
#endif
#ifndef PLANT_BIG_WARPS
#define PLANT_BIG_WARPS 4      /* warps per CTA at n >= 512 (tuning knob; with 12, PLANT_BIG_CTAS 1) */
#endif
#ifndef PLANT_BIG_CTAS
#define PLANT_BIG_CTAS 2
#endif
#ifndef SPLANT_SMALL_CTAS
#define SPLANT_SMALL_CTAS 2    /* resident CTAs the signed kernel is compiled for at n <= 256 (register cap) */
#endif
template <int L> struct PlantCfg {
  static constexpr int WARPS = (L >= 9) ? PLANT_BIG_WARPS : PLANT_WARPS;
  static constexpr bool TWREG = (L <= 8);
};
/* resident CTAs per SM the kernel is compiled for (register cap).  Measured on B200 (c2:
 * 996 vs 939 M polymul/s, c3: 1095 vs 1026, c4: 218 vs 213): 2 CTAs with 127 registers beat 3
 * CTAs with 80 registers and a small spill.  NTTB200_PLANT_MINB=2|3 overrides (tuning knob). */
int plant_minb(int L) {
  static int env = -1;
  if (env < 0) {
    const char *e = getenv("NTTB200_PLANT_MINB");
    env = e ? atoi(e) : 0;
  }
  if (env == 2 || env == 3) return env;
  (void)L;
  return 2;
}

/* NTTB200_PDL=0 turns programmatic dependent launch off (tuning / debugging knob) */
int plant_pdl() {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("NTTB200_PDL");
    v = e ? atoi(e) != 0 : 1;
  }
  return v;
}

/* tail scheduler knob: NTTB200_PLANT_DYN_PCT = per cent of the tiles handed out dynamically
 * (0 = static assignment only).  c4 on B200: 0 % 236, 8 % 245, 16 % 259, 24 % 263.5, 32 % 262,
 * 50 % 260, 100 % 253 M polymul/s */
int plant_dyn_pct() {
  static int v = -1;
  if (v < 0) { const char *e = getenv("NTTB200_PLANT_DYN_PCT"); v = e ? atoi(e) : 24; if (v < 0) v = 0; if (v > 100) v = 100; }
  return v;
}
template <int L, int MINB, typename IO = uint32_t, typename OIO = IO>
int run_plant(const nttb200_plan *P, void *c, const void *a, const void *b, size_t batch,
              cudaStream_t st) {
  using Gm = SmallGeom<L>;
  using Pg = PlantGeom<L, IO>;
  using Cfg = PlantCfg<L>;
  const bool cyclic = (P->flags & NTTB200_PLAN_CYCLIC) != 0;
  const DevTable &fwd = cyclic ? P->fwd_plain : P->fwd_mixed;
  const DevTable &inv = cyclic ? P->inv_plain : P->inv_mixed;
  PlantParams<Gm::R> p{};
  p.a = a; p.b = b; p.c = c; p.batch = batch;
  p.tw_fwd = fwd.d1; p.tw_inv = inv.d1;
  p.q = P->q; p.qinv = P->m.qinv;
  const uint32_t q = P->q;
  /* -n^-1 2^32: cancels the -2^-32 of the Plantard pointwise product */
  const uint64_t fs = (q - (uint64_t)P->n_inv * ((1ull << 32) % q) % q) % q;
  p.last_x = nttb200_plant_form((uint32_t)fs, q, p.qinv);
  p.last_y = nttb200_plant_form((uint32_t)(fs * inv.h[1].x % q), q, p.qinv);
  for (int i = 0; i < (int)(sizeof p.qmul / sizeof p.qmul[0]); i++) p.qmul[i] = (uint32_t)i * q;
  for (int i = 0; i < (1 << Gm::R); i++) {
    p.ufwd[i] = (size_t)i < fwd.h1.size() ? fwd.h1[i] : 0;
    p.uinv[i] = (size_t)i < inv.h1.size() ? inv.h1[i] : 0;
  }
  auto kernel = polymul_plant_kernel<L, Cfg::WARPS, MINB, Cfg::TWREG, IO, OIO>;
  const int smem = Cfg::WARPS * Pg::WARP_WORDS * (int)sizeof(uint32_t);
  /* attributes and occupancy are asked once per kernel instance and device, not per launch */
  static int per_sm_dev[64] = {0};
  int &per_sm = per_sm_dev[P->device & 63];
  if (!per_sm) {
    int v = 0;
    NTT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    NTT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, kernel, Cfg::WARPS * 32, smem));
    if (v < 1) return nttb200_fail(NTTB200_ECUDA, "plant kernel does not fit on an SM");
    per_sm = v;
  }
  const unsigned long long tiles = (batch + Gm::PPW - 1) / Gm::PPW;
  const unsigned long long want = (tiles + Cfg::WARPS - 1) / Cfg::WARPS;
  /* Sizes that keep the static loop (n <= 256) launch up to 4 times the resident CTAs, as long as
   * every warp still gets 6 tiles or more: the hardware then hands CTAs to the SMs as they free up,
   * which evens out the SM-to-SM spread at the price of one start-up per CTA.
   * Measured: batch 2^20 x1 1 349, x2 1 367, x3 1 375, x4 1 378, x8 1 375 M polymul/s; batch 2^16 x1
   * 1 223, x2 1 245, x3 1 221, x4 1 210, x8 986 M.  NTTB200_PLANT_GRIDX overrides. */
  static const int gridx_env = [] { const char *e = getenv("NTTB200_PLANT_GRIDX"); int v = e ? atoi(e) : 0; return v < 0 ? 0 : (v > 64 ? 64 : v); }();
  unsigned long long gridx = 1;
  if (gridx_env > 0) gridx = (unsigned long long)gridx_env;
  else if (L < PLANT_DYN_MINL)
    gridx = std::min<unsigned long long>(4, std::max<unsigned long long>(1, tiles / ((unsigned long long)P->sm_count * per_sm * Cfg::WARPS * 6)));
  const unsigned long long cap = (unsigned long long)P->sm_count * per_sm * gridx;
  const int grid = (int)(want < cap ? (want ? want : 1) : cap);
  /* n >= 2^PLANT_DYN_MINL, many more tiles than warps: the last quarter of the tiles is handed out
   * from a device counter (PlantParams::sched) */
  const unsigned long long warps = (unsigned long long)grid * Cfg::WARPS;
  p.sched = nullptr;
  if (L >= PLANT_DYN_MINL && P->sched_ring && plant_dyn_pct() > 0 && tiles > 4 * warps) {
    const unsigned long long stat = tiles * (100 - plant_dyn_pct()) / 100 / warps;      /* whole rounds */
    p.static_rounds = (uint32_t)std::min<unsigned long long>(std::max<unsigned long long>(stat, 2), 0x7fffffffull);
    p.sched = P->sched_ring + 2 * (size_t)(P->sched_seq.fetch_add(1) % NTTB200_SCHED_SLOTS);
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(Cfg::WARPS * 32);
  cfg.dynamicSmemBytes = (size_t)smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = plant_pdl() ? 1 : 0;
  p.nowait = (plant_pdl() && nttb200_launch_independent(st, a, batch * Gm::N * sizeof(IO), b, batch * Gm::N * sizeof(IO),
                                                        c, batch * Gm::N * sizeof(OIO))) ? 1u : 0u;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  NTT_CUDA(cudaLaunchKernelEx(&cfg, kernel, p));
  nttb200_count_launch(1);
  NTT_CUDA(cudaGetLastError());
  return 0;
}

/* Which of the two Plantard kernels serves a product.  The signed kernel of ntt_small_splant.cuh
 * (five-instruction butterflies levelled over the two integer pipes, incomplete transform with a pair
 * multiplication) is the default at every size: measured on B200 c2 1 340 -> 1 473,
 * c3 1 407 -> 1 484, c4 263 -> 273 M polymul/s, and 1 325 -> 1 392 M at c2 when the load lasts a second
 * and the chip runs into its power cap.  NTTB200_PLANT_SIGNED=0 brings the unsigned kernel of
 * ntt_small_plant.cuh back (2: signed for n >= 512 only, 3: for n <= 256 only); read per call so that
 * the tests cover both. */
int plant_signed(int logn) {
  const char *e = getenv("NTTB200_PLANT_SIGNED");
  const int v = e ? atoi(e) : 1;
  return v == 1 || (v == 2 && logn >= 9) || (v == 3 && logn <= 8);
}

/* the signed kernel */
template <int L, typename IO = uint32_t, typename OIO = IO>
int run_splant(const nttb200_plan *P, void *c, const void *a, const void *b, size_t batch, cudaStream_t st) {
  using Gm = SmallGeom<L>;
  using Pg = PlantGeom<L, IO>;
  using Cfg = PlantCfg<L>;
  constexpr int WARPS = Cfg::WARPS;
  const bool cyclic = (P->flags & NTTB200_PLAN_CYCLIC) != 0;
  const DevTable &fwd = cyclic ? P->fwd_plain : P->fwd_mixed;
  const DevTable &inv = cyclic ? P->inv_plain : P->inv_mixed;
  SPlantParams<Gm::R> p{};
  p.a = a; p.b = b; p.c = c; p.batch = batch;
  constexpr int V = SpDrop<L>::V;                   /* stages the transforms leave to the group multiplication */
  p.tw_fwd = fwd.d2; p.tw_inv = inv.d2; p.zeta = fwd.d3 + (Gm::N >> V);
  const uint32_t q = P->q;
  p.q = q; p.qinv = P->m.qinv;
  /* |Y W| <= mmax is what the second product tolerates (ntt_small_splant.cuh); the kernel multiplies
   * differences up to 16 q by |W| <= q/2 */
  const uint64_t mmax = ((1ull << 32) - 65536ull * (q + 4)) / 2;
  if (8ull * q * q > mmax) return nttb200_fail(NTTB200_EPARAM, "q=%u is too large for the signed Plantard kernel", q);
  /* the group multiplication reduces sums of 2^V raw products a b with |a| <= q/2 + 20 (Barrett step),
   * |b| <= (L + 2 - V) q / 2 (forward values after L - V stages), plus one product of two centred values */
  /* (three stages left out: both operands come out of the last forward stage within q + 20 of zero) */
  const uint64_t gsum = SpDrop<L>::XR ? (uint64_t)(1 << V) * (q + 21) * (q + 21)
                                      : (uint64_t)(1 << V) * (q / 2 + 21) * ((uint64_t)(L + 2 - V) * q / 2 + 1);
  if (V > 0 && gsum + (uint64_t)q * q / 4 > mmax)
    return nttb200_fail(NTTB200_EPARAM, "q=%u is too large for the signed Plantard kernel at n=%u", q, Gm::N);
  p.dd = (uint32_t)((mmax + 65535ull * q + 65535ull) / 65536ull);
  p.cbar = (uint32_t)(((1ull << SP_RED_SHIFT) + q / 2) / q);
  /* -n^-1 2^32: cancels the -2^-32 of the Plantard pointwise product; (n / 2^V)^-1 when the inverse
   * network runs without its first V stages (SPLANT_INCOMPLETE) */
  const uint64_t ninv = (uint64_t)P->n_inv * (1u << V) % q;
  const uint64_t fs = (q - ninv * ((1ull << 32) % q) % q) % q;
  p.last_x = nttb200_plant_form_centred((uint32_t)fs, q, p.qinv);
  p.last_y = nttb200_plant_form_centred((uint32_t)(fs * inv.h[1].x % q), q, p.qinv);
  for (int i = 0; i < (1 << Gm::R); i++) {
    p.ufwd[i] = (size_t)i < fwd.h2.size() ? fwd.h2[i] : 0;
    p.uinv[i] = (size_t)i < inv.h2.size() ? inv.h2[i] : 0;
  }
  auto kernel = polymul_splant_kernel<L, WARPS, (L >= 9) ? PLANT_BIG_CTAS : SPLANT_SMALL_CTAS, Cfg::TWREG, IO, OIO>;
  const int smem = WARPS * Pg::WARP_WORDS * (int)sizeof(uint32_t);
  static int per_sm_dev[64] = {0};
  int &per_sm = per_sm_dev[P->device & 63];
  if (!per_sm) {
    int v = 0;
    NTT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    NTT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, kernel, WARPS * 32, smem));
    if (v < 1) return nttb200_fail(NTTB200_ECUDA, "signed plant kernel does not fit on an SM");
    per_sm = v;
  }
  const unsigned long long tiles = (batch + Gm::PPW - 1) / Gm::PPW;
  const unsigned long long want = (tiles + WARPS - 1) / WARPS;
  static const int gridx_env = [] { const char *e = getenv("NTTB200_PLANT_GRIDX"); int v = e ? atoi(e) : 0; return v < 0 ? 0 : (v > 64 ? 64 : v); }();
  unsigned long long gridx = 1;
  if (gridx_env > 0) gridx = (unsigned long long)gridx_env;
  else if (L < PLANT_DYN_MINL)
    gridx = std::min<unsigned long long>(4, std::max<unsigned long long>(1, tiles / ((unsigned long long)P->sm_count * per_sm * WARPS * 6)));
  const unsigned long long cap = (unsigned long long)P->sm_count * per_sm * gridx;
  const int grid = (int)(want < cap ? (want ? want : 1) : cap);
  /* tail scheduler, as run_plant */
  const unsigned long long warps = (unsigned long long)grid * WARPS;
  p.sched = nullptr;
  if (L >= PLANT_DYN_MINL && P->sched_ring && plant_dyn_pct() > 0 && tiles > 4 * warps) {
    const unsigned long long stat = tiles * (100 - plant_dyn_pct()) / 100 / warps;
    p.static_rounds = (uint32_t)std::min<unsigned long long>(std::max<unsigned long long>(stat, 2), 0x7fffffffull);
    p.sched = P->sched_ring + 2 * (size_t)(P->sched_seq.fetch_add(1) % NTTB200_SCHED_SLOTS);
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(WARPS * 32);
  cfg.dynamicSmemBytes = (size_t)smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = plant_pdl() ? 1 : 0;
  p.nowait = (plant_pdl() && nttb200_launch_independent(st, a, batch * Gm::N * sizeof(IO), b, batch * Gm::N * sizeof(IO),
                                                        c, batch * Gm::N * sizeof(OIO))) ? 1u : 0u;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  NTT_CUDA(cudaLaunchKernelEx(&cfg, kernel, p));
  nttb200_count_launch(1);
  NTT_CUDA(cudaGetLastError());
  return 0;
}

/* n = 1024: the three-layout kernel of ntt_splant_wide.cuh (NTTB200_PLANT_N1024=0 keeps the one-layout-
 * per-phase kernel polymul_splant_kernel<10>).  The same kernel is instantiated for n = 512 and measured
 * there level with polymul_splant_kernel<9> (655 against 659 M polymul/s), so n = 512 takes it only with
 * NTTB200_PLANT_N1024=2 (the tests do). */
#ifndef SPLANT_N1024_WARPS
#define SPLANT_N1024_WARPS 4
#endif
#ifndef SPLANT_N1024_CTAS
#define SPLANT_N1024_CTAS 3
#endif
#ifndef SPLANT_N512_CTAS
#define SPLANT_N512_CTAS 6
#endif
int plant_n1024() {
  const char *e = getenv("NTTB200_PLANT_N1024");
  return e ? atoi(e) : 1;
}
template <int L, typename IO = uint32_t, typename OIO = IO>
int run_splant_wide(const nttb200_plan *P, void *c, const void *a, const void *b, size_t batch, cudaStream_t st) {
  using W = WideGeom<L>;
  constexpr int N = 1 << L, V = W::V;
  constexpr int WARPS = SPLANT_N1024_WARPS;
  const bool cyclic = (P->flags & NTTB200_PLAN_CYCLIC) != 0;
  const DevTable &fwd = cyclic ? P->fwd_plain : P->fwd_mixed;
  const DevTable &inv = cyclic ? P->inv_plain : P->inv_mixed;
  SPlantParams<4> p{};
  p.a = a; p.b = b; p.c = c; p.batch = batch;
  p.tw_fwd = fwd.d2; p.tw_inv = inv.d2; p.zeta = fwd.d3 + (N >> V);
  const uint32_t q = P->q;
  p.q = q; p.qinv = P->m.qinv;
  const uint64_t mmax = ((1ull << 32) - 65536ull * (q + 4)) / 2;       /* as run_splant */
  if (8ull * q * q > mmax ||
      (V >= 3 ? (uint64_t)(1 << V) * (q + 21) * (q + 21)
              : (uint64_t)(1 << V) * (q / 2 + 21) * ((uint64_t)(L + 2 - V) * q / 2 + 1)) + (uint64_t)q * q / 4 > mmax)
    return nttb200_fail(NTTB200_EPARAM, "q=%u is too large for the signed Plantard kernel at n=%d", q, N);
  p.dd = (uint32_t)((mmax + 65535ull * q + 65535ull) / 65536ull);
  p.cbar = (uint32_t)(((1ull << SP_RED_SHIFT) + q / 2) / q);
  const uint64_t ninv = (uint64_t)P->n_inv * (1u << V) % q;
  const uint64_t fs = (q - ninv * ((1ull << 32) % q) % q) % q;
  p.last_x = nttb200_plant_form_centred((uint32_t)fs, q, p.qinv);
  p.last_y = nttb200_plant_form_centred((uint32_t)(fs * inv.h[1].x % q), q, p.qinv);
  for (int i = 0; i < 16; i++) {
    p.ufwd[i] = (size_t)i < fwd.h2.size() ? fwd.h2[i] : 0;
    p.uinv[i] = (size_t)i < inv.h2.size() ? inv.h2[i] : 0;
  }
  auto kernel = polymul_splant_wide_kernel<L, WARPS, (L == 10) ? SPLANT_N1024_CTAS : SPLANT_N512_CTAS, IO, OIO>;
  const int smem = WARPS * (2 * (N * (int)sizeof(IO) / 4) + 2 * W::WK) * (int)sizeof(uint32_t);
  static int per_sm_dev[64] = {0};
  int &per_sm = per_sm_dev[P->device & 63];
  if (!per_sm) {
    int v = 0;
    NTT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    NTT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, kernel, WARPS * 32, smem));
    if (v < 1) return nttb200_fail(NTTB200_ECUDA, "n = %d plant kernel does not fit on an SM", N);
    per_sm = v;
  }
  const unsigned long long tiles = batch;
  const unsigned long long want = (tiles + WARPS - 1) / WARPS;
  const unsigned long long cap = (unsigned long long)P->sm_count * per_sm;
  const int grid = (int)(want < cap ? (want ? want : 1) : cap);
  const unsigned long long warps = (unsigned long long)grid * WARPS;
  p.sched = nullptr;
  if (P->sched_ring && plant_dyn_pct() > 0 && tiles > 4 * warps) {
    const unsigned long long stat = tiles * (100 - plant_dyn_pct()) / 100 / warps;
    p.static_rounds = (uint32_t)std::min<unsigned long long>(std::max<unsigned long long>(stat, 2), 0x7fffffffull);
    p.sched = P->sched_ring + 2 * (size_t)(P->sched_seq.fetch_add(1) % NTTB200_SCHED_SLOTS);
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(WARPS * 32);
  cfg.dynamicSmemBytes = (size_t)smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = plant_pdl() ? 1 : 0;
  p.nowait = (plant_pdl() && nttb200_launch_independent(st, a, batch * N * sizeof(IO), b, batch * N * sizeof(IO),
                                                        c, batch * N * sizeof(OIO))) ? 1u : 0u;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  NTT_CUDA(cudaLaunchKernelEx(&cfg, kernel, p));
  nttb200_count_launch(1);
  NTT_CUDA(cudaGetLastError());
  return 0;
}

template <int L>
int info_splant(int *regs, int *smem_bytes, int *blocks_per_sm) {
  using Pg = PlantGeom<L>;
  using Cfg = PlantCfg<L>;
  auto kernel = polymul_splant_kernel<L, Cfg::WARPS, (L >= 9) ? PLANT_BIG_CTAS : SPLANT_SMALL_CTAS, Cfg::TWREG>;
  cudaFuncAttributes at;
  NTT_CUDA(cudaFuncGetAttributes(&at, kernel));
  const int smem = Cfg::WARPS * Pg::WARP_WORDS * (int)sizeof(uint32_t);
  NTT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  int per_sm = 0;
  NTT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, Cfg::WARPS * 32, smem));
  *regs = at.numRegs; *smem_bytes = smem; *blocks_per_sm = per_sm;
  return 0;
}
template <int L, int MINB>
int info_plant(int *regs, int *smem_bytes, int *blocks_per_sm) {
  using Pg = PlantGeom<L>;
  using Cfg = PlantCfg<L>;
  auto kernel = polymul_plant_kernel<L, Cfg::WARPS, MINB, Cfg::TWREG>;
  cudaFuncAttributes at;
  NTT_CUDA(cudaFuncGetAttributes(&at, kernel));
  const int smem = Cfg::WARPS * Pg::WARP_WORDS * (int)sizeof(uint32_t);
  NTT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  int per_sm = 0;
  NTT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, Cfg::WARPS * 32, smem));
  *regs = at.numRegs; *smem_bytes = smem; *blocks_per_sm = per_sm;
  return 0;
}
}  // namespace

#define PLANT_SWITCH(expr_of_L)                      \
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

#define SPLANT_SWITCH(expr_of_L)                     \
  switch (P->logn) {                                 \
    case 3: { constexpr int L = 3; expr_of_L; }      \
    case 4: { constexpr int L = 4; expr_of_L; }      \
    case 5: { constexpr int L = 5; expr_of_L; }      \
    case 6: { constexpr int L = 6; expr_of_L; }      \
    case 7: { constexpr int L = 7; expr_of_L; }      \
    case 8: { constexpr int L = 8; expr_of_L; }      \
    case 9: { constexpr int L = 9; expr_of_L; }      \
    case 10: { constexpr int L = 10; expr_of_L; }    \
    default: break;                                  \
  }

int launch_polymul_small_plant(const nttb200_plan *P, uint32_t *c, const uint32_t *a, const uint32_t *b,
                               size_t batch, cudaStream_t st) {
  if (plant_signed(P->logn) && P->logn == 10 && plant_n1024()) return run_splant_wide<10>(P, c, a, b, batch, st);
  if (plant_signed(P->logn) && P->logn == 9 && plant_n1024() >= 2) return run_splant_wide<9>(P, c, a, b, batch, st);
  if (plant_signed(P->logn)) { SPLANT_SWITCH(return (run_splant<L>(P, c, a, b, batch, st))) }
  if (plant_minb(P->logn) == 2) { PLANT_SWITCH(return (run_plant<L, 2>(P, c, a, b, batch, st))) }
  PLANT_SWITCH(return (run_plant<L, 3>(P, c, a, b, batch, st)))
}
/* packed 16-bit operands and result (extension outside the reference API) */
int launch_polymul_small_plant_u16(const nttb200_plan *P, uint16_t *c, const uint16_t *a, const uint16_t *b,
                                   size_t batch, cudaStream_t st) {
  if (plant_signed(P->logn) && P->logn == 10 && plant_n1024()) return run_splant_wide<10, uint16_t>(P, c, a, b, batch, st);
  if (plant_signed(P->logn) && P->logn == 9 && plant_n1024() >= 2) return run_splant_wide<9, uint16_t>(P, c, a, b, batch, st);
  if (plant_signed(P->logn)) { SPLANT_SWITCH(return (run_splant<L, uint16_t>(P, c, a, b, batch, st))) }
  PLANT_SWITCH(return (run_plant<L, 2, uint16_t>(P, c, a, b, batch, st)))
}
/* 16-bit operands, 32-bit result: the wire pipeline of the host-buffer call narrows a and b on
 * the host and lets the kernel write the caller's int32 rows (nttb200.cu, polymul_batch_wire) */
int launch_polymul_small_plant_u16in(const nttb200_plan *P, uint32_t *c, const uint16_t *a, const uint16_t *b,
                                     size_t batch, cudaStream_t st) {
  if (plant_signed(P->logn) && P->logn == 10 && plant_n1024()) return run_splant_wide<10, uint16_t, uint32_t>(P, c, a, b, batch, st);
  if (plant_signed(P->logn) && P->logn == 9 && plant_n1024() >= 2) return run_splant_wide<9, uint16_t, uint32_t>(P, c, a, b, batch, st);
  if (plant_signed(P->logn)) { SPLANT_SWITCH(return (run_splant<L, uint16_t, uint32_t>(P, c, a, b, batch, st))) }
  PLANT_SWITCH(return (run_plant<L, 2, uint16_t, uint32_t>(P, c, a, b, batch, st)))
}
template <int L, int DIR>
int run_ntt_plant(const nttb200_plan *P, const DevTable &tab, int scale, uint32_t *a, size_t batch,
                  cudaStream_t st) {
  using Gm = SmallGeom<L>;
  constexpr int WARPS = PlantCfg<L>::WARPS;
  PlantParams<Gm::R> p{};
  p.a = nullptr; p.b = nullptr; p.c = a; p.batch = batch;
  p.tw_fwd = tab.d1; p.tw_inv = tab.d1;
  p.q = P->q; p.qinv = P->m.qinv;
  const uint32_t q = P->q;
  const uint64_t sc = scale ? P->n_inv : 1;
  p.last_x = nttb200_plant_form((uint32_t)sc, q, p.qinv);
  p.last_y = nttb200_plant_form((uint32_t)(sc * tab.h[1].x % q), q, p.qinv);
  for (int i = 0; i < (int)(sizeof p.qmul / sizeof p.qmul[0]); i++) p.qmul[i] = (uint32_t)i * q;
  for (int i = 0; i < (1 << Gm::R); i++) {
    const uint32_t v = (size_t)i < tab.h1.size() ? tab.h1[i] : 0;
    p.ufwd[i] = v;
    p.uinv[i] = v;
  }
  auto kernel = ntt_plant_kernel<L, WARPS, DIR>;
  const int smem = WARPS * ((DIR == 0 ? PlantGeom<L, uint32_t>::PF_WORDS : 0) + Gm::PPW * Gm::STRIDE) *
                   (int)sizeof(uint32_t);
  int per_sm = 0;
  NTT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  NTT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, WARPS * 32, smem));
  if (per_sm < 1) return nttb200_fail(NTTB200_ECUDA, "plant transform kernel does not fit on an SM");
  const unsigned long long tiles = (batch + Gm::PPW - 1) / Gm::PPW;
  const unsigned long long want = (tiles + WARPS - 1) / WARPS;
  unsigned long long cap = (unsigned long long)P->sm_count * per_sm;
  cap *= std::min<unsigned long long>(4, std::max<unsigned long long>(1, want / (cap * 6)));   /* as run_plant */
  const int grid = (int)(want < cap ? (want ? want : 1) : cap);
  p.nowait = (plant_pdl() && nttb200_launch_independent(st, a, batch * Gm::N * sizeof(uint32_t), nullptr, 0, a,
                                                        batch * Gm::N * sizeof(uint32_t))) ? 1u : 0u;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(WARPS * 32);
  cfg.dynamicSmemBytes = (size_t)smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = plant_pdl() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  NTT_CUDA(cudaLaunchKernelEx(&cfg, kernel, p));
  nttb200_count_launch(1);
  NTT_CUDA(cudaGetLastError());
  return 0;
}

/* standalone transform of a half-word-modulus plan: dir 0 forward CT, dir 1 inverse GS */
int launch_ntt_small_plant(const nttb200_plan *P, const DevTable &tab, int dir, int scale, uint32_t *a,
                           size_t batch, cudaStream_t st) {
  if (dir == 0) { PLANT_SWITCH(return (run_ntt_plant<L, 0>(P, tab, scale, a, batch, st))) }
  PLANT_SWITCH(return (run_ntt_plant<L, 1>(P, tab, scale, a, batch, st)))
}

int small_plant_signed(const nttb200_plan *P) { return plant_signed((int)P->logn); }

int small_kernel_info_plant(const nttb200_plan *P, int *regs, int *smem_bytes, int *blocks_per_sm) {
  if (plant_signed(P->logn)) { SPLANT_SWITCH(return (info_splant<L>(regs, smem_bytes, blocks_per_sm))) }
  if (plant_minb(P->logn) == 2) { PLANT_SWITCH(return (info_plant<L, 2>(regs, smem_bytes, blocks_per_sm))) }
  PLANT_SWITCH(return (info_plant<L, 3>(regs, smem_bytes, blocks_per_sm)))
}
