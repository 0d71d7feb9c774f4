/*
 * ntt_large.cuh -- multi-pass global-memory NTT for n = 2^L > 1024 (north_star item (g)).
 *
 * A polynomial is viewed as an n1 x n2 matrix (n1 = 2^K1 rows of n2 = 2^LR contiguous
 * coefficients, index i = row * n2 + col).  The Cooley-Tukey std->rev dataflow of the
 * reference (R/NTT/ntt.C:342-371, block j of stage t uses p[t + j]) splits exactly there:
 *
 *   stages 0 .. K1-1   touch the row bits only: n2 independent n1-point column transforms
 *                      whose twiddles p[2^s + j] do not depend on the column   -> COLUMN pass
 *   stages K1 .. L-1   stay inside one row: row j is an n2-point transform with the
 *                      twiddles p[2^(K1+s) + (j << s) + u]                      -> ROW pass
 *
 * and the Gentleman-Sande rev->std inverse (R/NTT/ntt.C:428-451) is the mirror image (rows
 * first, then columns).  A product is therefore three launches per batch chunk:
 *
 *   K1  column pass on a and b                      (read a,b          write a',b')
 *   K2  row pass fwd(a'), fwd(b'), pointwise, row pass inv   (read a',b'  write c')
 *   K3  column pass inv on c' with n^-1 folded into its last stage (read c' write c)
 *
 * a', b', c' live in a per-chunk scratch area (the host pipelines batch chunks over a few
 * internal streams so that the column pass of one chunk overlaps the row pass of the previous
 * one); the scratch is small enough to stay in L2, but with a 31-bit modulus the passes are
 * instruction-bound, not memory-bound.
 *
 * Column pass: one CTA owns 32 or 64 adjacent columns (one or two per lane: every global access
 * is a 128- or 256-byte row segment) and all n1 rows; the K1 stages run as one or two register phases
 * (RA then RB stages) with one exchange through shared memory.  Twiddle reads are
 * warp-uniform (they depend on the row bits only).
 * Row pass: the register/shared-memory machinery of ntt_small.cuh, one row of one
 * polynomial per half-warp, with the twiddles of row j fetched from the level table.
 */
#pragma once
#include <stdint.h>
#include "ntt_small.cuh"

namespace nttb200 {

struct LargeParams {
  const uint32_t *src[2];     /* K1: a, b    K2: a', b'    K3 / transforms: src[0]          */
  uint32_t *dst[2];           /* K1: a', b'  K2: c' in dst[0]   K3: c in dst[0]             */
  const uint2 *tab;           /* forward level table, n entries (w, floor(w 2^32/q))        */
  const uint2 *tab_inv;       /* inverse level table                                        */
  unsigned long long batch;   /* polynomials per operand                                    */
  uint32_t logn;
  uint32_t nops;              /* operands handled by a column launch (1 or 2)               */
  ModQ m;
  uint2 last_x;               /* K3: multiplier of the sum branch of the very last stage     */
  uint2 last_y;               /* K3: multiplier of its diff branch (twiddle p[1] * scale)    */
  uint2 one;                  /* (1, floor(2^32/q)): Shoup pair that only reduces            */
  uint32_t zero;              /* always 0 (modq_regs)                                        */
  uint32_t pdl;               /* programmatic dependent launch: bit 0 = the forward column pass may READ its
                                 operands before the previous kernel of its stream has finished (it
                                 then waits just before its first store into the scratch that kernel
                                 may still be reading)                                              */
  uint2 utw[32];              /* column passes: entries [1, 32) of the table of this launch's
                                 direction -- the twiddles of the register phase on the high row
                                 bits, whose indices are compile-time numbers (constant bank)   */
};

constexpr int LARGE_LR = 8;       /* rows of 2^8 coefficients: fixed, so row strides are immediates */

template <int K1, int CPL = 1>
struct ColGeom {
  static constexpr int RA = (K1 <= 4) ? K1 : (K1 + 1) / 2;   /* stages of the phase on the high row bits */
  static constexpr int RB = K1 - RA;                          /* stages of the phase on the low row bits  */
  static constexpr int NV = 1 << RA;                          /* column vectors per thread                 */
  static constexpr int WARPS = 1 << RB;
  static constexpr int GB = 1 << (RA - RB);                   /* RB-groups per thread (RB > 0)             */
  static constexpr int ROWS = 1 << K1;
  static constexpr int TILE_COLS = 32 * CPL;                  /* adjacent columns per CTA                  */
  static constexpr int LOG_TILE = (CPL == 2) ? 6 : 5;
  static constexpr int SMEM_BYTES = (RB > 0) ? ROWS * TILE_COLS * 4 : 0;
  /* resident CTAs asked of ptxas: every warp slot of the SM while a thread holds <= 16 vectors of one
   * column (32 registers then suffice); the latency of the strided row loads is hidden by warps only */
  static constexpr int MINB = (NV * CPL <= 16) ? ((2048 / (WARPS * 32)) > 32 ? 32 : (2048 / (WARPS * 32))) : 1;
};
/* columns per lane: 1 (32-bit accesses, 128-byte row segments per warp) or 2 (64-bit accesses,
 * 256-byte segments, half the load/store instructions; operands must be 8-byte aligned).
 * Measured at n = 2^16: two columns per lane help the classes with cheap butterflies (HARVEY
 * 1.55 -> 1.75 M polymul/s) and not CANON (1.51 -> 1.48 M), whose 8-instruction butterflies
 * dominate; large_dispatch.inl picks per class. */
template <int CPL> struct CVec { uint32_t v[CPL]; };
template <int CPL>
__device__ __forceinline__ CVec<CPL> cv_ldg(const uint32_t *p) {
  CVec<CPL> r;
  if (CPL == 1) {
    r.v[0] = __ldcg(p);             /* L2, not the non-coherent path: see gload_cols in ntt_small.cuh */
  } else {
    const uint2 t = __ldcg(reinterpret_cast<const uint2 *>(p));
    r.v[0] = t.x;
    r.v[CPL - 1] = t.y;
  }
  return r;
}
template <int CPL>
__device__ __forceinline__ CVec<CPL> cv_ld(const uint32_t *p) {
  CVec<CPL> r;
  if (CPL == 1) {
    r.v[0] = *p;
  } else {
    const uint2 t = *reinterpret_cast<const uint2 *>(p);
    r.v[0] = t.x;
    r.v[CPL - 1] = t.y;
  }
  return r;
}
template <int CPL>
__device__ __forceinline__ void cv_st(uint32_t *p, const CVec<CPL> &x) {
  if (CPL == 1) *p = x.v[0];
  else *reinterpret_cast<uint2 *>(p) = make_uint2(x.v[0], x.v[CPL - 1]);
}
template <int ARITH, int CPL, bool LAZYOUT = false>
__device__ __forceinline__ void cv_ct(CVec<CPL> &X, CVec<CPL> &Y, uint2 tw, const ModQ &m) {
#pragma unroll
  for (int c = 0; c < CPL; c++) ct_bfly<ARITH, LAZYOUT>(X.v[c], Y.v[c], tw.x, tw.y, m);
}
/* CANON: the results may stay in [0, 2q) when the next stage (register bit `bit` - 1) multiplies both */
template <int ARITH, int CPL>
__device__ __forceinline__ void cv_ct_at(CVec<CPL> &X, CVec<CPL> &Y, uint2 tw, const ModQ &m, int k, int bit) {
  if (ARITH == ARITH_CANON && bit > 0 && ((k >> (bit - 1)) & 1)) cv_ct<ARITH, CPL, true>(X, Y, tw, m);
  else cv_ct<ARITH, CPL>(X, Y, tw, m);
}
template <int ARITH, int CPL>
__device__ __forceinline__ void cv_gs(CVec<CPL> &X, CVec<CPL> &Y, uint2 tw, const ModQ &m, uint32_t yb) {
#pragma unroll
  for (int c = 0; c < CPL; c++) gs_bfly<ARITH>(X.v[c], Y.v[c], tw.x, tw.y, m, yb);
}

/* final reduction of a value produced by the forward column+row passes to [0, q) */
template <int ARITH>
__device__ __forceinline__ uint32_t canon_fwd(uint32_t x, uint2 one, const ModQ &m) {
  if (ARITH == ARITH_LAZY) return csub(shoup_mul(x, one.x, one.y, m), m.q);
  if (ARITH == ARITH_HARVEY) return csub(csub(x, m.q2), m.q);
  return x;
}

/* =====================================================================================
 * Column pass, forward: stages 0 .. K1-1 of the CT std->rev dataflow.
 * grid.x = batch * nops * (n2 / 32); block = WARPS * 32 threads.
 * Output range: LAZY < (2 K1 + 1) q, HARVEY [0,4q), CANON [0,q) -- the row pass continues
 * in the same class, so nothing is reduced here.
 * ===================================================================================== */
template <int K1, int ARITH, int CPL>
__global__ void __launch_bounds__(ColGeom<K1>::WARPS * 32, ColGeom<K1, CPL>::MINB)
large_cols_fwd_kernel(const __grid_constant__ LargeParams P) {
  using G = ColGeom<K1, CPL>;
  using V = CVec<CPL>;
  extern __shared__ __align__(16) uint32_t smem[];
  const int lane = threadIdx.x & 31;
  const int w = threadIdx.x >> 5;
  constexpr int lr = LARGE_LR;                            /* log2 of the row length */
  constexpr int lt = lr - G::LOG_TILE;                    /* log2 of the tiles per row */
  const unsigned long long unit = blockIdx.x;
  const uint32_t tile = (uint32_t)(unit & ((1u << lt) - 1));
  const unsigned long long po = unit >> lt;
  const uint32_t op = (P.nops == 2) ? (uint32_t)(po & 1) : 0u;
  const unsigned long long poly = (P.nops == 2) ? (po >> 1) : po;
  const size_t base = ((size_t)poly << (K1 + lr)) + tile * G::TILE_COLS + lane * CPL;
  const uint32_t *src = P.src[op] + base;
  uint32_t *dst = P.dst[op] + base;
  uint32_t *sm = smem + lane * CPL;                       /* [row][32 lanes][CPL] */
  const ModQ m = modq_regs(P.m, P.zero);
  asm volatile("griddepcontrol.launch_dependents;");
  const bool early = (P.pdl & 1u) != 0;
  if (!early) asm volatile("griddepcontrol.wait;" ::: "memory");

  V x[G::NV];
  /* phase A: row bits K1-1 .. K1-RA are register bits; this warp's fixed low row bits = w */
#pragma unroll
  for (int k = 0; k < G::NV; k++) x[k] = cv_ldg<CPL>(src + ((size_t)((k << G::RB) | w) << lr));
#pragma unroll
  for (int s = 0; s < G::RA; s++) {
    const int bit = G::RA - 1 - s;
#pragma unroll
    for (int k = 0; k < G::NV; k++) {
      if (k & (1 << bit)) continue;
      const uint2 tw = P.utw[(1 << s) + (k >> (bit + 1))];      /* constant bank: the index is a compile-time number */
      cv_ct_at<ARITH, CPL>(x[k], x[k | (1 << bit)], tw, m, k, bit);
    }
  }
  if (G::RB == 0) {
    if (early) asm volatile("griddepcontrol.wait;" ::: "memory");
#pragma unroll
    for (int k = 0; k < G::NV; k++) cv_st<CPL>(dst + ((size_t)k << lr), x[k]);
    return;
  }
#pragma unroll
  for (int k = 0; k < G::NV; k++) cv_st<CPL>(sm + ((k << G::RB) | w) * G::TILE_COLS, x[k]);
  __syncthreads();
  /* phase B: row bits RB-1 .. 0 are register bits; fixed high row bits hfix = w*GB + g */
#pragma unroll
  for (int g = 0; g < G::GB; g++) {
    const int hfix = w * G::GB + g;
#pragma unroll
    for (int kk = 0; kk < (1 << G::RB); kk++)
      x[(g << G::RB) + kk] = cv_ld<CPL>(sm + ((hfix << G::RB) | kk) * G::TILE_COLS);
  }
#pragma unroll
  for (int s = 0; s < G::RB; s++) {
    const int bit = G::RB - 1 - s;
#pragma unroll
    for (int g = 0; g < G::GB; g++) {
      const int hfix = w * G::GB + g;
#pragma unroll
      for (int kk = 0; kk < (1 << G::RB); kk++) {
        if (kk & (1 << bit)) continue;
        const uint2 tw = __ldg(P.tab + (1 << (G::RA + s)) + (hfix << s) + (kk >> (bit + 1)));
        cv_ct_at<ARITH, CPL>(x[(g << G::RB) + kk], x[(g << G::RB) + (kk | (1 << bit))], tw, m, kk, bit);
      }
    }
  }
  if (early) asm volatile("griddepcontrol.wait;" ::: "memory");
#pragma unroll
  for (int g = 0; g < G::GB; g++) {
    const int hfix = w * G::GB + g;
#pragma unroll
    for (int kk = 0; kk < (1 << G::RB); kk++)
      cv_st<CPL>(dst + ((size_t)((hfix << G::RB) | kk) << lr), x[(g << G::RB) + kk]);
  }
}

/* =====================================================================================
 * Column pass, inverse: the last K1 stages of the GS rev->std dataflow (row bit beta has
 * t = 2^(K1-1-beta) blocks, block j = row >> (beta+1), twiddle p[t + j]).  Inputs must be
 * < 2q (the row pass canonicalises); the very last stage multiplies its sum branch by
 * last_x and its diff branch by last_y, outputs canonical [0, q).
 * ===================================================================================== */
template <int ARITH>
__device__ __forceinline__ void gs_last(uint32_t &X, uint32_t &Y, uint32_t yb, uint2 lx, uint2 ly,
                                        const ModQ &m) {
  uint32_t s, d;
  if (ARITH == ARITH_LAZY) { d = X - Y + yb; s = X + Y; }
  else if (ARITH == ARITH_HARVEY) { d = X - Y + m.q2; s = X + Y; }
  else { s = X + Y; d = X - Y + m.q; }
  Y = csub(shoup_mul(d, ly.x, ly.y, m), m.q);
  X = csub(shoup_mul(s, lx.x, lx.y, m), m.q);
}

template <int K1, int ARITH, int CPL>
__global__ void __launch_bounds__(ColGeom<K1>::WARPS * 32, ColGeom<K1, CPL>::MINB)
large_cols_inv_kernel(const __grid_constant__ LargeParams P) {
  using G = ColGeom<K1, CPL>;
  using V = CVec<CPL>;
  extern __shared__ __align__(16) uint32_t smem[];
  const int lane = threadIdx.x & 31;
  const int w = threadIdx.x >> 5;
  constexpr int lr = LARGE_LR;
  constexpr int lt = lr - G::LOG_TILE;
  const unsigned long long unit = blockIdx.x;
  const uint32_t tile = (uint32_t)(unit & ((1u << lt) - 1));
  const unsigned long long poly = unit >> lt;
  const size_t base = ((size_t)poly << (K1 + lr)) + tile * G::TILE_COLS + lane * CPL;
  const uint32_t *src = P.src[0] + base;
  uint32_t *dst = P.dst[0] + base;
  uint32_t *sm = smem + lane * CPL;
  const ModQ m = modq_regs(P.m, P.zero);
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");

  V x[G::NV];
  if (G::RB > 0) {
#pragma unroll
    for (int g = 0; g < G::GB; g++) {
      const int hfix = w * G::GB + g;
#pragma unroll
      for (int kk = 0; kk < (1 << G::RB); kk++)
        x[(g << G::RB) + kk] = cv_ldg<CPL>(src + ((size_t)((hfix << G::RB) | kk) << lr));
    }
#pragma unroll
    for (int bit = 0; bit < G::RB; bit++) {
      const uint32_t yb = m.q2 << bit;
#pragma unroll
      for (int g = 0; g < G::GB; g++) {
        const int hfix = w * G::GB + g;
#pragma unroll
        for (int kk = 0; kk < (1 << G::RB); kk++) {
          if (kk & (1 << bit)) continue;
          const int j = (hfix << (G::RB - 1 - bit)) | (kk >> (bit + 1));
          const uint2 tw = __ldg(P.tab_inv + (1 << (K1 - 1 - bit)) + j);
          cv_gs<ARITH, CPL>(x[(g << G::RB) + kk], x[(g << G::RB) + (kk | (1 << bit))], tw, m, yb);
        }
      }
    }
#pragma unroll
    for (int g = 0; g < G::GB; g++) {
      const int hfix = w * G::GB + g;
#pragma unroll
      for (int kk = 0; kk < (1 << G::RB); kk++)
        cv_st<CPL>(sm + ((hfix << G::RB) | kk) * G::TILE_COLS, x[(g << G::RB) + kk]);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < G::NV; k++) x[k] = cv_ld<CPL>(sm + ((k << G::RB) | w) * G::TILE_COLS);
  } else {
#pragma unroll
    for (int k = 0; k < G::NV; k++) x[k] = cv_ldg<CPL>(src + ((size_t)k << lr));
  }
#pragma unroll
  for (int bit = 0; bit < G::RA; bit++) {
    const uint32_t yb = m.q2 << (G::RB + bit);
#pragma unroll
    for (int k = 0; k < G::NV; k++) {
      if (k & (1 << bit)) continue;
      if (bit < G::RA - 1) {
        const uint2 tw = P.utw[(1 << (G::RA - 1 - bit)) + (k >> (bit + 1))];
        cv_gs<ARITH, CPL>(x[k], x[k | (1 << bit)], tw, m, yb);
      } else {
#pragma unroll
        for (int c = 0; c < CPL; c++)
          gs_last<ARITH>(x[k].v[c], x[k | (1 << bit)].v[c], yb, P.last_x, P.last_y, m);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < G::NV; k++) cv_st<CPL>(dst + ((size_t)((k << G::RB) | w) << lr), x[k]);
}

/* =====================================================================================
 * Row pass helpers: twiddles of row j at local level l are p[2^(k1+l) + (j << l) + ...]
 * ===================================================================================== */
template <int LR>
__device__ __forceinline__ void load_row_uniform_tw(uint2 (&tw)[1 << SmallGeom<LR>::R], const uint2 *tab,
                                                    uint32_t k1, uint32_t j) {
  using Gm = SmallGeom<LR>;
  tw[0] = make_uint2(0, 0);
#pragma unroll
  for (int s = 0; s < Gm::R; s++)
#pragma unroll
    for (int u = 0; u < (1 << s); u++) tw[(1 << s) + u] = __ldg(tab + (1u << (k1 + s)) + (j << s) + u);
}

template <int LR>
__device__ __forceinline__ void load_row_lane_tw(LaneTw<LR> &t, const uint2 *tab, int l, uint32_t k1,
                                                 uint32_t j) {
  using Gm = SmallGeom<LR>;
  constexpr int PER_ROW = LaneTw<LR>::PER_ROW;
#pragma unroll
  for (int g = 0; g < (1 << Gm::G); g++) {
    const uint32_t row = (j << Gm::R) | (uint32_t)((g << Gm::H) | l);
#pragma unroll
    for (int mm = 0; mm < Gm::H; mm++) {
      const uint2 *src = tab + (1u << (k1 + Gm::R + mm)) + (row << mm);
      if (mm == 0) {
        t.w[g * PER_ROW + 0] = __ldg(src);
      } else {
#pragma unroll
        for (int u = 0; u < (1 << mm); u += 2) {
          uint4 v = __ldg(reinterpret_cast<const uint4 *>(src + u));
          t.w[g * PER_ROW + ((1 << mm) - 1) + u] = make_uint2(v.x, v.y);
          t.w[g * PER_ROW + ((1 << mm) - 1) + u + 1] = make_uint2(v.z, v.w);
        }
      }
    }
  }
}

/* =====================================================================================
 * K2: row pass of the product.  grid.x = n1 * ceil(batch / (WARPS * PPW)); block (j, bg)
 * handles row j of WARPS*PPW consecutive polynomials, so the 8 warps of a CTA fetch the
 * same 4 KiB of row twiddles (one L2 read, then L1 hits).
 * ===================================================================================== */
#ifndef LARGE_ROW_MINB
#define LARGE_ROW_MINB 3
#endif
template <int LR, int ARITH, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, LARGE_ROW_MINB)
large_rows_polymul_kernel(const __grid_constant__ LargeParams P) {
  using Gm = SmallGeom<LR>;
  extern __shared__ __align__(16) uint32_t smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int sub = lane / Gm::T;
  const int l = lane % Gm::T;
  uint32_t *sm_a = smem + (warp * 2 * Gm::PPW + sub) * Gm::STRIDE;
  uint32_t *sm_b = sm_a + Gm::PPW * Gm::STRIDE;
  const ModQ m = modq_regs(P.m, P.zero);
  const uint32_t k1 = P.logn - LR;
  const uint32_t j = blockIdx.x & ((1u << k1) - 1);
  const unsigned long long bg = blockIdx.x >> k1;
  const unsigned long long poly = (bg * WARPS + warp) * Gm::PPW + sub;
  const bool live = poly < P.batch;
  const size_t off = ((size_t)(live ? poly : 0ull) << P.logn) + ((size_t)j << LR);
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");

  uint32_t xa[Gm::NV], xb[Gm::NV];
  gload_cols<LR>(xa, P.src[0] + off, l);
  gload_cols<LR>(xb, P.src[1] + off, l);
  {
    uint2 twu[1 << Gm::R];
    load_row_uniform_tw<LR>(twu, P.tab, k1, j);
    fwd_phase_cols<LR, ARITH>(xa, twu, m);
    fwd_phase_cols<LR, ARITH>(xb, twu, m);
  }
  store_cols<LR>(xa, sm_a, l);
  store_cols<LR>(xb, sm_b, l);
  __syncwarp();
  load_rows<LR>(xa, sm_a, l);
  load_rows<LR>(xb, sm_b, l);
  {
    LaneTw<LR> twl;
    load_row_lane_tw<LR>(twl, P.tab, l, k1, j);
    fwd_phase_rows<LR, ARITH, true>(xa, twl, m);
    fwd_phase_rows<LR, ARITH>(xb, twl, m);
  }
#pragma unroll
  for (int k = 0; k < Gm::NV; k++) {
    uint32_t av = xa[k], bv = xb[k];
    if (ARITH == ARITH_HARVEY) { av = csub(av, m.q2); bv = csub(bv, m.q2); }
    uint32_t v = mont_mul(av, bv, m);                 /* (0, 2q); the 2^-32 is cancelled in K3 */
    xa[k] = (ARITH == ARITH_CANON) ? csub(v, m.q) : v;
  }
  {
    LaneTw<LR> twl;
    load_row_lane_tw<LR>(twl, P.tab_inv, l, k1, j);
    inv_phase_rows<LR, ARITH>(xa, twl, m);
  }
  __syncwarp();
  store_rows<LR>(xa, sm_a, l);
  __syncwarp();
  load_cols<LR>(xa, sm_a, l);
  {
    uint2 twu[1 << Gm::R];
    load_row_uniform_tw<LR>(twu, P.tab_inv, k1, j);
    inv_phase_cols<LR, ARITH, true>(xa, twu, m, P.one, twu[1]);
  }
  canon_2q<Gm::NV, ARITH>(xa, m);
  if (live) gstore_cols<LR>(xa, P.dst[0] + off, l);
}

/* =====================================================================================
 * Row pass of a standalone transform, in place on dst[0] (src[0] == dst[0]).
 *   DIR 0: forward rows after the forward column pass, canonical output
 *   DIR 1: inverse rows before the inverse column pass, canonical output
 * ===================================================================================== */
template <int LR, int ARITH, int WARPS, int DIR>
__global__ void __launch_bounds__(WARPS * 32)
large_rows_ntt_kernel(const __grid_constant__ LargeParams P) {
  using Gm = SmallGeom<LR>;
  extern __shared__ __align__(16) uint32_t smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int sub = lane / Gm::T;
  const int l = lane % Gm::T;
  uint32_t *sm_a = smem + (warp * Gm::PPW + sub) * Gm::STRIDE;
  const ModQ m = modq_regs(P.m, P.zero);
  const uint32_t k1 = P.logn - LR;
  const uint32_t j = blockIdx.x & ((1u << k1) - 1);
  const unsigned long long bg = blockIdx.x >> k1;
  const unsigned long long poly = (bg * WARPS + warp) * Gm::PPW + sub;
  const bool live = poly < P.batch;
  const size_t off = ((size_t)(live ? poly : 0ull) << P.logn) + ((size_t)j << LR);
  uint32_t x[Gm::NV];
  if (DIR == 0) {
    gload_cols<LR>(x, P.src[0] + off, l);
    {
      uint2 twu[1 << Gm::R];
      load_row_uniform_tw<LR>(twu, P.tab, k1, j);
      fwd_phase_cols<LR, ARITH>(x, twu, m);
    }
    store_cols<LR>(x, sm_a, l);
    __syncwarp();
    load_rows<LR>(x, sm_a, l);
    {
      LaneTw<LR> twl;
      load_row_lane_tw<LR>(twl, P.tab, l, k1, j);
      fwd_phase_rows<LR, ARITH>(x, twl, m);
    }
#pragma unroll
    for (int k = 0; k < Gm::NV; k++) x[k] = canon_fwd<ARITH>(x[k], P.one, m);
    if (live) gstore_rows<LR>(x, P.dst[0] + off, l);
  } else {
    gload_rows<LR>(x, P.src[0] + off, l);
    {
      LaneTw<LR> twl;
      load_row_lane_tw<LR>(twl, P.tab_inv, l, k1, j);
      inv_phase_rows<LR, ARITH>(x, twl, m);
    }
    store_rows<LR>(x, sm_a, l);
    __syncwarp();
    load_cols<LR>(x, sm_a, l);
    {
      uint2 twu[1 << Gm::R];
      load_row_uniform_tw<LR>(twu, P.tab_inv, k1, j);
      inv_phase_cols<LR, ARITH, true>(x, twu, m, P.one, twu[1]);
    }
    canon_2q<Gm::NV, ARITH>(x, m);
    if (live) gstore_cols<LR>(x, P.dst[0] + off, l);
  }
}

}  // namespace nttb200
