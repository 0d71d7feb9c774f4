/*
 * ntt_small.cuh -- register/shared-memory resident kernels for n = 2^L, 3 <= L <= 10.
 *
 * One polynomial is handled by a group of T = 2^H lanes of one warp (H = L - R), each lane
 * holding NV = 2^R coefficients in registers, R = ceil(L/2).  A transform runs in two
 * register phases separated by ONE transposition through shared memory:
 *
 *   layout 1 (global-memory side, "column" layout)
 *       register k  <->  index bits [L-1 .. H]      lane l  <->  index bits [H-1 .. 0]
 *       -> the butterflies on the top R index bits are register-local and their twiddles
 *          p[t+j] depend on register indices only: they are kernel-parameter (constant bank)
 *          operands, identical for every lane.
 *   layout 2 ("row" layout)
 *       register r = (g, r_lo)  <->  index bits [L-1 .. 2H] (G = R-H of them) and [H-1 .. 0]
 *       lane l'                 <->  index bits [2H-1 .. H]
 *       -> the butterflies on the low H index bits are register-local; their twiddles depend
 *          on the lane (p[2^(R+m) + (row << m) + u], a contiguous run per lane and level).
 *
 * Forward  = Cooley-Tukey, standard order in, bit-reversed order out, psi folded into the
 *            twiddles (dataflow of mulntt_ct_std2rev, R/NTT/ntt.C:342-371): layout 1 then 2.
 * Inverse  = Gentleman-Sande, bit-reversed in, standard out, psi^-1 folded in (dataflow of
 *            nttmul_gs_rev2std, R/NTT/ntt.C:428-451): layout 2 then 1, n^-1 folded into the
 *            last stage.
 * The fused product kernel chains  fwd(a), fwd(b) -> pointwise -> inv  without leaving the
 * SM; natural -> bit-reversed -> natural ordering means no permutation pass ever runs
 * (the reference's products never call bitrev_shuffle either, R/NTT/ntt256.C:5-24).
 */
#pragma once
#include <stdint.h>
#include "modarith.cuh"

namespace nttb200 {

template <int R>
struct UniformTw {
  /* entries [1, 2^R) of the level table p[t+j] (t < 2^R), with Shoup companions.
   * .x = w, .y = floor(w 2^32 / q).  Entry 0 unused. */
  uint2 fwd[1 << R];
  uint2 inv[1 << R];
};

template <int R>
struct SmallParams {
  const uint32_t *a;       /* [batch][n] (fused: operand a; transform: in/out) */
  const uint32_t *b;       /* [batch][n] */
  uint32_t *c;             /* [batch][n] */
  const uint2 *tw_fwd;     /* device level table, n entries of (w, w')          */
  const uint2 *tw_inv;     /* device level table for the inverse                */
  unsigned long long batch;
  ModQ m;
  uint2 last_x;            /* multiplier of the SUM branch of the last inverse stage:
                              n^-1 * 2^32 (fused) / n^-1 (scaled intt) / unused          */
  uint2 last_y;            /* multiplier of the DIFF branch of the last inverse stage    */
  uint32_t flags;
  uint32_t zero;           /* always 0 (modq_regs) */
  UniformTw<R> u;
};

enum {
  SMALL_FLAG_SCALE_LAST = 1u,   /* multiply the sum branch of the last inverse stage by last_x */
  SMALL_FLAG_NOWAIT = 2u,       /* the launch shares no data with the launches still in flight on its stream
                                   (nttb200_launch_independent): it waits for them at its end, not its start */
};
/* programmatic dependent launch: let the next launch of the stream start its CTAs while this grid
 * drains; a launch whose operands may come from its predecessors waits for them before touching
 * memory, an independent one only before it ends (so the stream still completes in order).
 * A launch that waits triggers its dependents only AFTER its wait: the host checks a new launch against
 * the launches since the last one that waited (nttb200_launch_independent), which is only enough if
 * nothing older can still be running when the new launch starts -- a waiting launch that triggered
 * first would let its successor run next to ITS predecessors, unchecked (found by
 * test_random_launch_sequences_on_one_stream_equal_their_sequential_meaning). */
__device__ __forceinline__ void pdl_begin(uint32_t flags) {
  if (!(flags & SMALL_FLAG_NOWAIT)) asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;");
}
__device__ __forceinline__ void pdl_end(uint32_t flags) {
  if (flags & SMALL_FLAG_NOWAIT) asm volatile("griddepcontrol.wait;" ::: "memory");
}

template <int L>
struct SmallGeom {
  static constexpr int R = (L + 1) / 2;
  static constexpr int H = L - R;
  static constexpr int G = R - H;                 /* 0 or 1 */
  static constexpr int N = 1 << L;
  static constexpr int NV = 1 << R;               /* coefficients per lane */
  static constexpr int T = 1 << H;                /* lanes per polynomial  */
  static constexpr int PPW = 32 / T;              /* polynomials per warp  */
  static constexpr int STRIDE = N + (T < 32 ? T : 0);   /* smem words per polynomial */
  static constexpr int CPR = (H >= 2) ? (1 << (H - 2)) : 1;   /* 16-byte chunks per row */
};

/* xor-swizzle of the 16-byte chunk index inside a row so that both the scalar column
 * accesses and the 128-bit row accesses are bank-conflict free */
template <int H>
__device__ __forceinline__ int swz(int row) {
  if (H >= 5) return row & 7;
  if (H == 4) return (row >> 1) & 3;
  if (H == 3) return (row >> 2) & 1;
  return 0;
}

/* ---- transpositions (sm points at this polynomial's STRIDE-word region) ------------- */

template <int L>
__device__ __forceinline__ void store_cols(const uint32_t (&x)[SmallGeom<L>::NV], uint32_t *sm, int l) {
  using Gm = SmallGeom<L>;
#pragma unroll
  for (int k = 0; k < Gm::NV; k++) {
    int chunk = (l >> 2) ^ swz<Gm::H>(k);
    sm[(k << Gm::H) + (chunk << 2) + (l & 3)] = x[k];
  }
}
template <int L>
__device__ __forceinline__ void load_cols(uint32_t (&x)[SmallGeom<L>::NV], const uint32_t *sm, int l) {
  using Gm = SmallGeom<L>;
#pragma unroll
  for (int k = 0; k < Gm::NV; k++) {
    int chunk = (l >> 2) ^ swz<Gm::H>(k);
    x[k] = sm[(k << Gm::H) + (chunk << 2) + (l & 3)];
  }
}
template <int L>
__device__ __forceinline__ void load_rows(uint32_t (&x)[SmallGeom<L>::NV], const uint32_t *sm, int l) {
  using Gm = SmallGeom<L>;
#pragma unroll
  for (int g = 0; g < (1 << Gm::G); g++) {
    int row = (g << Gm::H) | l;
    if (Gm::H >= 2) {
#pragma unroll
      for (int c = 0; c < Gm::CPR; c++) {
        uint4 v = *reinterpret_cast<const uint4 *>(sm + (row << Gm::H) + ((c ^ swz<Gm::H>(row)) << 2));
        x[(g << Gm::H) + 4 * c + 0] = v.x;
        x[(g << Gm::H) + 4 * c + 1] = v.y;
        x[(g << Gm::H) + 4 * c + 2] = v.z;
        x[(g << Gm::H) + 4 * c + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int r = 0; r < Gm::T; r++) x[(g << Gm::H) + r] = sm[(row << Gm::H) + r];
    }
  }
}
template <int L>
__device__ __forceinline__ void store_rows(const uint32_t (&x)[SmallGeom<L>::NV], uint32_t *sm, int l) {
  using Gm = SmallGeom<L>;
#pragma unroll
  for (int g = 0; g < (1 << Gm::G); g++) {
    int row = (g << Gm::H) | l;
    if (Gm::H >= 2) {
#pragma unroll
      for (int c = 0; c < Gm::CPR; c++) {
        uint4 v;
        v.x = x[(g << Gm::H) + 4 * c + 0];
        v.y = x[(g << Gm::H) + 4 * c + 1];
        v.z = x[(g << Gm::H) + 4 * c + 2];
        v.w = x[(g << Gm::H) + 4 * c + 3];
        *reinterpret_cast<uint4 *>(sm + (row << Gm::H) + ((c ^ swz<Gm::H>(row)) << 2)) = v;
      }
    } else {
#pragma unroll
      for (int r = 0; r < Gm::T; r++) sm[(row << Gm::H) + r] = x[(g << Gm::H) + r];
    }
  }
}

/* ---- global memory, layout 1: lane l, register k  <->  coefficient (k << H) | l ------
 * Operands are read with ld.global.cg (L2), NOT through the non-coherent path (__ldg): the kernels
 * are launched with programmatic stream serialization, so their lifetime overlaps the previous
 * kernel of the stream -- which may be the one that WRITES these operands.  After
 * griddepcontrol.wait its stores are visible in L2, but a line it read earlier can still sit in this
 * SM's non-coherent cache; "read-only for the lifetime of the kernel" does not hold.  (Found by
 * tests/test_gpu_parity.py::test_random_launch_sequences_...; twiddle tables stay on __ldg.) */
template <int L>
__device__ __forceinline__ void gload_cols(uint32_t (&x)[SmallGeom<L>::NV], const uint32_t *g, int l) {
  using Gm = SmallGeom<L>;
#pragma unroll
  for (int k = 0; k < Gm::NV; k++) x[k] = __ldcg(g + (k << Gm::H) + l);
}
template <int L>
__device__ __forceinline__ void gstore_cols(const uint32_t (&x)[SmallGeom<L>::NV], uint32_t *g, int l) {
  using Gm = SmallGeom<L>;
#pragma unroll
  for (int k = 0; k < Gm::NV; k++) g[(k << Gm::H) + l] = x[k];
}
/* global memory, layout 2: lane l' owns rows (g << H) | l', 2^H contiguous coefficients each */
template <int L>
__device__ __forceinline__ void gload_rows(uint32_t (&x)[SmallGeom<L>::NV], const uint32_t *gp, int l) {
  using Gm = SmallGeom<L>;
#pragma unroll
  for (int g = 0; g < (1 << Gm::G); g++) {
    const uint32_t *rowp = gp + ((((g << Gm::H) | l)) << Gm::H);
    if (Gm::H >= 2) {
#pragma unroll
      for (int c = 0; c < Gm::CPR; c++) {
        uint4 v = __ldcg(reinterpret_cast<const uint4 *>(rowp) + c);
        x[(g << Gm::H) + 4 * c + 0] = v.x;
        x[(g << Gm::H) + 4 * c + 1] = v.y;
        x[(g << Gm::H) + 4 * c + 2] = v.z;
        x[(g << Gm::H) + 4 * c + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int r = 0; r < Gm::T; r++) x[(g << Gm::H) + r] = __ldcg(rowp + r);
    }
  }
}
template <int L>
__device__ __forceinline__ void gstore_rows(const uint32_t (&x)[SmallGeom<L>::NV], uint32_t *gp, int l) {
  using Gm = SmallGeom<L>;
#pragma unroll
  for (int g = 0; g < (1 << Gm::G); g++) {
    uint32_t *rowp = gp + ((((g << Gm::H) | l)) << Gm::H);
    if (Gm::H >= 2) {
#pragma unroll
      for (int c = 0; c < Gm::CPR; c++) {
        uint4 v;
        v.x = x[(g << Gm::H) + 4 * c + 0];
        v.y = x[(g << Gm::H) + 4 * c + 1];
        v.z = x[(g << Gm::H) + 4 * c + 2];
        v.w = x[(g << Gm::H) + 4 * c + 3];
        reinterpret_cast<uint4 *>(rowp)[c] = v;
      }
    } else {
#pragma unroll
      for (int r = 0; r < Gm::T; r++) rowp[r] = x[(g << Gm::H) + r];
    }
  }
}

/* ---- per-lane twiddles of the layout-2 phase ---------------------------------------
 * level m (0 <= m < H) handles index bit H-1-m; the lane needs, for each of its rows,
 * the 2^m consecutive table entries starting at 2^(R+m) + (row << m).               */
template <int L>
struct LaneTw {
  using Gm = SmallGeom<L>;
  static constexpr int PER_ROW = (1 << Gm::H) - 1;       /* 1 + 2 + ... + 2^(H-1) */
  uint2 w[(1 << Gm::G) * (PER_ROW > 0 ? PER_ROW : 1)];
  __device__ __forceinline__ void load(const uint2 *tab, int l) {
#pragma unroll
    for (int g = 0; g < (1 << Gm::G); g++) {
      int row = (g << Gm::H) | l;
#pragma unroll
      for (int m = 0; m < Gm::H; m++) {
        const uint2 *src = tab + (1 << (Gm::R + m)) + (row << m);
        if (m == 0) {
          w[g * PER_ROW + 0] = __ldg(src);
        } else {
#pragma unroll
          for (int u = 0; u < (1 << m); u += 2) {
            uint4 v = __ldg(reinterpret_cast<const uint4 *>(src + u));
            w[g * PER_ROW + ((1 << m) - 1) + u] = make_uint2(v.x, v.y);
            w[g * PER_ROW + ((1 << m) - 1) + u + 1] = make_uint2(v.z, v.w);
          }
        }
      }
    }
  }
  __device__ __forceinline__ uint2 get(int g, int m, int u) const {
    return w[g * PER_ROW + ((1 << m) - 1) + u];
  }
};

/* ---- register phases ------------------------------------------------------------- */

/* forward, layout 1: index bits L-1 .. H  (register bits R-1 .. 0), uniform twiddles */
template <int L, int ARITH>
__device__ __forceinline__ void fwd_phase_cols(uint32_t (&x)[SmallGeom<L>::NV],
                                               const uint2 (&tw)[1 << SmallGeom<L>::R], const ModQ &m) {
  using Gm = SmallGeom<L>;
#pragma unroll
  for (int s = 0; s < Gm::R; s++) {
    const int kb = Gm::R - 1 - s;                     /* register bit handled by this stage */
#pragma unroll
    for (int k = 0; k < Gm::NV; k++) {
      if (k & (1 << kb)) continue;
      const int j = k >> (kb + 1);
      const uint2 w = tw[(1 << s) + j];
      /* both results go into the next stage's multiplication when the next register bit is set */
      if (ARITH == ARITH_CANON && kb > 0 && ((k >> (kb - 1)) & 1))
        ct_bfly<ARITH, true>(x[k], x[k | (1 << kb)], w.x, w.y, m);
      else
        ct_bfly<ARITH>(x[k], x[k | (1 << kb)], w.x, w.y, m);
    }
  }
}

/* forward, layout 2: index bits H-1 .. 0 (register bits H-1 .. 0 of r_lo), lane twiddles.
 * LASTLAZY (CANON): the results of the very last stage stay in [0, 2q) -- for the ONE operand of
 * the pointwise Montgomery product that may (a b < 2q q < q 2^32). */
template <int L, int ARITH, bool LASTLAZY = false>
__device__ __forceinline__ void fwd_phase_rows(uint32_t (&x)[SmallGeom<L>::NV], const LaneTw<L> &tw,
                                               const ModQ &m) {
  using Gm = SmallGeom<L>;
#pragma unroll
  for (int lv = 0; lv < Gm::H; lv++) {
    const int bit = Gm::H - 1 - lv;
#pragma unroll
    for (int r = 0; r < Gm::NV; r++) {
      if (r & (1 << bit)) continue;
      const int g = r >> Gm::H;
      const int u = (r & (Gm::T - 1)) >> (bit + 1);
      const uint2 w = tw.get(g, lv, u);
      if (ARITH == ARITH_CANON && ((bit > 0 && ((r >> (bit - 1)) & 1)) || (bit == 0 && LASTLAZY)))
        ct_bfly<ARITH, true>(x[r], x[r | (1 << bit)], w.x, w.y, m);
      else
        ct_bfly<ARITH>(x[r], x[r | (1 << bit)], w.x, w.y, m);
    }
  }
}

/* inverse, layout 2: index bits 0 .. H-1, lane twiddles.  `stage0` = number of GS stages
 * already applied (for the LAZY bound 2q * 2^stage). */
template <int L, int ARITH>
__device__ __forceinline__ void inv_phase_rows(uint32_t (&x)[SmallGeom<L>::NV], const LaneTw<L> &tw,
                                               const ModQ &m) {
  using Gm = SmallGeom<L>;
#pragma unroll
  for (int bit = 0; bit < Gm::H; bit++) {
    const int lv = Gm::H - 1 - bit;
    const uint32_t yb = m.q2 << bit;                 /* LAZY: inputs < 2q * 2^bit */
#pragma unroll
    for (int r = 0; r < Gm::NV; r++) {
      if (r & (1 << bit)) continue;
      const int g = r >> Gm::H;
      const int u = (r & (Gm::T - 1)) >> (bit + 1);
      const uint2 w = tw.get(g, lv, u);
      gs_bfly<ARITH>(x[r], x[r | (1 << bit)], w.x, w.y, m, yb);
    }
  }
}

/* inverse, layout 1: index bits H .. L-1 (register bits 0 .. R-1), uniform twiddles.
 * The last stage multiplies its sum branch by last_x when SCALE is set; its diff-branch
 * twiddle is last_y (= p[1], times the scale when SCALE). */
template <int L, int ARITH, bool SCALE>
__device__ __forceinline__ void inv_phase_cols(uint32_t (&x)[SmallGeom<L>::NV],
                                               const uint2 (&tw)[1 << SmallGeom<L>::R], const ModQ &m,
                                               uint2 last_x, uint2 last_y) {
  using Gm = SmallGeom<L>;
#pragma unroll
  for (int kb = 0; kb < Gm::R; kb++) {
    const int t = 1 << (Gm::R - 1 - kb);
    const uint32_t yb = m.q2 << (Gm::H + kb);
#pragma unroll
    for (int k = 0; k < Gm::NV; k++) {
      if (k & (1 << kb)) continue;
      const int j = k >> (kb + 1);
      if (kb < Gm::R - 1) {
        const uint2 w = tw[t + j];
        gs_bfly<ARITH>(x[k], x[k | (1 << kb)], w.x, w.y, m, yb);
      } else {
        uint32_t &X = x[k], &Y = x[k | (1 << kb)];
        if (ARITH == ARITH_LAZY) {
          uint32_t d = X - Y + yb, s = X + Y;
          Y = shoup_mul(d, last_y.x, last_y.y, m);
          X = SCALE ? shoup_mul(s, last_x.x, last_x.y, m) : s;
        } else if (ARITH == ARITH_HARVEY) {
          uint32_t d = X - Y + m.q2, s = X + Y;         /* s < 4q */
          Y = shoup_mul(d, last_y.x, last_y.y, m);
          X = SCALE ? shoup_mul(s, last_x.x, last_x.y, m) : csub(s, m.q2);
        } else {
          uint32_t s = X + Y, d = X - Y + m.q;
          Y = shoup_mul(d, last_y.x, last_y.y, m);
          X = SCALE ? shoup_mul(s, last_x.x, last_x.y, m) : csub(s, m.q);
        }
      }
    }
  }
}

/* bring every register to canonical [0, q).  `bound_2q`: values are known < 2q */
template <int NV, int ARITH>
__device__ __forceinline__ void canon_2q(uint32_t (&x)[NV], const ModQ &m) {
#pragma unroll
  for (int k = 0; k < NV; k++) x[k] = csub(x[k], m.q);
}

/* generic reduction of a lazily grown value (any 32-bit x) to [0,q): Shoup-multiply by 1 */
__device__ __forceinline__ uint32_t reduce_any(uint32_t x, uint32_t one_p, const ModQ &m) {
  return csub(shoup_mul(x, 1u, one_p, m), m.q);
}

/* =====================================================================================
 * Fused product kernel:  c = INTT( NTT(a) o NTT(b) ) * n^-1
 * grid-stride over "warp tiles" of PPW polynomials; one smem region per warp.
 * ===================================================================================== */
template <int L, int ARITH, int WARPS, int MINB, bool TWREG>
__global__ void __launch_bounds__(WARPS * 32, MINB)
polymul_small_kernel(const __grid_constant__ SmallParams<SmallGeom<L>::R> P) {
  using Gm = SmallGeom<L>;
  extern __shared__ __align__(16) uint32_t smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int sub = lane / Gm::T;            /* polynomial within the warp tile */
  const int l = lane % Gm::T;              /* lane within the polynomial      */
  uint32_t *sm_a = smem + (warp * 2 * Gm::PPW + sub) * Gm::STRIDE;
  uint32_t *sm_b = sm_a + Gm::PPW * Gm::STRIDE;
  const ModQ m = modq_regs(P.m, P.zero);

  LaneTw<L> twf, twi;
  if (TWREG) {
    twf.load(P.tw_fwd, l);
    twi.load(P.tw_inv, l);
  }
  pdl_begin(P.flags);

  const unsigned long long ntiles = (P.batch + Gm::PPW - 1) / Gm::PPW;
  const unsigned long long wstride = (unsigned long long)gridDim.x * WARPS;
  for (unsigned long long tile = (unsigned long long)blockIdx.x * WARPS + warp; tile < ntiles;
       tile += wstride) {
    const unsigned long long poly = tile * Gm::PPW + sub;
    const bool live = poly < P.batch;
    const unsigned long long off = (live ? poly : 0ull) << L;

    uint32_t xa[Gm::NV], xb[Gm::NV];
    gload_cols<L>(xa, P.a + off, l);
    gload_cols<L>(xb, P.b + off, l);

    fwd_phase_cols<L, ARITH>(xa, P.u.fwd, m);
    fwd_phase_cols<L, ARITH>(xb, P.u.fwd, m);
    if (Gm::H > 0) {
      store_cols<L>(xa, sm_a, l);
      store_cols<L>(xb, sm_b, l);
      __syncwarp();
      load_rows<L>(xa, sm_a, l);
      load_rows<L>(xb, sm_b, l);
      if (!TWREG) twf.load(P.tw_fwd, l);
      fwd_phase_rows<L, ARITH, true>(xa, twf, m);
      fwd_phase_rows<L, ARITH>(xb, twf, m);
    }

    /* pointwise product in the bit-reversed domain (mul_array, R/NTT/ntt.C:131-137);
     * Montgomery REDC leaves a factor 2^-32 that last_x/last_y cancel */
#pragma unroll
    for (int k = 0; k < Gm::NV; k++) {
      uint32_t av = xa[k], bv = xb[k];
      if (ARITH == ARITH_HARVEY) { av = csub(av, m.q2); bv = csub(bv, m.q2); }
      uint32_t v = mont_mul(av, bv, m);              /* (0, 2q) */
      xa[k] = (ARITH == ARITH_CANON) ? csub(v, m.q) : v;
    }

    if (Gm::H > 0) {
      if (!TWREG) twi.load(P.tw_inv, l);
      inv_phase_rows<L, ARITH>(xa, twi, m);
      __syncwarp();                                   /* all lanes done reading sm_a */
      store_rows<L>(xa, sm_a, l);
      __syncwarp();
      load_cols<L>(xa, sm_a, l);
    }
    inv_phase_cols<L, ARITH, true>(xa, P.u.inv, m, P.last_x, P.last_y);
    canon_2q<Gm::NV, ARITH>(xa, m);
    if (live) gstore_cols<L>(xa, P.c + off, l);
    __syncwarp();                                     /* smem reuse by the next tile */
  }
  pdl_end(P.flags);
}

/* =====================================================================================
 * Standalone transforms on the same building blocks (in place on P.c).
 *   DIR 0: forward CT std->rev  (table tw_fwd / u.fwd), canonical output
 *   DIR 1: inverse GS rev->std  (table tw_inv / u.inv), canonical output,
 *          SCALE_LAST folds last_x / last_y into the last stage
 * The table decides which reference function this is (plain omega table: ntt_ct_std2rev /
 * ntt_gs_rev2std; mixed table: mulntt_ct_std2rev / nttmul_gs_rev2std).
 * ===================================================================================== */
template <int L, int ARITH, int WARPS, int DIR>
__global__ void __launch_bounds__(WARPS * 32)
ntt_small_kernel(const __grid_constant__ SmallParams<SmallGeom<L>::R> P) {
  using Gm = SmallGeom<L>;
  extern __shared__ __align__(16) uint32_t smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int sub = lane / Gm::T;
  const int l = lane % Gm::T;
  uint32_t *sm_a = smem + (warp * Gm::PPW + sub) * Gm::STRIDE;
  const ModQ m = modq_regs(P.m, P.zero);
  const uint32_t one_p = 0xFFFFFFFFu / m.q;          /* floor(2^32/q) for q not a power of 2 */

  LaneTw<L> tw;
  tw.load(DIR == 0 ? P.tw_fwd : P.tw_inv, l);
  pdl_begin(P.flags);

  const unsigned long long ntiles = (P.batch + Gm::PPW - 1) / Gm::PPW;
  const unsigned long long wstride = (unsigned long long)gridDim.x * WARPS;
  for (unsigned long long tile = (unsigned long long)blockIdx.x * WARPS + warp; tile < ntiles;
       tile += wstride) {
    const unsigned long long poly = tile * Gm::PPW + sub;
    const bool live = poly < P.batch;
    const unsigned long long off = (live ? poly : 0ull) << L;
    uint32_t x[Gm::NV];
    if (DIR == 0) {
      gload_cols<L>(x, P.c + off, l);
      fwd_phase_cols<L, ARITH>(x, P.u.fwd, m);
      if (Gm::H > 0) {
        store_cols<L>(x, sm_a, l);
        __syncwarp();
        load_rows<L>(x, sm_a, l);
        fwd_phase_rows<L, ARITH>(x, tw, m);
      }
#pragma unroll
      for (int k = 0; k < Gm::NV; k++) {
        if (ARITH == ARITH_LAZY) x[k] = reduce_any(x[k], one_p, m);
        else if (ARITH == ARITH_HARVEY) x[k] = csub(csub(x[k], m.q2), m.q);
      }
      __syncwarp();
      if (live) gstore_rows<L>(x, P.c + off, l);
    } else {
      gload_rows<L>(x, P.c + off, l);
      if (Gm::H > 0) {
        inv_phase_rows<L, ARITH>(x, tw, m);
        store_rows<L>(x, sm_a, l);
        __syncwarp();
        load_cols<L>(x, sm_a, l);
      }
      if (P.flags & SMALL_FLAG_SCALE_LAST) {
        inv_phase_cols<L, ARITH, true>(x, P.u.inv, m, P.last_x, P.last_y);
        canon_2q<Gm::NV, ARITH>(x, m);
      } else {
        inv_phase_cols<L, ARITH, false>(x, P.u.inv, m, P.last_x, P.last_y);
#pragma unroll
        for (int k = 0; k < Gm::NV; k++) {
          if (k & (1 << (Gm::R - 1))) x[k] = csub(x[k], m.q);          /* diff branch: < 2q */
          else if (ARITH == ARITH_LAZY) x[k] = reduce_any(x[k], one_p, m);
          else if (ARITH == ARITH_HARVEY) x[k] = csub(x[k], m.q);      /* sum branch: < 2q  */
        }
      }
      __syncwarp();
      if (live) gstore_cols<L>(x, P.c + off, l);
    }
  }
  pdl_end(P.flags);
}

}  // namespace nttb200
