/*
 * ntt_small_splant.cuh -- fused product kernel for half-word moduli (q <= 12385), n = 2^L <= 256,
 * with SIGNED Plantard arithmetic and FIVE-instruction Cooley-Tukey butterflies.
 *
 * Same dataflow, layouts, prefetch and launch machinery as ntt_small_plant.cuh (CT forward std->rev
 * with psi folded in, pointwise, GS inverse rev->std with n^-1 folded into the last stage;
 * R/NTT/ntt.C:342-371, 131-137, 428-451).  What changes is the representation: coefficients are
 * signed 32-bit values, congruent to the reference's canonical ones and bounded at compile time, and
 * a multiplication by a constant w is
 *
 *      p = Y * w~   (mod 2^32, read as a SIGNED word)          w~ = W q^-1 mod 2^32,
 *      u = (p >> 16) * q + D                                   W  = (-w 2^32) mod q, centred
 *      T = u >> 16                                             (arithmetic shifts)
 *
 *   [p q = k 2^32 + Y W with k == Y w (mod q); writing p = h 2^16 + p_lo, 0 <= p_lo < 2^16:
 *    (h q + D) 2^16 = k 2^32 + (Y W - p_lo q + D 2^16), so T = k exactly whenever
 *    0 <= Y W - p_lo q + D 2^16 < 2^32, i.e. for |Y W| <= M with D = ceil((M + 65535 q) / 65536) and
 *    2 M + 65536 (q + 4) <= 2^32.  |p| <= 2^31 makes |k| <= (q - 1)/2: the result is CENTRED.
 *    At q = 12289: M = 1.74e9, |W| <= 6144, so any |Y| <= 23 q may be multiplied.
 *    tests/test_plantard_arith.py pins this on the CPU.]
 *
 * The point of it: T is the upper half of u, and the one instruction `X + (u >> 16)` is
 * LEA.HI.SX32 on sm_100a; the other leg is X - T = 2X - X', one IADD3.  So
 *
 *      CT butterfly = IMAD, SHF.R.S32, IMAD, LEA.HI.SX32, IADD3          (5 instructions)
 *
 * against IMAD, SHF, IMAD, SHF, IADD, IADD3 of the unsigned half-word form: no `+ q` term (signed
 * legs need none) and no stand-alone second shift.  Measured on B200 (profiles/
 * r2_microbench_butterflies.txt, nttb200_measure_int_peak 14 vs 11): 20.7 against 16.2
 * lane-butterflies per clock and SM, +28 %.  The kernel being issue-bound (84 % of the issue slots),
 * instructions are what counts.
 *
 * Value bounds, in units of q/2 and all known at compile time:
 *   inputs [0, q)                    2
 *   every product                    1      (centred)
 *   forward, after s stages          2 + s  <= 12 at n = 1024: never reduced
 *   pointwise  a b q^-1              operand a is first brought to 1 by a 3-instruction Barrett step
 *                                    a - q ((a C + 2^25) >> 26), C = round(2^26 / q); |a b| <= 3.1 q^2
 *   inverse    sums double per stage; a sum whose bound would pass SP_CAP = 16 (8 q) takes the
 *              same Barrett step, so every difference that is multiplied stays <= 32 (16 q)
 *   last stage both legs multiplied (n^-1 folded in), then min(T, T + q) as unsigned: canonical.
 */
#pragma once
#include <stdint.h>
#include "ntt_small_plant.cuh"

namespace nttb200 {

constexpr int SP_CAP = 16;                 /* units of q/2 */
constexpr int SP_RED_SHIFT = 26;

template <int R>
struct SPlantParams {
  const void *a;
  const void *b;
  void *c;
  const uint32_t *tw_fwd;    /* device level table of centred w~, n entries */
  const uint32_t *tw_inv;
  const uint32_t *zeta;      /* n / 2^V entries: -(w^2) 2^32 mod q, centred, w = forward table entry n / 2^V + j */
  unsigned long long batch;
  uint32_t q, qinv;
  uint32_t dd;               /* D: addend of the second product                             */
  uint32_t cbar;             /* C = round(2^26 / q): the Barrett step                        */
  uint32_t last_x, last_y;   /* centred forms of -n^-1 2^32 and -n^-1 2^32 p_inv[1]         */
  uint32_t zero;
  uint32_t nowait;           /* see PlantParams::nowait                                     */
  unsigned long long *sched; /* tail scheduler of the sizes from 2^PLANT_DYN_MINL up, see PlantParams::sched */
  uint32_t static_rounds;
  uint32_t ufwd[1 << R];     /* entries [1, 2^R) of the tables (constant-bank operands)     */
  uint32_t uinv[1 << R];
};

struct SpRegs { int q; uint32_t qinv; int dd; int cbar; };

/* u = (Y w~ >> 16) q + D; the product is T = u >> 16, which the callers fold into their adds */
__device__ __forceinline__ int sp_mul_u(int y, uint32_t wt, const SpRegs &G) {
  const int p = (int)((uint32_t)y * wt);
  return (p >> 16) * G.q + G.dd;
}
/* Barrett step: x - q round(x / q) for |x| <= 16 q; result within q/2 + 20 of zero */
__device__ __forceinline__ int sp_red(int x, const SpRegs &G) {
  const int r = (x * G.cbar + (1 << (SP_RED_SHIFT - 1))) >> SP_RED_SHIFT;
  return x - r * G.q;
}
/* Which pipe takes the second leg.  Written as an addition, Y' = 2X - X' becomes IADD3 and the
 * butterfly is 2 multiplier-pipe + 3 ALU instructions; both pipes issue one warp instruction every
 * two cycles, so the ALU pipe then binds at 6 cycles per warp butterfly.  Written as mad.lo (X, 2, -X')
 * ptxas picks IMAD or LEA per instance and levels the two pipes (SASS of microbench 18: 1 287 IMAD
 * against 1 274 ALU instructions): 5 cycles.  SPLANT_MAD_NUM of every SPLANT_MAD_DEN butterflies use
 * the mad form.  In the kernel the group multiplication is three quarters IMAD, so the butterflies give the
 * multiplier pipe some room: every other butterfly in the mad form measured best (c2: all 1 540, one in two
 * 1 556, one in four 1 548 M polymul/s; c3 1 546 / 1 560 / 1 540; c4 level). */
#ifndef SPLANT_MAD_NUM
#define SPLANT_MAD_NUM 1
#endif
#ifndef SPLANT_MAD_DEN
#define SPLANT_MAD_DEN 2
#endif
__host__ __device__ constexpr bool sp_use_mad(int i) {
  return SPLANT_MAD_NUM > 0 && (i % SPLANT_MAD_DEN) < SPLANT_MAD_NUM;
}
/* CT: X' = X + T, Y' = X - T = 2X - X'.  redx: X takes a Barrett step first, so that both results are within
 * q + 20 of zero (the last forward stage ahead of a group multiplication of eight, see SpDrop) */
__device__ __forceinline__ void sp_ct(uint32_t &X, uint32_t &Y, uint32_t wt, const SpRegs &G, bool mad = false,
                                      bool redx = false) {
  if (redx) X = (uint32_t)sp_red((int)X, G);
  const int u = sp_mul_u((int)Y, wt, G);
  const int xn = (int)X + (u >> 16);
  if (mad) {
    int yn;
    asm("{ .reg .s32 t; neg.s32 t, %2; mad.lo.s32 %0, %1, 2, t; }" : "=r"(yn) : "r"((int)X), "r"(xn));
    Y = (uint32_t)yn;
  } else {
    Y = 2u * X - (uint32_t)xn;
  }
  X = (uint32_t)xn;
}
/* GS: X' = X + Y (Barrett step when its bound would pass the cap), Y' = (X - Y) w.
 * PENDING SHIFTS (SPLANT_PENDING, tuning experiment, off): a product is T = u >> 16, and the butterfly that
 * consumes two products X = uX >> 16, Y = uY >> 16 gets its sum as LEA.HI.SX32 (uY, X) -- so the
 * inverse network keeps every product as u and only the X leg pays its shift: SHF, LEA.HI.SX32, IADD3
 * for the sum and the difference of two products instead of SHF, SHF, IADD, IADD.  Which registers hold
 * a pending product is known at compile time: at the stage on register bit b > first, those whose bit
 * b - 1 is set (the previous stage left its products there); at the first stage all of them (the
 * outputs of the pair multiplication).  PEND_IN: both legs are pending; the product this butterfly
 * leaves in Y is pending again when PEND_OUT.
 * Result: not one instruction fewer (3 173 / 1 324 per tile at n = 1024 / 256 either way).  ptxas does
 * this by itself from the plain form: given Y = u >> 16, S = X + Y, D = X - Y it emits S = LEA.HI.SX32
 * (u, X) and D = LEA (X, -S, 1) = 2X - S, and drops the shift -- in 144 butterflies per n = 1024 tile
 * against the 72 of the hand-written version. */
#ifndef SPLANT_PENDING
#define SPLANT_PENDING 0
#endif
template <bool REDUCE, bool PEND_IN, bool PEND_OUT>
__device__ __forceinline__ void sp_gs(uint32_t &X, uint32_t &Y, uint32_t wt, const SpRegs &G) {
  int d, s;
  if (PEND_IN) {
    const int xs = (int)X >> 16;
    s = xs + ((int)Y >> 16);
    asm("" : "+r"(s));            /* or the front end rewrites d as xs - (Y >> 16) and the shift of Y is back */
    d = 2 * xs - s;
  } else {
    d = (int)X - (int)Y;
    s = (int)X + (int)Y;
  }
  if (REDUCE) s = sp_red(s, G);
  X = (uint32_t)s;
  const int u = sp_mul_u(d, wt, G);
  Y = (uint32_t)(PEND_OUT ? u : (u >> 16));
}
template <bool PEND_IN, bool PEND_OUT>
__device__ __forceinline__ void sp_gs_r(bool reduce, uint32_t &X, uint32_t &Y, uint32_t wt, const SpRegs &G) {
  if (reduce) sp_gs<true, PEND_IN, PEND_OUT>(X, Y, wt, G);      /* `reduce` is a constant after unrolling */
  else sp_gs<false, PEND_IN, PEND_OUT>(X, Y, wt, G);
}

/* bound (units of q/2) of both legs of the GS butterfly on register bit `bit` of a register whose
 * low bits are `k`, in a register phase entered with every register <= b_in: stage s < bit left a
 * product (1) where bit s of k is set, else a sum (doubled; back to 1 when it passed the cap) */
__host__ __device__ constexpr int sp_leg_bound(int k, int bit, int b_in, int first = 0) {
  int b = b_in;
  for (int s = first; s < bit; s++) {
    if ((k >> s) & 1) b = 1;
    else { b = 2 * b; if (b > SP_CAP) b = 1; }
  }
  return b;
}
/* worst bound of any register after a phase of stages first .. bits-1 */
__host__ __device__ constexpr int sp_phase_out(int bits, int b_in, int first = 0) {
  int worst = 1;
  for (int k = 0; k < (1 << bits); k++) {
    int b = sp_leg_bound(k, bits, b_in, first);
    if (b > worst) worst = b;
  }
  return worst;
}
/* the same over the registers that took at least one product (some index bit >= first set) */
__host__ __device__ constexpr int sp_phase_out_mixed(int bits, int b_in, int first = 0) {
  int worst = 1;
  for (int k = 0; k < (1 << bits); k++) {
    if ((k >> first) == 0) continue;
    int b = sp_leg_bound(k, bits, b_in, first);
    if (b > worst) worst = b;
  }
  return worst;
}

/* ---- incomplete transform (SPLANT_INCOMPLETE = number of stages left out, default 3) -------------
 * The last D forward stages, the pointwise product and the first D inverse stages are replaced by one
 * multiplication of polynomials of degree 2^D - 1 per group of 2^D registers: before its last D stages
 * the Cooley-Tukey network holds a_0 + a_1 x + ... in Z_q[x]/(x^(2^D) - w^2) in the group at positions
 * 2^D j .. 2^D j + 2^D - 1, w being the twiddle of the group's next butterflies (table entry
 * n / 2^D + j): the remaining stages would only evaluate it at the 2^D roots.  So
 *        c_k = sum_{i+j=k} a_i b_j  +  w^2 sum_{i+j=k+2^D} a_i b_j
 * is what the inverse network holds after ITS first D stages, up to the factor 2^D those stages would
 * have contributed (the last stage scales by (n / 2^D)^-1 instead of n^-1).  Same ring, same canonical
 * result.  Raw 32-bit products are added BEFORE they are reduced:
 *        a_i <- Barrett step (|.| <= q/2 + 20)                                  3 each
 *        h_k = redc(sum_{i+j=k+2^D} a_i b_j)   (= -(...) 2^-32, centred)        products + 4
 *        c_k = redc(sum_{i+j=k} a_i b_j + h_k Z), Z = -(w^2) 2^32 mod q centred products + 1 + 4
 * D = 1: 23 instructions per pair against 2 x 5 (butterflies of a and b) + 2 x 8 (pointwise) + 6
 * (inverse butterfly) = 32; D = 2: 59 per group of four against 2 x 23 + 20 + 12 = 78 (measured: c2
 * 1 371 -> 1 473 M polymul/s with D = 1); and the lane phase keeps a half / a quarter of the twiddles.
 * D = 3 does not fit THIS WAY: redc(x) = ((x q^-1 >> 16) q + D) >> 16 needs |x| <= M = 11.5 q^2 (header), and
 * with |b| <= (L + 2 - D) q / 2 a sum of 2^D products reaches 2^D (q/2 + 20)(L + 2 - D) q / 2 + q^2 / 4
 * = 10.3 q^2 at n = 1024, D = 2 (run_splant checks it), 14 q^2 at D = 3 -- see XR below for how it does. */
#ifndef SPLANT_INCOMPLETE
#define SPLANT_INCOMPLETE 3
#endif
/* THREE stages (groups of eight, degree-7 polynomials mod x^8 - w^2) fit after all if BOTH operands come
 * out of the forward transform small: the X leg of every butterfly of the last forward stage takes the
 * Barrett step (as many Barrett steps as reducing all of a: n per product), so both results of that stage
 * are within q + 20 of zero, a sum of eight products stays below 8.03 q^2 + q^2/4 < M = 11.5 q^2, and the
 * group multiplication needs no Barrett step of its own.  64 + 7 products, 15 reductions and the 8 Barrett
 * steps: 155 instructions per eight coefficients against 182 for two groups of four and the stage between
 * them (XR below; run_splant checks the bound for the plan's q). */
template <int L>
struct SpDrop {
  static constexpr int H = SmallGeom<L>::H;
  static constexpr int V = SPLANT_INCOMPLETE < H ? SPLANT_INCOMPLETE : H;    /* stages left out */
  static constexpr int D = 1 << V;                                           /* registers per group */
  static constexpr bool XR = (V >= 3);           /* operands reduced by the last forward stage, not by the group */
  static constexpr bool XR_ROWS = XR && (H - V > 0);   /* that stage is the last one of the lane phase ...     */
  static constexpr bool XR_COLS = XR && (H - V == 0);  /* ... or of the register phase on the high index bits  */
};

/* per-lane twiddles of the lane phase, levels 0 .. H-1-V (LaneTw1 holds all H levels) */
template <int L>
struct LaneTwS {
  using Gm = SmallGeom<L>;
  static constexpr int LV = Gm::H - SpDrop<L>::V;
  static constexpr int PER_ROW = (1 << LV) - 1;
  uint32_t w[(1 << Gm::G) * (PER_ROW > 0 ? PER_ROW : 1)];
  __device__ __forceinline__ void load(const uint32_t *tab, int l) {
#pragma unroll
    for (int g = 0; g < (1 << Gm::G); g++) {
      const int row = (g << Gm::H) | l;
#pragma unroll
      for (int m = 0; m < LV; m++) {
        const uint32_t *src = tab + (1 << (Gm::R + m)) + (row << m);
        if (m == 0) {
          w[g * PER_ROW + 0] = __ldg(src);
        } else if (m == 1) {
          uint2 v = __ldg(reinterpret_cast<const uint2 *>(src));
          w[g * PER_ROW + 1] = v.x;
          w[g * PER_ROW + 2] = v.y;
        } else {
#pragma unroll
          for (int u = 0; u < (1 << m); u += 4) {
            uint4 v = __ldg(reinterpret_cast<const uint4 *>(src + u));
            w[g * PER_ROW + ((1 << m) - 1) + u + 0] = v.x;
            w[g * PER_ROW + ((1 << m) - 1) + u + 1] = v.y;
            w[g * PER_ROW + ((1 << m) - 1) + u + 2] = v.z;
            w[g * PER_ROW + ((1 << m) - 1) + u + 3] = v.w;
          }
        }
      }
    }
  }
  __device__ __forceinline__ uint32_t get(int g, int m, int u) const {
    return w[g * PER_ROW + ((1 << m) - 1) + u];
  }
};
/* per-lane Z of the group multiplication: 2^(H-V) per row, laid out like table level R + H - V */
template <int L>
struct LaneZeta {
  using Gm = SmallGeom<L>;
  static constexpr int M = Gm::H - SpDrop<L>::V;
  static constexpr int ZPR = 1 << M;
  int z[(1 << Gm::G) * ZPR];
  __device__ __forceinline__ void load(const uint32_t *tab, int l) {
#pragma unroll
    for (int g = 0; g < (1 << Gm::G); g++) {
      const int row = (g << Gm::H) | l;
      const uint32_t *src = tab + (row << M);
      if (M == 0) {
        z[g * ZPR] = (int)__ldg(src);
      } else if (M == 1) {
        uint2 v = __ldg(reinterpret_cast<const uint2 *>(src));
        z[g * ZPR + 0] = (int)v.x;
        z[g * ZPR + 1] = (int)v.y;
      } else {
#pragma unroll
        for (int u = 0; u < ZPR; u += 4) {
          uint4 v = __ldg(reinterpret_cast<const uint4 *>(src + u));
          z[g * ZPR + u + 0] = (int)v.x;
          z[g * ZPR + u + 1] = (int)v.y;
          z[g * ZPR + u + 2] = (int)v.z;
          z[g * ZPR + u + 3] = (int)v.w;
        }
      }
    }
  }
  __device__ __forceinline__ int get(int g, int u) const { return z[g * ZPR + u]; }
};

/* redc of a raw product (or sum of raw products): -x 2^-32 mod q, centred; _u leaves the last shift
 * pending */
__device__ __forceinline__ int sp_redc_u(int x, const SpRegs &G) {
  const int p = (int)((uint32_t)x * G.qinv);
  return (p >> 16) * G.q + G.dd;
}
__device__ __forceinline__ int sp_redc(int x, const SpRegs &G) { return sp_redc_u(x, G) >> 16; }
/* the group multiplication, lane-phase layout: group = registers r .. r + 2^V - 1, r a multiple of 2^V */
template <int L>
__device__ __forceinline__ void sp_groupmul(uint32_t (&xa)[SmallGeom<L>::NV], const uint32_t (&xb)[SmallGeom<L>::NV],
                                            const LaneZeta<L> &zt, const SpRegs &G) {
  using Gm = SmallGeom<L>;
  constexpr int V = SpDrop<L>::V, D = SpDrop<L>::D;
  static_assert(Gm::H >= 1, "groups live in the lane-phase layout");
#pragma unroll
  for (int r = 0; r < Gm::NV; r += D) {
    uint32_t a[D], lo[D], hi[D];
#pragma unroll
    for (int i = 0; i < D; i++) a[i] = SpDrop<L>::XR ? xa[r + i] : (uint32_t)sp_red((int)xa[r + i], G);
    const uint32_t zeta = (uint32_t)zt.get(r >> Gm::H, (r & (Gm::T - 1)) >> V);
#pragma unroll
    for (int k = 0; k < D; k++) { lo[k] = 0; hi[k] = 0; }
#pragma unroll
    for (int i = 0; i < D; i++)
#pragma unroll
      for (int j = 0; j < D; j++) {
        if (i + j < D) lo[i + j] += a[i] * xb[r + j];
        else hi[i + j - D] += a[i] * xb[r + j];
      }
#pragma unroll
    for (int k = 0; k < D; k++) {
      uint32_t c = lo[k];
      if (k < D - 1) c += (uint32_t)sp_redc((int)hi[k], G) * zeta;
      xa[r + k] = (uint32_t)(SPLANT_PENDING ? sp_redc_u((int)c, G) : sp_redc((int)c, G));
    }
  }
}
template <int L, bool XLAST = false>
__device__ __forceinline__ void sp_fwd_cols(uint32_t (&x)[SmallGeom<L>::NV], const SPlantParams<SmallGeom<L>::R> &P,
                                            const SpRegs &G) {
  using Gm = SmallGeom<L>;
#pragma unroll
  for (int s = 0; s < Gm::R; s++) {
    const int kb = Gm::R - 1 - s;
#pragma unroll
    for (int k = 0; k < Gm::NV; k++) {
      if (k & (1 << kb)) continue;
      sp_ct(x[k], x[k | (1 << kb)], P.ufwd[(1 << s) + (k >> (kb + 1))], G, sp_use_mad(pl_ord(k, kb) + s),
            XLAST && s == Gm::R - 1);
    }
  }
}
template <int L>
__device__ __forceinline__ void sp_fwd_rows(uint32_t (&x)[SmallGeom<L>::NV], const LaneTwS<L> &tw, const SpRegs &G) {
  using Gm = SmallGeom<L>;
#pragma unroll
  for (int lv = 0; lv < Gm::H - SpDrop<L>::V; lv++) {
    const int bit = Gm::H - 1 - lv;
#pragma unroll
    for (int r = 0; r < Gm::NV; r++) {
      if (r & (1 << bit)) continue;
      sp_ct(x[r], x[r | (1 << bit)], tw.get(r >> Gm::H, lv, (r & (Gm::T - 1)) >> (bit + 1)), G,
            sp_use_mad(pl_ord(r, bit) + lv + 1), SpDrop<L>::XR_ROWS && lv == Gm::H - SpDrop<L>::V - 1);
    }
  }
}
/* inverse, layout 2: register bits V .. H-1; inputs are products (bound 1), pending when
 * SPLANT_PENDING; the outputs are values (the products of the last stage are shifted here, and so are
 * the inputs when the phase has no stage at all) */
template <int L>
__device__ __forceinline__ void sp_inv_rows(uint32_t (&x)[SmallGeom<L>::NV], const LaneTwS<L> &tw, const SpRegs &G) {
  using Gm = SmallGeom<L>;
  constexpr bool PD = SPLANT_PENDING != 0;
  constexpr int SP_DROP = SpDrop<L>::V;
#pragma unroll
  for (int bit = SP_DROP; bit < Gm::H; bit++) {
    const int lv = Gm::H - 1 - bit;
#pragma unroll
    for (int r = 0; r < Gm::NV; r++) {
      if (r & (1 << bit)) continue;
      const int rl = r & (Gm::T - 1);
      const uint32_t wt = tw.get(r >> Gm::H, lv, rl >> (bit + 1));
      const bool red = 2 * sp_leg_bound(rl, bit, 1, SP_DROP) > SP_CAP;
      const bool pin = PD && (bit == SP_DROP || ((rl >> (bit - 1)) & 1));
      if (pin) sp_gs_r<true, PD>(red, x[r], x[r | (1 << bit)], wt, G);
      else sp_gs_r<false, PD>(red, x[r], x[r | (1 << bit)], wt, G);
    }
  }
  if (PD) {
#pragma unroll
    for (int r = 0; r < Gm::NV; r++) {
      const int rl = r & (Gm::T - 1);
      const bool pend = (Gm::H > SP_DROP) ? ((rl >> (Gm::H - 1)) & 1) : true;
      if (pend) x[r] = (uint32_t)((int)x[r] >> 16);
    }
  }
}
/* inverse, layout 1: register bits 0 .. R-1, inputs <= B_IN; the last stage multiplies both legs
 * and makes them canonical */
template <int L, int B_IN>
__device__ __forceinline__ void sp_inv_cols(uint32_t (&x)[SmallGeom<L>::NV], const SPlantParams<SmallGeom<L>::R> &P,
                                            const SpRegs &G) {
  using Gm = SmallGeom<L>;
#pragma unroll
  for (int kb = 0; kb < Gm::R; kb++) {
    const int t = 1 << (Gm::R - 1 - kb);
#pragma unroll
    for (int k = 0; k < Gm::NV; k++) {
      if (k & (1 << kb)) continue;
      const int k2 = k | (1 << kb);
      const bool pin = (SPLANT_PENDING != 0) && kb > 0 && ((k >> (kb - 1)) & 1);
      if (kb < Gm::R - 1) {
        const uint32_t wt = P.uinv[t + (k >> (kb + 1))];
        const bool red = 2 * sp_leg_bound(k, kb, B_IN) > SP_CAP;
        if (pin) sp_gs_r<true, SPLANT_PENDING != 0>(red, x[k], x[k2], wt, G);
        else sp_gs_r<false, SPLANT_PENDING != 0>(red, x[k], x[k2], wt, G);
      } else {
        static_assert(2 * SP_CAP <= 40, "sums and differences of two capped legs must stay multipliable");
        int d, s;
        if (pin) {
          const int xs = (int)x[k] >> 16;
          s = xs + ((int)x[k2] >> 16);
          asm("" : "+r"(s));
          d = 2 * xs - s;
        } else {
          d = (int)x[k] - (int)x[k2];
          s = (int)x[k] + (int)x[k2];
        }
        const uint32_t ty = (uint32_t)(sp_mul_u(d, P.last_y, G) >> 16);
        const uint32_t tx = (uint32_t)(sp_mul_u(s, P.last_x, G) >> 16);
        x[k2] = min(ty, ty + (uint32_t)G.q);            /* centred -> [0, q): one VIADDMNMX.U32 */
        x[k] = min(tx, tx + (uint32_t)G.q);
      }
    }
  }
}

/* TWREG: the lane-phase twiddles stay in registers for the whole kernel (n <= 256); from n = 512 up
 * they are read per tile through L1, as in polymul_plant_kernel, and the last quarter of the tiles comes
 * from the tail scheduler (PlantParams::sched). */
template <int L, int WARPS, int MINB, bool TWREG, typename IO = uint32_t, typename OIO = IO>
__global__ void __launch_bounds__(WARPS * 32, MINB)
polymul_splant_kernel(const __grid_constant__ SPlantParams<SmallGeom<L>::R> P) {
  using Gm = SmallGeom<L>;
  using Pg = PlantGeom<L, IO>;
  constexpr int SP_DROP = SpDrop<L>::V;
  static_assert(TWREG ? L <= 8 : L <= 10, "twiddles of the lane phase live in registers up to n = 256");
  static_assert(L + 2 <= 2 * SP_CAP, "forward values are never reduced");
  static_assert(Gm::H >= 1, "n >= 8: there is a lane phase (the pair multiplication and the pending products rely on it)");
  extern __shared__ __align__(16) uint32_t smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int sub = lane / Gm::T;
  const int l = lane % Gm::T;
  IO *pf_a = reinterpret_cast<IO *>(smem + warp * Pg::WARP_WORDS);
  IO *pf_b = reinterpret_cast<IO *>(smem + warp * Pg::WARP_WORDS + Pg::PF_WORDS);
  uint32_t *sm_a = smem + warp * Pg::WARP_WORDS + 2 * Pg::PF_WORDS + sub * Gm::STRIDE;
  uint32_t *sm_b = Pg::SHARE_XCHG ? sm_a : sm_a + Gm::PPW * Gm::STRIDE;
  const IO *ga = static_cast<const IO *>(P.a), *gb = static_cast<const IO *>(P.b);
  OIO *gc = static_cast<OIO *>(P.c);
  /* the constants every butterfly reads sit in ordinary registers (ntt_small_plant.cuh, PlRegs) */
  SpRegs G;
  G.q = (int)(P.q + P.zero);
  G.qinv = P.qinv;
  G.dd = (int)(P.dd + P.zero);
  G.cbar = (int)P.cbar;

  const unsigned long long ntiles = (P.batch + Gm::PPW - 1) / Gm::PPW;
  const unsigned long long wstride = (unsigned long long)gridDim.x * WARPS;
  constexpr bool DYN = (L >= PLANT_DYN_MINL);
  const bool dyn = DYN && P.sched != nullptr;
  uint32_t rounds_left = dyn ? P.static_rounds : 0xffffffffu;
  const unsigned long long dyn_base = (unsigned long long)P.static_rounds * wstride;
  unsigned long long tile = (unsigned long long)blockIdx.x * WARPS + warp;
  unsigned long long next = tile + wstride;
  unsigned long long pend = 0;
  if (dyn) rounds_left -= 2;

  const bool nowait = P.nowait != 0;
  if (nowait) asm volatile("griddepcontrol.launch_dependents;");    /* a launch that waits triggers after its wait */
  if (nowait && tile < ntiles) plant_prefetch<L, IO>(pf_a, pf_b, ga, gb, tile, P.batch, lane);
  LaneTwS<L> twf, twi;
  LaneZeta<L> zt;
  if (TWREG) {
    twf.load(P.tw_fwd, l);
    twi.load(P.tw_inv, l);
    if (SP_DROP) zt.load(P.zeta, l);
  }
  if (!nowait) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;");
    if (tile < ntiles) plant_prefetch<L, IO>(pf_a, pf_b, ga, gb, tile, P.batch, lane);
  }

  /* From n = 512 up the unrolled tile is longer than the 32 KB instruction cache of the SM (3 173
   * instructions = 50 KB at n = 1024) and every resident warp runs through it at its own place: ncu
   * shows `no_instruction` as the top stall (1.45 warps per issue cycle).  SPLANT_CTASYNC (tuning
   * experiment, off) keeps the warps of a CTA in step -- one block barrier at the head of every tile (2:
   * another one before the pair multiplication) -- so that the SM fetches one instruction stream per CTA
   * instead of one per warp.  Measured at c4: 264 M polymul/s with the barrier against 273 M without
   * (two barriers 262 M; ONE CTA of 12 warps with the barrier 272 M): what the barrier gives the
   * instruction cache, it takes back in waiting. */
#ifndef SPLANT_CTASYNC
#define SPLANT_CTASYNC 0
#endif
  constexpr bool CSYNC = (SPLANT_CTASYNC > 0) && (L >= 9) && (WARPS > 1);
  unsigned long long next2 = 0;
  for (;; tile = next, next = next2) {
    const bool have = tile < ntiles;
    if (CSYNC) {
      if (!__syncthreads_or(have ? 1 : 0)) break;
    } else if (!have) {
      break;
    }
    const unsigned long long poly = tile * Gm::PPW + sub;
    const bool live = have && poly < P.batch;
    uint32_t xa[Gm::NV], xb[Gm::NV];
    bool grab = false;
    if (have) {
    cp_async_wait_all();
    __syncwarp();
#pragma unroll
    for (int k = 0; k < Gm::NV; k++) {
      xa[k] = pf_a[sub * Pg::PSTRIDE + (k << Gm::H) + l];
      xb[k] = pf_b[sub * Pg::PSTRIDE + (k << Gm::H) + l];
    }
    __syncwarp();                                     /* prefetch buffers are free again */
    if (next < ntiles) plant_prefetch<L, IO>(pf_a, pf_b, ga, gb, next, P.batch, lane);
    /* the tile after next: round-robin while static rounds are left, then one grab from the counter */
    grab = dyn && rounds_left == 0;
    if (grab) {
      if (lane == 0) pend = atomicAdd(P.sched, 1ULL);
    } else {
      next2 = next + wstride;
      if (dyn) rounds_left--;
    }

    sp_fwd_cols<L, SpDrop<L>::XR_COLS>(xa, P, G);
    sp_fwd_cols<L, SpDrop<L>::XR_COLS>(xb, P, G);
    if (Gm::H > 0) {
      store_cols<L>(xa, sm_a, l);
      if (Pg::SHARE_XCHG) {                           /* one buffer: a goes through, then b */
        __syncwarp();
        load_rows<L>(xa, sm_a, l);
        __syncwarp();
      }
      store_cols<L>(xb, sm_b, l);
      __syncwarp();
      if (!Pg::SHARE_XCHG) load_rows<L>(xa, sm_a, l);
      load_rows<L>(xb, sm_b, l);
      if (!TWREG) twf.load(P.tw_fwd, l);
      sp_fwd_rows<L>(xa, twf, G);
      sp_fwd_rows<L>(xb, twf, G);
    }
    }
    if (CSYNC && SPLANT_CTASYNC >= 2) __syncthreads();
    if (have) {

    /* pointwise product (mul_array, R/NTT/ntt.C:131-137) as a Plantard product of two variables:
     * p = a b q^-1 gives -a b 2^-32 mod q, centred; the constant is cancelled by last_x / last_y.
     * With SPLANT_INCOMPLETE it is the pair multiplication above (same constant). */
    if (SP_DROP && Gm::H > 0) {
      if (!TWREG) zt.load(P.zeta, l);
      sp_groupmul<L>(xa, xb, zt, G);
    } else {
#pragma unroll
      for (int k = 0; k < Gm::NV; k++) {
        const int av = sp_red((int)xa[k], G);
        xa[k] = (uint32_t)(SPLANT_PENDING ? sp_redc_u(av * (int)xb[k], G) : sp_redc(av * (int)xb[k], G));
      }
    }

    if (Gm::H > 0) {
      if (!TWREG) twi.load(P.tw_inv, l);
      sp_inv_rows<L>(xa, twi, G);
      /* the registers that only ever took sums (index bits SP_DROP .. H-1 all zero) go back to the
       * centre, so that the second phase starts from the bound of the others */
      constexpr int worst = sp_phase_out(Gm::H, 1, SP_DROP);
      constexpr int second = sp_phase_out_mixed(Gm::H, 1, SP_DROP) > 2 ? sp_phase_out_mixed(Gm::H, 1, SP_DROP) : 2;
      if (worst > second) {
#pragma unroll
        for (int g = 0; g < (1 << Gm::G); g++)
#pragma unroll
          for (int e = 0; e < (1 << SP_DROP); e++)
            xa[(g << Gm::H) + e] = (uint32_t)sp_red((int)xa[(g << Gm::H) + e], G);
      }
      __syncwarp();
      store_rows<L>(xa, sm_a, l);
      __syncwarp();
      load_cols<L>(xa, sm_a, l);
      constexpr int b_in = (worst > second) ? second : worst;
      sp_inv_cols<L, b_in>(xa, P, G);
    } else {
      sp_inv_cols<L, 1>(xa, P, G);
    }
    if (live) {
      OIO *cp = gc + (poly << L);
#pragma unroll
      for (int k = 0; k < Gm::NV; k++) cp[(k << Gm::H) + l] = (OIO)xa[k];
    }
    __syncwarp();                                     /* smem reuse by the next tile */
    if (grab) next2 = dyn_base + __shfl_sync(0xffffffffu, pend, 0);
    }
  }
  if (dyn) plant_sched_done(P.sched, lane, wstride);
  if (nowait) asm volatile("griddepcontrol.wait;" ::: "memory");
}

}  // namespace nttb200
