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
  unsigned long long batch;
  uint32_t q, qinv;
  uint32_t dd;               /* D: addend of the second product                             */
  uint32_t cbar;             /* C = round(2^26 / q): the Barrett step                        */
  uint32_t last_x, last_y;   /* centred forms of -n^-1 2^32 and -n^-1 2^32 p_inv[1]         */
  uint32_t zero;
  uint32_t nowait;           /* see PlantParams::nowait                                     */
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
/* CT: X' = X + T, Y' = X - T = 2X - X' */
__device__ __forceinline__ void sp_ct(uint32_t &X, uint32_t &Y, uint32_t wt, const SpRegs &G) {
  const int u = sp_mul_u((int)Y, wt, G);
  const int xn = (int)X + (u >> 16);
  Y = 2u * X - (uint32_t)xn;
  X = (uint32_t)xn;
}
/* GS: X' = X + Y (Barrett step when its bound would pass the cap), Y' = (X - Y) w */
template <bool REDUCE>
__device__ __forceinline__ void sp_gs(uint32_t &X, uint32_t &Y, uint32_t wt, const SpRegs &G) {
  const int d = (int)X - (int)Y;
  int s = (int)X + (int)Y;
  if (REDUCE) s = sp_red(s, G);
  X = (uint32_t)s;
  Y = (uint32_t)(sp_mul_u(d, wt, G) >> 16);
}

/* bound (units of q/2) of both legs of the GS butterfly on register bit `bit` of a register whose
 * low bits are `k`, in a register phase entered with every register <= b_in: stage s < bit left a
 * product (1) where bit s of k is set, else a sum (doubled; back to 1 when it passed the cap) */
__host__ __device__ constexpr int sp_leg_bound(int k, int bit, int b_in) {
  int b = b_in;
  for (int s = 0; s < bit; s++) {
    if ((k >> s) & 1) b = 1;
    else { b = 2 * b; if (b > SP_CAP) b = 1; }
  }
  return b;
}
/* worst bound of any register after a phase of `bits` stages */
__host__ __device__ constexpr int sp_phase_out(int bits, int b_in) {
  int worst = 1;
  for (int k = 0; k < (1 << bits); k++) {
    int b = sp_leg_bound(k, bits, b_in);
    if (b > worst) worst = b;
  }
  return worst;
}

template <int L>
__device__ __forceinline__ void sp_fwd_cols(uint32_t (&x)[SmallGeom<L>::NV], const SPlantParams<SmallGeom<L>::R> &P,
                                            const SpRegs &G) {
  using Gm = SmallGeom<L>;
#pragma unroll
  for (int s = 0; s < Gm::R; s++) {
    const int kb = Gm::R - 1 - s;
#pragma unroll
    for (int k = 0; k < Gm::NV; k++) {
      if (k & (1 << kb)) continue;
      sp_ct(x[k], x[k | (1 << kb)], P.ufwd[(1 << s) + (k >> (kb + 1))], G);
    }
  }
}
template <int L>
__device__ __forceinline__ void sp_fwd_rows(uint32_t (&x)[SmallGeom<L>::NV], const LaneTw1<L> &tw, const SpRegs &G) {
  using Gm = SmallGeom<L>;
#pragma unroll
  for (int lv = 0; lv < Gm::H; lv++) {
    const int bit = Gm::H - 1 - lv;
#pragma unroll
    for (int r = 0; r < Gm::NV; r++) {
      if (r & (1 << bit)) continue;
      sp_ct(x[r], x[r | (1 << bit)], tw.get(r >> Gm::H, lv, (r & (Gm::T - 1)) >> (bit + 1)), G);
    }
  }
}
/* inverse, layout 2: register bits 0 .. H-1; inputs are products (bound 1) */
template <int L>
__device__ __forceinline__ void sp_inv_rows(uint32_t (&x)[SmallGeom<L>::NV], const LaneTw1<L> &tw, const SpRegs &G) {
  using Gm = SmallGeom<L>;
#pragma unroll
  for (int bit = 0; bit < Gm::H; bit++) {
    const int lv = Gm::H - 1 - bit;
#pragma unroll
    for (int r = 0; r < Gm::NV; r++) {
      if (r & (1 << bit)) continue;
      const int rl = r & (Gm::T - 1);
      const uint32_t wt = tw.get(r >> Gm::H, lv, rl >> (bit + 1));
      if (2 * sp_leg_bound(rl, bit, 1) > SP_CAP) sp_gs<true>(x[r], x[r | (1 << bit)], wt, G);
      else sp_gs<false>(x[r], x[r | (1 << bit)], wt, G);
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
      if (kb < Gm::R - 1) {
        const uint32_t wt = P.uinv[t + (k >> (kb + 1))];
        if (2 * sp_leg_bound(k, kb, B_IN) > SP_CAP) sp_gs<true>(x[k], x[k2], wt, G);
        else sp_gs<false>(x[k], x[k2], wt, G);
      } else {
        static_assert(2 * SP_CAP <= 40, "sums and differences of two capped legs must stay multipliable");
        const int d = (int)x[k] - (int)x[k2];
        const int s = (int)x[k] + (int)x[k2];
        const uint32_t ty = (uint32_t)(sp_mul_u(d, P.last_y, G) >> 16);
        const uint32_t tx = (uint32_t)(sp_mul_u(s, P.last_x, G) >> 16);
        x[k2] = min(ty, ty + (uint32_t)G.q);            /* centred -> [0, q): one VIADDMNMX.U32 */
        x[k] = min(tx, tx + (uint32_t)G.q);
      }
    }
  }
}

template <int L, int WARPS, int MINB, typename IO = uint32_t, typename OIO = IO>
__global__ void __launch_bounds__(WARPS * 32, MINB)
polymul_splant_kernel(const __grid_constant__ SPlantParams<SmallGeom<L>::R> P) {
  using Gm = SmallGeom<L>;
  using Pg = PlantGeom<L, IO>;
  static_assert(L <= 8, "twiddles of the lane phase live in registers: n <= 256");
  static_assert(L + 2 <= 2 * SP_CAP, "forward values are never reduced");
  extern __shared__ __align__(16) uint32_t smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int sub = lane / Gm::T;
  const int l = lane % Gm::T;
  IO *pf_a = reinterpret_cast<IO *>(smem + warp * Pg::WARP_WORDS);
  IO *pf_b = reinterpret_cast<IO *>(smem + warp * Pg::WARP_WORDS + Pg::PF_WORDS);
  uint32_t *sm_a = smem + warp * Pg::WARP_WORDS + 2 * Pg::PF_WORDS + sub * Gm::STRIDE;
  uint32_t *sm_b = sm_a + Gm::PPW * Gm::STRIDE;
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
  unsigned long long tile = (unsigned long long)blockIdx.x * WARPS + warp;

  asm volatile("griddepcontrol.launch_dependents;");
  const bool nowait = P.nowait != 0;
  if (nowait && tile < ntiles) plant_prefetch<L, IO>(pf_a, pf_b, ga, gb, tile, P.batch, lane);
  LaneTw1<L> twf, twi;
  twf.load(P.tw_fwd, l);
  twi.load(P.tw_inv, l);
  if (!nowait) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (tile < ntiles) plant_prefetch<L, IO>(pf_a, pf_b, ga, gb, tile, P.batch, lane);
  }

  for (; tile < ntiles; tile += wstride) {
    const unsigned long long poly = tile * Gm::PPW + sub;
    const bool live = poly < P.batch;
    uint32_t xa[Gm::NV], xb[Gm::NV];
    cp_async_wait_all();
    __syncwarp();
#pragma unroll
    for (int k = 0; k < Gm::NV; k++) {
      xa[k] = pf_a[sub * Pg::PSTRIDE + (k << Gm::H) + l];
      xb[k] = pf_b[sub * Pg::PSTRIDE + (k << Gm::H) + l];
    }
    __syncwarp();                                     /* prefetch buffers are free again */
    if (tile + wstride < ntiles) plant_prefetch<L, IO>(pf_a, pf_b, ga, gb, tile + wstride, P.batch, lane);

    sp_fwd_cols<L>(xa, P, G);
    sp_fwd_cols<L>(xb, P, G);
    if (Gm::H > 0) {
      store_cols<L>(xa, sm_a, l);
      store_cols<L>(xb, sm_b, l);
      __syncwarp();
      load_rows<L>(xa, sm_a, l);
      load_rows<L>(xb, sm_b, l);
      sp_fwd_rows<L>(xa, twf, G);
      sp_fwd_rows<L>(xb, twf, G);
    }

    /* pointwise product (mul_array, R/NTT/ntt.C:131-137) as a Plantard product of two variables:
     * p = a b q^-1 gives -a b 2^-32 mod q, centred; the constant is cancelled by last_x / last_y */
#pragma unroll
    for (int k = 0; k < Gm::NV; k++) {
      const int av = sp_red((int)xa[k], G);
      const int p = (int)((uint32_t)(av * (int)xb[k]) * G.qinv);
      xa[k] = (uint32_t)(((p >> 16) * G.q + G.dd) >> 16);
    }

    if (Gm::H > 0) {
      sp_inv_rows<L>(xa, twi, G);
      /* the one register that only ever took sums (index bits all zero) goes back to the centre, so
       * that the second phase starts from the bound of the others */
      constexpr int worst = sp_phase_out(Gm::H, 1);
      constexpr int second = (Gm::H >= 1) ? sp_leg_bound(1, Gm::H, 1) : 1;
      if (worst > second) {
#pragma unroll
        for (int g = 0; g < (1 << Gm::G); g++) xa[g << Gm::H] = (uint32_t)sp_red((int)xa[g << Gm::H], G);
      }
      __syncwarp();
      store_rows<L>(xa, sm_a, l);
      __syncwarp();
      load_cols<L>(xa, sm_a, l);
      constexpr int b_in = (worst > second) ? (second > 2 ? second : 2) : worst;
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
  }
  if (nowait) asm volatile("griddepcontrol.wait;" ::: "memory");
}

}  // namespace nttb200
