/*
 * ntt_small_plant.cuh -- fused product kernel for HALF-WORD moduli (q <= 12385: 12289, 7681,
 * 3329 ...), n = 2^L <= 1024.  Same dataflow and layouts as ntt_small.cuh (CT forward
 * std->rev with psi folded in, pointwise, GS inverse rev->std with n^-1 folded into the last
 * stage; R/NTT/ntt.C:342-371, 131-137, 428-451) but a cheaper modular multiplication:
 *
 *   Plantard word-size multiplication by a constant w, stored as
 *        w~ = ((-w 2^32) mod q) * q^-1  mod 2^32 :
 *        T = umulhi(Y * w~, q)           (IMAD + IMAD.WIDE = 3 fmaheavy issue slots)
 *   gives  T = Y w mod q  EXACTLY CANONICAL in [0, q)  whenever  Y * q < 2^32.
 *   [p = Y w~ mod 2^32 satisfies p q = Y W + k 2^32 with W = (-w 2^32) mod q, so
 *    k = floor(p q / 2^32) when 0 <= Y W < 2^32, k == Y w (mod q) and 0 <= k < q.]
 *
 * Against the Shoup multiplication of modarith.cuh (IMAD.HI + 2 IMAD = 4 slots, result in
 * [0,2q), two table words per twiddle) this is 3 slots, a canonical result and ONE table
 * word: the butterfly is IMAD, IMAD.WIDE (or IMAD.HI), IADD3, IADD3 and -- measured with
 * nttb200_measure_int_peak(13) vs (3) -- runs 1.31x faster on the B200 integer pipe,
 * which is the unit that binds this kernel (profiles/).  The lane twiddles held in
 * registers halve, which is what leaves room for the next tile's operands to be
 * prefetched into shared memory with cp.async (128-bit, fully coalesced) while the current
 * tile is being transformed.
 *
 * Values are kept lazily in [0, b q) with the integer bound b of every register known AT
 * COMPILE TIME from closed forms (see "compile-time value bounds" below): a CT stage adds 1
 * to b, a GS sum doubles it, every product resets it to 1, and a single IADD+IMNMX
 * conditional subtraction is inserted exactly where a bound would pass the cap; all
 * multiplication inputs stay below PLANT_LIMB q = 28 q  (28 q^2 < 2^32 for q <= 12385).
 */
#pragma once
#include <stdint.h>
#include "ntt_small.cuh"

namespace nttb200 {

#ifndef PLANT_OPAQUE
#define PLANT_OPAQUE 1      /* tuning experiments, see DESIGN.md "what did not work" */
#endif
#ifndef PLANT_MULHI
#define PLANT_MULHI 3
#endif
#ifndef PLANT_BULK_STORE
#define PLANT_BULK_STORE 0  /* 1: results leave through cp.async.bulk (TMA) -- measured slower, DESIGN.md */
#endif
#ifndef PLANT_ADD3
#define PLANT_ADD3 0
#endif
#ifndef PLANT_STAGGER_NS
#define PLANT_STAGGER_NS 0     /* tuning experiment, see below: +1 % at best, inside the box-to-box spread */
#endif
#ifndef PLANT_STAGGER_MIN_TILES
#define PLANT_STAGGER_MIN_TILES 4
#endif
#ifndef PLANT_DYN_MINL
#define PLANT_DYN_MINL 9   /* sizes from 2^9 up hand the last tiles out dynamically (see PlantParams::sched) */
#endif
#ifndef PLANT_QREG
#define PLANT_QREG 1   /* which constants live in ordinary registers, see PlRegs */
#endif
constexpr int PLANT_LIMB = 28;
constexpr uint32_t PLANT_QMAX = 12385;     /* 28 * q * q < 2^32 */

template <int R>
struct PlantParams {
  const void *a;             /* [batch][n] of IO (uint32_t for the reference API, uint16_t for the  */
  const void *b;             /*  packed extension nttb200_polymul_batch_u16)                        */
  void *c;
  const uint32_t *tw_fwd;    /* device level table of w~, n entries           */
  const uint32_t *tw_inv;
  unsigned long long batch;
  uint32_t q, qinv;
  uint32_t last_x;           /* (-n^-1 2^32)~ : multiplier of the sum branch of the last stage */
  uint32_t last_y;           /* (-n^-1 2^32 p_inv[1])~ : multiplier of its diff branch         */
  uint32_t qmul[16];         /* i * q, so that "+ b q" is a constant-bank operand         */
  uint32_t zero;             /* always 0 (see add_alu)                                    */
  unsigned long long *sched; /* tail scheduler: [0] next chunk, [1] warps finished (both 0 between
                                launches); NULL = static round-robin assignment only          */
  uint32_t static_rounds;    /* tiles every warp takes round-robin before it turns to sched  */
  uint32_t nowait;           /* 1: this launch neither reads what the launches still in flight on its
                                stream write nor writes what they touch (nttb200_launch_independent):
                                it does not wait for them before it starts, only before it ENDS     */
  uint32_t ufwd[1 << R];     /* entries [1, 2^R) of the forward level table (w~)          */
  uint32_t uinv[1 << R];
};

/* Constants that the butterflies read most often, held in ordinary registers.  ptxas keeps kernel
 * parameters in UNIFORM registers (LDCU -> URx operands); measured on B200, the same kernel with q
 * in an ordinary register instead runs 2.5 % faster (c2 1 159 -> 1 188 M polymul/s), so the hottest
 * constants are copied once per thread through an addition with the always-zero parameter, which
 * ptxas cannot fold back into a parameter read.  PLANT_QREG: 0 none, 1 q, 2 + q 2q 4q 8q,
 * 3 + the twiddles of the first two forward stages, 4 + of the last inverse stages. */
struct PlRegs {
  uint32_t q, qinv;
  uint32_t qm[4];            /* q, 2q, 4q, 8q                                   */
  uint32_t tf[4], ti[4];     /* entries 1..3 of the forward / inverse level table */
};
template <typename PT>
__device__ __forceinline__ PlRegs pl_regs(const PT &P) {
  PlRegs r;
  const uint32_t z = (PLANT_QREG >= 1) ? P.zero : 0u;
  r.q = P.q + z;
  r.qinv = P.qinv + ((PLANT_QREG >= 2) ? z : 0u);
#pragma unroll
  for (int i = 0; i < 4; i++) {
    r.qm[i] = P.qmul[1 << i] + z;
    r.tf[i] = P.ufwd[i] + z;
    r.ti[i] = P.uinv[i] + z;
  }
  return r;
}
template <typename PT>
__device__ __forceinline__ uint32_t pl_qmul(int b, const PT &P, const PlRegs &G) {
  if (PLANT_QREG >= 2 && (b == 1 || b == 2 || b == 4 || b == 8)) return G.qm[b == 1 ? 0 : b == 2 ? 1 : b == 4 ? 2 : 3];
  return P.qmul[b];
}

/* floor(p q / 2^32).  Written as mul.wide + unpack rather than mul.hi: ptxas folds an add that
 * follows an IMAD.HI into its addend and, when the product has two such consumers (every
 * value in the GS network: X + Y and X - Y), DUPLICATES the 2-slot IMAD.HI -- 32 extra per
 * n=256 product.  IMAD.WIDE has the same issue cost and is left alone (-10 % fmaheavy slots,
 * +6..9 % throughput measured). */
__device__ __forceinline__ uint32_t mulhi_nofold(uint32_t p, uint32_t q) {
#if PLANT_MULHI == 3
  uint32_t hi, lo;
  asm("{ .reg .u64 w; mul.wide.u32 w, %2, %3; mov.b64 {%1, %0}, w; }" : "=r"(hi), "=r"(lo) : "r"(p), "r"(q));
  (void)lo;
  return hi;
#else
  return __umulhi(p, q);
#endif
}
__device__ __forceinline__ uint32_t plant_mul(uint32_t y, uint32_t wt, uint32_t q) {
  uint32_t t = mulhi_nofold(y * wt, q);
#if PLANT_OPAQUE
  asm("" : "+r"(t));
#endif
  return t;
}
/* Variant B of the same multiplication: only the UPPER HALF of p = Y w~ mod 2^32 meets q,
 *        T = ((p >> 16) + 1) q >> 16          (IMAD, SHF, IMAD, SHF)
 * [(p>>16 + 1) q 2^16 = p q + (2^16 - p_lo) q, so the quotient by 2^32 is the same k as above
 *  whenever Y W + 2^16 q < 2^32: every Y < 22 q for q <= 12385; tests/test_plantard_arith.py.]
 * Two multiplier slots and two ALU instructions instead of three multiplier slots.  Measured on
 * B200: with q in an ordinary register (PlRegs) every butterfly of the
 * n <= 256 kernels in this form is worth +7 % (c3 1 252 -> 1 343 M polymul/s; PLANT_B_NUM of every
 * PLANT_B_DEN butterflies: 1/4 1 243, 1/2 1 274, 3/4 1 316, all 1 343) although ptxas answers the
 * extra ALU instructions by turning butterfly additions into IMAD.IADD; at n = 1024 the longer
 * code misses the instruction cache (249 -> 213 M), hence PLANT_B_MAXL = 8. */
#ifndef PLANT_B_NUM
#define PLANT_B_NUM 1
#endif
#ifndef PLANT_B_DEN
#define PLANT_B_DEN 1
#endif
#ifndef PLANT_B_MAXL
#define PLANT_B_MAXL 8
#endif
constexpr int PLANT_B_LIMB = 22;
__host__ __device__ constexpr bool pl_use_b(int L, int i) {
  return PLANT_B_NUM > 0 && L <= PLANT_B_MAXL && (i % PLANT_B_DEN) < PLANT_B_NUM;
}
/* ordinal of the butterfly whose lower leg is register r at register bit `bit` */
__host__ __device__ constexpr int pl_ord(int r, int bit) { return ((r >> (bit + 1)) << bit) | (r & ((1 << bit) - 1)); }
__device__ __forceinline__ uint32_t plant_mul_b(uint32_t y, uint32_t wt, uint32_t q) {
  const uint32_t p = y * wt;
  uint32_t t = ((p >> 16) * q + q) >> 16;
#if PLANT_OPAQUE
  asm("" : "+r"(t));
#endif
  return t;
}
__device__ __forceinline__ uint32_t plant_mul_v(bool vb, uint32_t y, uint32_t wt, uint32_t q) {
  return vb ? plant_mul_b(y, wt, q) : plant_mul(y, wt, q);
}
/* a + b as a THREE-input add (z is a kernel parameter that is always 0): ptxas turns plain
 * two-input adds into IMAD.IADD to "balance" the pipes, but the fmaheavy pipe is the one
 * that binds this kernel; a three-input add can only be an ALU-pipe IADD3 */
__device__ __forceinline__ uint32_t add_alu(uint32_t a, uint32_t b, uint32_t z) {
#if PLANT_ADD3
  return a + b + z;
#else
  (void)z;
  return a + b;
#endif
}

/* ---- compile-time value bounds ----------------------------------------------------------
 * Every register holds a value in [0, b q) with b known at compile time from closed forms:
 *   forward CT      b = (stages done) + 1, the same for every register (<= 11 <= PLANT_LIMB)
 *   inverse GS      inside a register phase, at the stage on register bit `bit`, the two legs
 *                   of a butterfly share their history h = r & (2^bit - 1):
 *                     h != 0: last product at stage p = msb(h), then bit-1-p sums:  2^(bit-1-p)
 *                     h == 0: only sums since the phase input bound b_in:           b_in 2^bit
 *                   capped at PLANT_CAP = 8 by ONE conditional subtraction of 8q on a sum whose
 *                   inputs already were at the cap (so d = X - Y + b q < 16 q <= 28 q always).
 */
constexpr int PLANT_CAP = 8;

__host__ __device__ constexpr int pl_min(int a, int b) { return a < b ? a : b; }
__host__ __device__ constexpr int pl_msb(int h) { int p = -1; while (h) { p++; h >>= 1; } return p; }
/* bound of both legs of the GS butterfly on register bit `bit` of register r */
__host__ __device__ constexpr int pl_gs_bound(int r, int bit, int b_in) {
  const int h = r & ((1 << bit) - 1);
  if (h == 0) return pl_min(b_in << bit, PLANT_CAP);
  return pl_min(1 << (bit - 1 - pl_msb(h)), PLANT_CAP);
}
/* worst bound of any register after a GS register phase of `bits` stages */
__host__ __device__ constexpr int pl_gs_phase_out(int bits, int b_in) {
  return pl_min(b_in << bits, PLANT_CAP);
}
/* pointwise a*b needs ba*bb <= PLANT_LIMB: halve the larger bound until it does */
struct PlPointwise { int na, nb; int ha[4], hb[4]; };
__host__ __device__ constexpr PlPointwise pl_pointwise_plan(int b) {
  PlPointwise p{0, 0, {0, 0, 0, 0}, {0, 0, 0, 0}};
  int ba = b, bb = b;
  while (ba * bb > PLANT_LIMB) {
    if (ba >= bb) { ba = (ba + 1) / 2; p.ha[p.na++] = ba; }
    else { bb = (bb + 1) / 2; p.hb[p.nb++] = bb; }
  }
  return p;
}

/* per-lane twiddles of the layout-2 phase, one word each */
template <int L>
struct LaneTw1 {
  using Gm = SmallGeom<L>;
  static constexpr int PER_ROW = (1 << Gm::H) - 1;
  uint32_t w[(1 << Gm::G) * (PER_ROW > 0 ? PER_ROW : 1)];
  __device__ __forceinline__ void load(const uint32_t *tab, int l) {
#pragma unroll
    for (int g = 0; g < (1 << Gm::G); g++) {
      const int row = (g << Gm::H) | l;
#pragma unroll
      for (int m = 0; m < Gm::H; m++) {
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

/* CT butterfly: T = Y w mod q in [0,q); X' = X + T, Y' = X - T + q  (bounds grow by 1) */
__device__ __forceinline__ void pl_ct(uint32_t &X, uint32_t &Y, uint32_t wt, uint32_t q, uint32_t z,
                                      bool vb = false) {
  const uint32_t T = plant_mul_v(vb, Y, wt, q);
  Y = X - T + q;
  X = add_alu(X, T, z);
}
/* GS butterfly on two legs < b q: X' = X + Y (capped), Y' = (X - Y) w mod q in [0,q) */
template <int B, typename PT>
__device__ __forceinline__ void pl_gs(uint32_t &X, uint32_t &Y, uint32_t wt, const PT &P, const PlRegs &G, bool vb) {
  static_assert(2 * B <= PLANT_B_LIMB, "d = X - Y + B q < 2 B q must stay inside variant B's range");
  const uint32_t d = X - Y + pl_qmul(B, P, G);
  uint32_t s = add_alu(X, Y, P.zero);
  if (2 * B > PLANT_CAP) s = csub(s, pl_qmul(PLANT_CAP, P, G));
  X = s;
  Y = plant_mul_v(vb, d, wt, G.q);
}
template <typename PT>
__device__ __forceinline__ void pl_gs_b(int b, uint32_t &X, uint32_t &Y, uint32_t wt, const PT &P,
                                        const PlRegs &G, bool vb = false) {
  switch (b) {                                     /* b is a compile-time constant after unrolling */
    case 1: pl_gs<1>(X, Y, wt, P, G, vb); break;
    case 2: pl_gs<2>(X, Y, wt, P, G, vb); break;
    case 4: pl_gs<4>(X, Y, wt, P, G, vb); break;
    default: pl_gs<8>(X, Y, wt, P, G, vb); break;
  }
}

template <int L>
__device__ __forceinline__ void pl_fwd_cols(uint32_t (&x)[SmallGeom<L>::NV],
                                            const PlantParams<SmallGeom<L>::R> &P, const PlRegs &G) {
  using Gm = SmallGeom<L>;
  static_assert(L + 1 <= PLANT_B_LIMB, "forward values < (L+1) q must stay inside variant B's range");
#pragma unroll
  for (int s = 0; s < Gm::R; s++) {
    const int kb = Gm::R - 1 - s;
#pragma unroll
    for (int k = 0; k < Gm::NV; k++) {
      if (k & (1 << kb)) continue;
      const int ti = (1 << s) + (k >> (kb + 1));
      pl_ct(x[k], x[k | (1 << kb)], (PLANT_QREG >= 3 && ti < 4) ? G.tf[ti] : P.ufwd[ti], G.q, P.zero,
            pl_use_b(L, pl_ord(k, kb) + s));
    }
  }
}
template <int L>
__device__ __forceinline__ void pl_fwd_rows(uint32_t (&x)[SmallGeom<L>::NV], const LaneTw1<L> &tw,
                                            const PlantParams<SmallGeom<L>::R> &P, const PlRegs &G) {
  using Gm = SmallGeom<L>;
#pragma unroll
  for (int lv = 0; lv < Gm::H; lv++) {
    const int bit = Gm::H - 1 - lv;
#pragma unroll
    for (int r = 0; r < Gm::NV; r++) {
      if (r & (1 << bit)) continue;
      const int g = r >> Gm::H;
      const int u = (r & (Gm::T - 1)) >> (bit + 1);
      pl_ct(x[r], x[r | (1 << bit)], tw.get(g, lv, u), G.q, P.zero, pl_use_b(L, pl_ord(r, bit) + lv + 1));
    }
  }
}
/* inverse, layout 2 (register bits 0..H-1 of r_lo; the g bit of r is not a history bit) */
template <int L>
__device__ __forceinline__ void pl_inv_rows(uint32_t (&x)[SmallGeom<L>::NV], const LaneTw1<L> &tw,
                                            const PlantParams<SmallGeom<L>::R> &P, const PlRegs &G) {
  using Gm = SmallGeom<L>;
#pragma unroll
  for (int bit = 0; bit < Gm::H; bit++) {
    const int lv = Gm::H - 1 - bit;
#pragma unroll
    for (int r = 0; r < Gm::NV; r++) {
      if (r & (1 << bit)) continue;
      const int g = r >> Gm::H;
      const int u = (r & (Gm::T - 1)) >> (bit + 1);
      pl_gs_b(pl_gs_bound(r & (Gm::T - 1), bit, 1), x[r], x[r | (1 << bit)], tw.get(g, lv, u), P, G,
              pl_use_b(L, pl_ord(r, bit) + bit + 2));
    }
  }
}
/* inverse, layout 1; inputs < B_IN q; the last stage multiplies both branches (n^-1 folded
 * in), so the outputs are canonical */
template <int L, int B_IN>
__device__ __forceinline__ void pl_inv_cols(uint32_t (&x)[SmallGeom<L>::NV],
                                            const PlantParams<SmallGeom<L>::R> &P, const PlRegs &G) {
  using Gm = SmallGeom<L>;
#pragma unroll
  for (int kb = 0; kb < Gm::R; kb++) {
    const int t = 1 << (Gm::R - 1 - kb);
#pragma unroll
    for (int k = 0; k < Gm::NV; k++) {
      if (k & (1 << kb)) continue;
      const int k2 = k | (1 << kb);
      const int b = pl_gs_bound(k, kb, B_IN);
      if (kb < Gm::R - 1) {
        const int ti = t + (k >> (kb + 1));
        pl_gs_b(b, x[k], x[k2], (PLANT_QREG >= 4 && ti < 4) ? G.ti[ti] : P.uinv[ti], P, G,
                pl_use_b(L, pl_ord(k, kb) + kb));
      } else {
        const uint32_t d = x[k] - x[k2] + pl_qmul(b, P, G);
        const uint32_t s = add_alu(x[k], x[k2], P.zero);
        x[k2] = plant_mul_v(pl_use_b(L, 2 * pl_ord(k, kb)), d, P.last_y, G.q);
        x[k] = plant_mul_v(pl_use_b(L, 2 * pl_ord(k, kb) + 1), s, P.last_x, G.q);
      }
    }
  }
}

/* Tail scheduler (n >= 2^PLANT_DYN_MINL).  With a purely static assignment every warp gets the
 * same number of tiles and the grid lasts as long as its slowest SM: two launches kept in flight on
 * two streams, which lets the fast SMs run ahead into the next launch, measured +8 % at n = 1024
 * (profiles/r1_two_streams.txt, bench.py "two_streams").  So the last quarter of the tiles is
 * handed out from a device counter instead: a warp that has done its `static_rounds` round-robin
 * tiles grabs one tile at a time (one atomic by lane 0, issued a whole tile ahead of the prefetch
 * that needs it and only looked at when that tile ends): c4 249 -> 263.5 M polymul/s, level with
 * the two-stream figure.  At n = 256 the same loop measured no gain (tiles are four times shorter,
 * 2 368 warps on one counter, and the extra loop state costs 1.6 % at the 127-register cap), so
 * those sizes keep the static loop at compile time.  The last warp to finish zeroes the two words
 * for the next launch that is handed this slot. */
__device__ __forceinline__ void plant_sched_done(unsigned long long *sched, int lane, unsigned long long warps) {
  if (lane == 0) {
    __threadfence();
    if (atomicAdd(sched + 1, 1ULL) + 1 == warps) {      /* every warp has made its last grab */
      sched[0] = 0;
      sched[1] = 0;
      __threadfence();
    }
  }
}

/* cp.async (LDGSTS) 16-byte copy global -> shared */
__device__ __forceinline__ void cp_async16(uint32_t *smem_dst, const uint32_t *gsrc) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

#ifndef PLANT_SHARE_XCHG
#define PLANT_SHARE_XCHG 0
#endif
#ifndef PLANT_SEQ_AB
#define PLANT_SEQ_AB 0       /* tuning experiment (n >= 512): see polymul_plant_kernel */
#endif
template <int L, typename IO = uint32_t>
struct PlantGeom {
  using Gm = SmallGeom<L>;
  static constexpr int EPC = 16 / (int)sizeof(IO);              /* elements per 16-byte chunk */
  /* elements per polynomial in the prefetch buffer: 16-byte aligned rows, and consecutive
   * polynomials of a warp land on different banks for the layout-1 reads */
  static constexpr int PSTRIDE = Gm::N + ((Gm::T >= 32) ? 0 : ((Gm::T % EPC == 0) ? Gm::T : EPC));
  static constexpr int CHUNKS = Gm::N / EPC;                    /* 16-byte chunks per polynomial */
  static constexpr int ITERS = (Gm::PPW * CHUNKS + 31) / 32;     /* cp.async per lane per operand */
  static constexpr int PF_WORDS = (Gm::PPW * PSTRIDE * (int)sizeof(IO) + 15) / 16 * 4;   /* per operand */
  /* per warp: prefetch a, prefetch b, transposition a, transposition b.  From n = 512 up the two
   * operands take turns in ONE transposition buffer (PLANT_SHARE_XCHG): 12 instead of 16 KiB per warp
   * at n = 1024, which is what lets a fourth CTA (16 warps) fit into the SM's shared memory */
  static constexpr bool SEQ_AB = PLANT_SEQ_AB && (L >= 9) && !PLANT_BULK_STORE;
  static constexpr bool SHARE_XCHG = PLANT_SHARE_XCHG && (L >= 9) && !SEQ_AB;
  static constexpr int WARP_WORDS = 2 * PF_WORDS + (SHARE_XCHG ? 1 : 2) * Gm::PPW * Gm::STRIDE;
};

template <int L, typename IO>
__device__ __forceinline__ void plant_prefetch(IO *pa, IO *pb, const IO *ga, const IO *gb,
                                               unsigned long long tile, unsigned long long batch, int lane) {
  using Gm = SmallGeom<L>;
  using Pg = PlantGeom<L, IO>;
#pragma unroll
  for (int i = 0; i < Pg::ITERS; i++) {
    const int c = i * 32 + lane;
    const int sub = c / Pg::CHUNKS;
    const int cc = c % Pg::CHUNKS;
    const unsigned long long poly = tile * Gm::PPW + sub;
    if (sub < Gm::PPW && poly < batch) {
      const size_t go = ((size_t)poly << L) + cc * Pg::EPC;
      cp_async16(reinterpret_cast<uint32_t *>(pa + sub * Pg::PSTRIDE + cc * Pg::EPC),
                 reinterpret_cast<const uint32_t *>(ga + go));
      cp_async16(reinterpret_cast<uint32_t *>(pb + sub * Pg::PSTRIDE + cc * Pg::EPC),
                 reinterpret_cast<const uint32_t *>(gb + go));
    }
  }
}

/* =====================================================================================
 * Fused product kernel, half-word moduli.
 * ===================================================================================== */
#ifndef PLANT_MINB_SCALE
#define PLANT_MINB_SCALE 1   /* with PLANT_WARPS=4: 2 keeps 16 warps per SM as 4 CTAs (tuning experiment) */
#endif
#ifndef PLANT_BIG_MINB
#define PLANT_BIG_MINB 0     /* tuning experiment: resident CTAs asked for at n >= 512 (0 = MINB) */
#endif
template <int L, int WARPS, int MINB, bool TWREG, typename IO = uint32_t, typename OIO = IO>
__global__ void __launch_bounds__(WARPS * 32, (L <= 8) ? MINB * PLANT_MINB_SCALE : (PLANT_BIG_MINB ? PLANT_BIG_MINB : MINB))
polymul_plant_kernel(const __grid_constant__ PlantParams<SmallGeom<L>::R> P) {
  using Gm = SmallGeom<L>;
  using Pg = PlantGeom<L, IO>;
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
  OIO *gc = static_cast<OIO *>(P.c);                  /* result rows may be wider than the operands */
  const PlRegs G = pl_regs(P);
  const uint32_t q = G.q;

  const unsigned long long ntiles = (P.batch + Gm::PPW - 1) / Gm::PPW;
  const unsigned long long wstride = (unsigned long long)gridDim.x * WARPS;
  /* compile-time switch: the sizes below PLANT_DYN_MINL keep the purely static loop */
  constexpr bool DYN = (L >= PLANT_DYN_MINL);
  const bool dyn = DYN && P.sched != nullptr;
  uint32_t rounds_left = dyn ? P.static_rounds : 0xffffffffu;     /* static tiles still to hand out */
  const unsigned long long dyn_base = (unsigned long long)P.static_rounds * wstride;
  unsigned long long tile = (unsigned long long)blockIdx.x * WARPS + warp;
  unsigned long long next = tile + wstride;
  unsigned long long pend = 0;                                     /* lane 0: the grab in flight */
  /* (static_rounds >= 2, so the first two tiles are static ones) */
  if (dyn) rounds_left -= 2;

  /* programmatic dependent launch: let the next launch on the stream start its CTAs (and
   * their twiddle loads, which depend on nothing) while this grid drains; operands may be the
   * previous kernel's output, so they are only touched after griddepcontrol.wait */
  const bool nowait = P.nowait != 0;
  if (nowait) asm volatile("griddepcontrol.launch_dependents;");
  /* A launch whose operands have nothing to do with the launches still running on the stream (the
   * host compares address ranges, nttb200.cu:launch_independent) starts its first prefetch right
   * away: its CTAs fill the SMs that the previous grid's tail leaves idle -- what a caller with two
   * streams gets (+8 % at batch 2^16), for a single-stream caller.  Such a launch waits for its
   * predecessors at its END instead, so completion on the stream stays in order.  A launch that waits
   * triggers its dependents only after its wait (ntt_small.cuh, pdl_begin). */
  if (nowait && tile < ntiles) plant_prefetch<L, IO>(pf_a, pf_b, ga, gb, tile, P.batch, lane);
  LaneTw1<L> twf, twi;
  if (TWREG) {
    twf.load(P.tw_fwd, l);
    twi.load(P.tw_inv, l);
  }
  if (!nowait) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;");
  }
#if PLANT_STAGGER_NS > 0
  /* The warps of a launch start together and stay in step from tile to tile, so the four that
   * share a scheduler want the same pipes at the same time.  Starting them a fraction of a tile
   * apart measured +1.0 % (batch 2^16) to +1.4 % (2^20) at n = 256 on one box, the same for 450,
   * 900 and 1800 ns, and nothing on another: off by default. */
  if (L <= 8 && ntiles > wstride * PLANT_STAGGER_MIN_TILES)
    __nanosleep((unsigned)(((warp >> 2) + 2 * (blockIdx.x & 1)) * PLANT_STAGGER_NS));
#endif
  if (!nowait && tile < ntiles) plant_prefetch<L, IO>(pf_a, pf_b, ga, gb, tile, P.batch, lane);

  /* bulk (TMA) stores need 16-byte aligned rows in shared memory: every size but the tiniest */
  constexpr bool BULK = PLANT_BULK_STORE && (Gm::T >= 4) && (Gm::N * sizeof(IO) >= 16) && sizeof(IO) == sizeof(OIO);
  bool bulk_pending = false;
  unsigned long long next2 = 0;
  for (; tile < ntiles; tile = next, next = next2) {
    const unsigned long long poly = tile * Gm::PPW + sub;
    const bool live = poly < P.batch;

    uint32_t xa[Gm::NV], xb[Gm::NV];
    cp_async_wait_all();
    __syncwarp();
    if (Pg::SEQ_AB) {
      /* n >= 512, PLANT_SEQ_AB: ONE copy of the forward transform's code, run for a and then for b
       * (a real loop: the operand is picked by a pointer), NTT(a) parked in the warp's second buffer in
       * the meantime -- a third less code for the instruction cache at the price of the a/b interleaving */
      uint32_t *park = smem + warp * Pg::WARP_WORDS + 2 * Pg::PF_WORDS + Gm::PPW * Gm::STRIDE + lane;
#pragma unroll 1
      for (int op = 0; op < 2; op++) {
        const IO *pf = op ? pf_b : pf_a;
#pragma unroll
        for (int k = 0; k < Gm::NV; k++) xb[k] = pf[sub * Pg::PSTRIDE + (k << Gm::H) + l];
        if (op == 1) {
          __syncwarp();                               /* prefetch buffers are free again */
          if (next < ntiles) plant_prefetch<L, IO>(pf_a, pf_b, ga, gb, next, P.batch, lane);
        }
        pl_fwd_cols<L>(xb, P, G);
        if (Gm::H > 0) {
          __syncwarp();
          store_cols<L>(xb, sm_a, l);
          __syncwarp();
          load_rows<L>(xb, sm_a, l);
          if (!TWREG) twf.load(P.tw_fwd, l);
          pl_fwd_rows<L>(xb, twf, P, G);
        }
        if (op == 0) {
#pragma unroll
          for (int k = 0; k < Gm::NV; k++) park[k * 32] = xb[k];
        }
      }
#pragma unroll
      for (int k = 0; k < Gm::NV; k++) xa[k] = park[k * 32];
    } else {
#pragma unroll
    for (int k = 0; k < Gm::NV; k++) {
      xa[k] = pf_a[sub * Pg::PSTRIDE + (k << Gm::H) + l];
      xb[k] = pf_b[sub * Pg::PSTRIDE + (k << Gm::H) + l];
    }
    __syncwarp();                                     /* prefetch buffers are free again */
    if (next < ntiles) plant_prefetch<L, IO>(pf_a, pf_b, ga, gb, next, P.batch, lane);
    }
    /* the tile after next: round-robin while static rounds are left, then one grab from the
     * counter, issued here and only looked at when this tile ends */
    const bool grab = dyn && rounds_left == 0;
    if (grab) {
      if (lane == 0) pend = atomicAdd(P.sched, 1ULL);
    } else {
      next2 = next + wstride;
      if (dyn) rounds_left--;
    }

    if (!Pg::SEQ_AB) {
    pl_fwd_cols<L>(xa, P, G);
    pl_fwd_cols<L>(xb, P, G);
    if (BULK && bulk_pending) {                       /* the previous tile's result row has left sm_a */
      if (l == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncwarp();
    }
    if (Gm::H > 0) {
      store_cols<L>(xa, sm_a, l);
      if (Pg::SHARE_XCHG) {                             /* one buffer: a goes through, then b */
        __syncwarp();
        load_rows<L>(xa, sm_a, l);
        __syncwarp();
      }
      store_cols<L>(xb, sm_b, l);
      __syncwarp();
      if (!Pg::SHARE_XCHG) load_rows<L>(xa, sm_a, l);
      load_rows<L>(xb, sm_b, l);
      if (!TWREG) twf.load(P.tw_fwd, l);
      pl_fwd_rows<L>(xa, twf, P, G);
      pl_fwd_rows<L>(xb, twf, P, G);
    }
    }

    /* pointwise product (mul_array, R/NTT/ntt.C:131-137) as a Plantard product of two
     * variables: umulhi(a b q^-1, q) = -a b 2^-32 mod q, canonical, needs a b < 2^32;
     * the constant -2^32 is cancelled by last_x / last_y */
    constexpr PlPointwise pw = pl_pointwise_plan(L + 1);    /* both operands < (L+1) q here */
#pragma unroll
    for (int k = 0; k < Gm::NV; k++) {
      uint32_t av = xa[k], bv = xb[k];
      if (pw.na > 0) av = csub(av, P.qmul[pw.ha[0]]);
      if (pw.na > 1) av = csub(av, P.qmul[pw.ha[1]]);
      if (pw.na > 2) av = csub(av, P.qmul[pw.ha[2]]);
      if (pw.nb > 0) bv = csub(bv, P.qmul[pw.hb[0]]);
      if (pw.nb > 1) bv = csub(bv, P.qmul[pw.hb[1]]);
      if (pw.nb > 2) bv = csub(bv, P.qmul[pw.hb[2]]);
      uint32_t v = mulhi_nofold(av * bv * G.qinv, q);
#if PLANT_OPAQUE
      asm("" : "+r"(v));
#endif
      xa[k] = v;
    }

    if (Gm::H > 0) {
      if (!TWREG) twi.load(P.tw_inv, l);
      pl_inv_rows<L>(xa, twi, P, G);
      __syncwarp();                                   /* all lanes done reading sm_a */
      store_rows<L>(xa, sm_a, l);
      __syncwarp();
      load_cols<L>(xa, sm_a, l);
    }
    pl_inv_cols<L, pl_gs_phase_out(Gm::H, 1)>(xa, P, G);
    if (BULK) {
    /* result row -> shared memory in natural order -> ONE bulk (TMA) copy per polynomial:
     * cp.async.bulk.global.shared::cta, issued by the polynomial's first lane */
    {
      OIO *row = reinterpret_cast<OIO *>(sm_a);
      __syncwarp();                                   /* load_cols of this tile is done with sm_a */
#pragma unroll
      for (int k = 0; k < Gm::NV; k++) row[(k << Gm::H) + l] = (OIO)xa[k];
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (l == 0 && live) {
        const unsigned sa = (unsigned)__cvta_generic_to_shared(row);
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                     :: "l"(gc + (poly << L)), "r"(sa), "n"(Gm::N * (int)sizeof(OIO)) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      bulk_pending = true;
    }
    } else {
      if (live) {
        OIO *cp = gc + (poly << L);
#pragma unroll
        for (int k = 0; k < Gm::NV; k++) cp[(k << Gm::H) + l] = (OIO)xa[k];
      }
      __syncwarp();                                   /* smem reuse by the next tile */
    }
    if (grab) next2 = dyn_base + __shfl_sync(0xffffffffu, pend, 0);
  }
  if (BULK && bulk_pending && l == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  if (dyn) plant_sched_done(P.sched, lane, wstride);
  if (nowait) asm volatile("griddepcontrol.wait;" ::: "memory");   /* no grid ends before its predecessors */
}

/* bring a value < B q to [0, q) with ceil(log2 B) conditional subtractions */
template <int B, typename PT>
__device__ __forceinline__ uint32_t pl_canon(uint32_t x, const PT &P) {
  if (B > 8) x = csub(x, P.qmul[8]);
  if (B > 4) x = csub(x, P.qmul[4]);
  if (B > 2) x = csub(x, P.qmul[2]);
  if (B > 1) x = csub(x, P.qmul[1]);
  return x;
}

/* =====================================================================================
 * Standalone transforms for half-word moduli, in place on P.c (uint32 rows), canonical output.
 *   DIR 0: forward CT std->rev with the table P.tw_fwd / P.ufwd
 *   DIR 1: inverse GS rev->std with P.tw_inv / P.uinv; the last stage multiplies its sum branch
 *          by last_x and its diff branch by last_y (1 and p[1] when no scaling is wanted)
 * Same reference functions as ntt_small_kernel (the table decides which).
 * ===================================================================================== */
template <int L, int WARPS, int DIR>
__global__ void __launch_bounds__(WARPS * 32)
ntt_plant_kernel(const __grid_constant__ PlantParams<SmallGeom<L>::R> P) {
  using Gm = SmallGeom<L>;
  using Pg = PlantGeom<L, uint32_t>;
  extern __shared__ __align__(16) uint32_t smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int sub = lane / Gm::T;
  const int l = lane % Gm::T;
  /* per warp: [prefetch buffer (forward only)] [transposition buffer] */
  constexpr int WARP_WORDS = (DIR == 0 ? Pg::PF_WORDS : 0) + Gm::PPW * Gm::STRIDE;
  uint32_t *pf = smem + warp * WARP_WORDS;
  uint32_t *sm_a = pf + (DIR == 0 ? Pg::PF_WORDS : 0) + sub * Gm::STRIDE;
  uint32_t *data = static_cast<uint32_t *>(P.c);

  LaneTw1<L> tw;
  tw.load(DIR == 0 ? P.tw_fwd : P.tw_inv, l);
  const PlRegs G = pl_regs(P);
  if (!P.nowait) asm volatile("griddepcontrol.wait;" ::: "memory");     /* wait, then trigger: pdl_begin */
  asm volatile("griddepcontrol.launch_dependents;");

  const unsigned long long ntiles = (P.batch + Gm::PPW - 1) / Gm::PPW;
  const unsigned long long wstride = (unsigned long long)gridDim.x * WARPS;
  unsigned long long tile = (unsigned long long)blockIdx.x * WARPS + warp;
  /* forward: rows arrive by 128-bit cp.async one tile ahead (the column layout needs 32-bit
   * accesses, which are cheap in shared memory and half-efficient on HBM) */
  auto prefetch = [&](unsigned long long t) {
#pragma unroll
    for (int i = 0; i < Pg::ITERS; i++) {
      const int c = i * 32 + lane;
      const int s2 = c / Pg::CHUNKS, cc = c % Pg::CHUNKS;
      const unsigned long long poly = t * Gm::PPW + s2;
      if (s2 < Gm::PPW && poly < P.batch)
        cp_async16(pf + s2 * Pg::PSTRIDE + cc * 4, data + (poly << L) + cc * 4);
    }
  };
  if (DIR == 0 && tile < ntiles) prefetch(tile);

  for (; tile < ntiles; tile += wstride) {
    const unsigned long long poly = tile * Gm::PPW + sub;
    const bool live = poly < P.batch;
    const unsigned long long off = (live ? poly : 0ull) << L;
    uint32_t x[Gm::NV];
    if (DIR == 0) {
      cp_async_wait_all();
      __syncwarp();
#pragma unroll
      for (int k = 0; k < Gm::NV; k++) x[k] = pf[sub * Pg::PSTRIDE + (k << Gm::H) + l];
      __syncwarp();
      if (tile + wstride < ntiles) prefetch(tile + wstride);
      pl_fwd_cols<L>(x, P, G);
      if (Gm::H > 0) {
        store_cols<L>(x, sm_a, l);
        __syncwarp();
        load_rows<L>(x, sm_a, l);
        pl_fwd_rows<L>(x, tw, P, G);
      }
#pragma unroll
      for (int k = 0; k < Gm::NV; k++) x[k] = pl_canon<L + 1>(x[k], P);
      __syncwarp();
      if (live) gstore_rows<L>(x, data + off, l);
    } else {
      gload_rows<L>(x, data + off, l);
      if (Gm::H > 0) {
        pl_inv_rows<L>(x, tw, P, G);
        store_rows<L>(x, sm_a, l);
        __syncwarp();
        load_cols<L>(x, sm_a, l);
      }
      pl_inv_cols<L, pl_gs_phase_out(Gm::H, 1)>(x, P, G);
      __syncwarp();
      if (live) gstore_cols<L>(x, data + off, l);
    }
  }
  if (P.nowait) asm volatile("griddepcontrol.wait;" ::: "memory");
}

}  // namespace nttb200
