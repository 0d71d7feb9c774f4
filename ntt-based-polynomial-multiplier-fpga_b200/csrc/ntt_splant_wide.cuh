/*
 * ntt_splant_wide.cuh -- fused product kernel for n = 1024 (and n = 512), half-word moduli: the arithmetic of
 * ntt_small_splant.cuh (signed Plantard, five-instruction butterflies, three stages left to a group
 * multiplication) in a geometry whose code FITS THE INSTRUCTION CACHE.
 *
 * Why: one warp per polynomial with 32 coefficients per lane (polymul_splant_kernel<10>) unrolls to
 * 3 036 instructions per tile = 48 KB, the SM's instruction cache holds 32 KB, and twelve resident warps
 * run through the loop each at its own place: ncu shows `no_instruction` as the top stall (1.45 warps per
 * issue cycle, issue slots 69 % busy against 82 % for the 20 KB loop of n = 256;
 * profiles/r2_c4_splant_ncu_full.txt).
 *
 * Geometry: still one warp per polynomial, but 16 registers per "virtual lane" and 64 virtual lanes: the
 * warp runs every phase for both halves of the virtual lanes (layouts A and C: a rolled loop, same code;
 * layout B: both halves at once, their elements arrive paired in 64-bit words) and the two operands of the
 * forward transform through ONE copy of the code (a rolled loop over a, b).  With 4
 * index bits in registers per layout it takes three layouts of the 10 index bits i9 .. i0:
 *
 *      layout A   registers i9 i8 i7 i6   half i5          lane i4 .. i0        stages on bits 9 .. 6
 *      layout B   registers i5 i4 i3 i2   half i0          lane i9 .. i6, i1    stages on bits 5 .. 3 (i2 rides along)
 *      layout C   registers i3 i2 i1 i0   half i9          lane i8 .. i4        two groups of eight: multiplication
 *
 * (forward A -> B -> C, inverse C -> B -> A; dataflow of R/NTT/ntt.C:342-371 and 428-451 as in
 * ntt_small_splant.cuh).  The polynomial lives in the warp's shared memory between the phases, every
 * phase reads the 16 words it owns and writes the same 16 words back, so the only synchronisation is
 * __syncwarp between phases.  Word i sits at  i + 2 (i >> 5)  (two words of padding per 32, so even words stay
 * even): layout A reads 32 consecutive words per register; layout B reads the pair i0 = 0, 1 -- its two
 * halves -- as ONE 64-bit word per register (per half-warp: banks 4 (i8 i7 i6) + 2 i1 + const, conflict-free);
 * layout C reads its 16 consecutive words as eight 64-bit words (per half-warp: banks 16 i4 + 2 (i7 i6 i5) +
 * const, conflict-free); and the address of register k is the lane's base plus a compile-time constant
 * (A: 68 k, B: 4 k + 2 (k >> 3), C: k), so no access costs address arithmetic.  Layout A reads the
 * operands straight from the cp.async prefetch buffer and writes the result straight to global memory, 128
 * contiguous bytes per register.  The twiddles of layout A are constant-bank operands (table entries 1 ..
 * 15), those of layout B depend on i9 .. i6 only, i.e. on the lane and not on the half: 15 + 15 words per
 * lane held in registers for the whole kernel; the four Z of layout C are one 16-byte load per half.
 *
 * Price: three more trips through shared memory per coefficient.  With two stages left out (the form the
 * captures were taken with): 3 194 instead of 3 087 warp instructions per polynomial (3 382 with 32-bit
 * accesses everywhere), loop 1 727 instructions = 28 KB, 95 registers; measured on B200
 * (profiles/r2_c4_splant_wide_v2_ncu_full.txt): `no_instruction` drops from 1.45 to 0.10 warps per issue cycle,
 * issue slots 69.5 % -> 83.1 % busy (3.32 instructions per clock and SM, what the n = 256 kernel reaches), c4
 * 279.5 -> 302 M polymul/s.  With three stages left out: loop 1 632 instructions, c4 305 - 307 M.
 *
 * n = 512 can run the same kernel (NTTB200_PLANT_N1024=2; measured level with the one-layout-per-phase
 * kernel there, 655 against 659 M polymul/s, so that one stays the default) with 32 virtual lanes (no halves) and three stages in layout B:
 *      layout A   registers i8 i7 i6 i5   lane i4 .. i0          stages on bits 8 .. 5
 *      layout B   registers i4 i3 i2 i1   lane i8 .. i5, i0      stages on bits 4, 3 (i2, i1 ride along)
 *      layout C   registers i3 i2 i1 i0   lane i8 .. i4          two groups of eight: multiplication
 * (32-bit accesses in B: its register pairs are not adjacent words).
 */
#pragma once
#include <stdint.h>
#include "ntt_small_splant.cuh"

namespace nttb200 {

/* geometry of the two sizes */
template <int L>
struct WideGeom {
  static_assert(L == 9 || L == 10, "n = 512 or 1024");
  static constexpr int N = 1 << L;
  static constexpr int HALVES = 1 << (L - 9);
  static constexpr int V = SPLANT_INCOMPLETE >= 3 ? 3 : 2;   /* stages left to the group multiplication */
  static constexpr int D = 1 << V;                  /* registers per group in layout C                  */
  static constexpr bool XR = (V >= 3);              /* the last stage of layout B reduces its X legs (SpDrop) */
  static constexpr int NB = L - 4 - V;              /* stages of layout B (on its upper register bits) */
  static constexpr int FIRST = 4 - NB;              /* lowest register bit of layout B that is a stage */
  static constexpr int PADMUL = 2;                  /* word i sits at i + 2 (i >> 5): even words stay even  */
  static constexpr int WK = N + PADMUL * (N >> 5);  /* words per padded polynomial                          */
  /* natural index of register k of layout A for virtual lane v = (half << 5) | lane */
  __device__ static __forceinline__ int a_index(int k, int v) { return (k << (L - 4)) | v; }
  /* padded addresses: the lane's base plus a compile-time constant per register */
  __device__ static __forceinline__ int a_base(int half, int lane) { return 34 * half + lane; }
  static constexpr int a_off(int k) { return (L == 10) ? 68 * k : 34 * k; }
  /* layout B.  n = 1024: the lane is (i9 .. i6, i1) and reads the PAIR i0 = 0, 1 -- the two halves -- as one
   * 64-bit word; n = 512: the lane is (i8 .. i5, i0), 32-bit accesses */
  __device__ static __forceinline__ int b_base(int hi4, int low) { return (L == 10) ? 68 * hi4 + 2 * low : 34 * hi4 + low; }
  static constexpr int b_off(int k) { return (L == 10) ? 4 * k + 2 * (k >> 3) : 2 * k; }
  /* layout C: 16 consecutive words from an even address, read and written as 64-bit words */
  __device__ static __forceinline__ int c_base(int v) { return (v << 4) + PADMUL * (v >> 1); }
};

/* NB Cooley-Tukey stages on register bits 3 .. 4-NB with the lane's twiddles (levels 0 .. NB-1 of
 * LaneTw1<8>: level m holds table entries (16 << m) + (row << m) + u, u < 2^m) */
template <int NB, bool XLAST>
__device__ __forceinline__ void wide_fwd_lane(uint32_t (&x)[16], const LaneTw1<8> &tw, const SpRegs &G) {
#pragma unroll
  for (int lv = 0; lv < NB; lv++) {
    const int bit = 3 - lv;
#pragma unroll
    for (int r = 0; r < 16; r++) {
      if (r & (1 << bit)) continue;
      sp_ct(x[r], x[r | (1 << bit)], tw.get(0, lv, r >> (bit + 1)), G, sp_use_mad(pl_ord(r, bit) + lv + 1),
            XLAST && lv == NB - 1);
    }
  }
}
/* NB Gentleman-Sande stages on register bits 4-NB .. 3; inputs are products (bound 1) */
template <int NB>
__device__ __forceinline__ void wide_inv_lane(uint32_t (&x)[16], const LaneTw1<8> &tw, const SpRegs &G) {
#pragma unroll
  for (int bit = 4 - NB; bit < 4; bit++) {
    const int lv = 3 - bit;
#pragma unroll
    for (int r = 0; r < 16; r++) {
      if (r & (1 << bit)) continue;
      const bool red = 2 * sp_leg_bound(r, bit, 1, 4 - NB) > SP_CAP;
      sp_gs_r<false, false>(red, x[r], x[r | (1 << bit)], tw.get(0, lv, r >> (bit + 1)), G);
    }
  }
}
/* group multiplication of layout C: 16 / D groups of D registers, Z of group g in z[g]; XR: the operands
 * arrive reduced (SpDrop), else a takes its Barrett step here */
template <int D, bool XR>
__device__ __forceinline__ void wide_groupmul(uint32_t (&xa)[16], const uint32_t (&xb)[16], const int (&z)[16 / D],
                                              const SpRegs &G) {
#pragma unroll
  for (int r = 0; r < 16; r += D) {
    uint32_t a[D], lo[D], hi[D];
#pragma unroll
    for (int i = 0; i < D; i++) a[i] = XR ? xa[r + i] : (uint32_t)sp_red((int)xa[r + i], G);
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
      if (k < D - 1) c += (uint32_t)sp_redc((int)hi[k], G) * (uint32_t)z[r / D];
      xa[r + k] = (uint32_t)sp_redc((int)c, G);
    }
  }
}

/* one polynomial per warp: 16-byte cp.async copies of both operands into the warp's prefetch buffers */
template <int L, typename IO>
__device__ __forceinline__ void wide_prefetch(IO *pa, IO *pb, const IO *ga, const IO *gb, unsigned long long poly,
                                              int lane) {
  constexpr int EPC = 16 / (int)sizeof(IO);
  constexpr int CHUNKS = (1 << L) / EPC;
  const size_t go = (size_t)poly << L;
#pragma unroll
  for (int c = 0; c < CHUNKS / 32; c++) {
    const int e = (c * 32 + lane) * EPC;
    cp_async16(reinterpret_cast<uint32_t *>(pa + e), reinterpret_cast<const uint32_t *>(ga + go + e));
    cp_async16(reinterpret_cast<uint32_t *>(pb + e), reinterpret_cast<const uint32_t *>(gb + go + e));
  }
}

template <int L, int WARPS, int MINB, typename IO = uint32_t, typename OIO = IO>
__global__ void __launch_bounds__(WARPS * 32, MINB)
polymul_splant_wide_kernel(const __grid_constant__ SPlantParams<4> P) {
  using W = WideGeom<L>;
  constexpr int N = W::N;
  constexpr int PF_WORDS = N * (int)sizeof(IO) / 4;
  extern __shared__ __align__(16) uint32_t smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  constexpr int WARP_WORDS = 2 * PF_WORDS + 2 * W::WK;
  IO *pf_a = reinterpret_cast<IO *>(smem + warp * WARP_WORDS);
  IO *pf_b = reinterpret_cast<IO *>(smem + warp * WARP_WORDS + PF_WORDS);
  uint32_t *wk_a = smem + warp * WARP_WORDS + 2 * PF_WORDS;
  uint32_t *wk_b = wk_a + W::WK;
  const IO *ga = static_cast<const IO *>(P.a), *gb = static_cast<const IO *>(P.b);
  OIO *gc = static_cast<OIO *>(P.c);
  SpRegs G;
  G.q = (int)(P.q + P.zero);
  G.qinv = P.qinv;
  G.dd = (int)(P.dd + P.zero);
  G.cbar = (int)P.cbar;

  const unsigned long long ntiles = P.batch;
  const unsigned long long wstride = (unsigned long long)gridDim.x * WARPS;
  const bool dyn = P.sched != nullptr;
  uint32_t rounds_left = dyn ? P.static_rounds : 0xffffffffu;
  const unsigned long long dyn_base = (unsigned long long)P.static_rounds * wstride;
  unsigned long long tile = (unsigned long long)blockIdx.x * WARPS + warp;
  unsigned long long next = tile + wstride;
  unsigned long long pend = 0;
  if (dyn) rounds_left -= 2;

  const bool nowait = P.nowait != 0;
  if (nowait) asm volatile("griddepcontrol.launch_dependents;");    /* a launch that waits triggers after its wait */
  if (nowait && tile < ntiles) wide_prefetch<L, IO>(pf_a, pf_b, ga, gb, tile, lane);
  /* layout B: the lane is (the four highest index bits, i0) */
  const int hi4 = lane >> 1, b0 = lane & 1;
  LaneTw1<8> twf, twi;
  twf.load(P.tw_fwd, hi4);
  twi.load(P.tw_inv, hi4);
  if (!nowait) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;");
    if (tile < ntiles) wide_prefetch<L, IO>(pf_a, pf_b, ga, gb, tile, lane);
  }

  unsigned long long next2 = 0;
  for (; tile < ntiles; tile = next, next = next2) {
    cp_async_wait_all();
    __syncwarp();
    /* forward transforms: a, then b, through one copy of the code */
#pragma unroll 1
    for (int op = 0; op < 2; op++) {
      const IO *pf = op ? pf_b : pf_a;
      uint32_t *wk = op ? wk_b : wk_a;
#pragma unroll 1
      for (int half = 0; half < W::HALVES; half++) {       /* layout A: the four highest stages */
        uint32_t x[16];
        const int v = (half << 5) | lane;
        const int base = W::a_base(half, lane);
#pragma unroll
        for (int k = 0; k < 16; k++) x[k] = (uint32_t)pf[W::a_index(k, v)];
        sp_fwd_cols<8>(x, P, G);
#pragma unroll
        for (int k = 0; k < 16; k++) wk[base + W::a_off(k)] = x[k];
      }
      __syncwarp();
      if (L == 10) {                                       /* layout B: both halves from 64-bit words */
        uint32_t x0[16], x1[16];
        const int base = W::b_base(hi4, b0);
#pragma unroll
        for (int k = 0; k < 16; k++) {
          const uint2 v = *reinterpret_cast<const uint2 *>(wk + base + W::b_off(k));
          x0[k] = v.x;
          x1[k] = v.y;
        }
        wide_fwd_lane<W::NB, W::XR>(x0, twf, G);
        wide_fwd_lane<W::NB, W::XR>(x1, twf, G);
#pragma unroll
        for (int k = 0; k < 16; k++) *reinterpret_cast<uint2 *>(wk + base + W::b_off(k)) = make_uint2(x0[k], x1[k]);
      } else {                                             /* layout B: the next NB stages */
        uint32_t x[16];
        const int base = W::b_base(hi4, b0);
#pragma unroll
        for (int k = 0; k < 16; k++) x[k] = wk[base + W::b_off(k)];
        wide_fwd_lane<W::NB, W::XR>(x, twf, G);
#pragma unroll
        for (int k = 0; k < 16; k++) wk[base + W::b_off(k)] = x[k];
      }
    }
    __syncwarp();                                          /* prefetch buffers are free again */
    if (next < ntiles) wide_prefetch<L, IO>(pf_a, pf_b, ga, gb, next, lane);
    const bool grab = dyn && rounds_left == 0;
    if (grab) {
      if (lane == 0) pend = atomicAdd(P.sched, 1ULL);
    } else {
      next2 = next + wstride;
      if (dyn) rounds_left--;
    }

#pragma unroll 1
    for (int half = 0; half < W::HALVES; half++) {         /* layout C: group multiplication */
      uint32_t xa[16], xb[16];
      const int v = (half << 5) | lane;                    /* index bits L-1 .. 4 */
      const int base = W::c_base(v);
#pragma unroll
      for (int k = 0; k < 16; k += 2) {
        const uint2 va = *reinterpret_cast<const uint2 *>(wk_a + base + k);
        const uint2 vb = *reinterpret_cast<const uint2 *>(wk_b + base + k);
        xa[k] = va.x; xa[k + 1] = va.y;
        xb[k] = vb.x; xb[k + 1] = vb.y;
      }
      int z[16 / W::D];
      if (W::D == 4) {
        const uint4 zv = __ldg(reinterpret_cast<const uint4 *>(P.zeta) + v);
        z[0] = (int)zv.x; z[1] = (int)zv.y; z[16 / W::D - 2] = (int)zv.z; z[16 / W::D - 1] = (int)zv.w;
      } else {
        const uint2 zv = __ldg(reinterpret_cast<const uint2 *>(P.zeta) + v);
        z[0] = (int)zv.x; z[16 / W::D - 1] = (int)zv.y;
      }
      wide_groupmul<W::D, W::XR>(xa, xb, z, G);
#pragma unroll
      for (int k = 0; k < 16; k += 2) *reinterpret_cast<uint2 *>(wk_a + base + k) = make_uint2(xa[k], xa[k + 1]);
    }
    __syncwarp();
    constexpr int worst = sp_phase_out(4, 1, W::FIRST), mixed = sp_phase_out_mixed(4, 1, W::FIRST);
    if (L == 10) {                                         /* layout B, inverse stages, both halves */
      uint32_t x0[16], x1[16];
      const int base = W::b_base(hi4, b0);
#pragma unroll
      for (int k = 0; k < 16; k++) {
        const uint2 v = *reinterpret_cast<const uint2 *>(wk_a + base + W::b_off(k));
        x0[k] = v.x;
        x1[k] = v.y;
      }
      wide_inv_lane<W::NB>(x0, twi, G);
      wide_inv_lane<W::NB>(x1, twi, G);
      /* the registers that only ever took sums go back to the centre (polymul_splant_kernel) */
      if (worst > mixed) {
#pragma unroll
        for (int e = 0; e < (1 << W::FIRST); e++) {
          x0[e] = (uint32_t)sp_red((int)x0[e], G);
          x1[e] = (uint32_t)sp_red((int)x1[e], G);
        }
      }
#pragma unroll
      for (int k = 0; k < 16; k++) *reinterpret_cast<uint2 *>(wk_a + base + W::b_off(k)) = make_uint2(x0[k], x1[k]);
    } else {                                               /* layout B, inverse stages */
      uint32_t x[16];
      const int base = W::b_base(hi4, b0);
#pragma unroll
      for (int k = 0; k < 16; k++) x[k] = wk_a[base + W::b_off(k)];
      wide_inv_lane<W::NB>(x, twi, G);
      if (worst > mixed) {
#pragma unroll
        for (int e = 0; e < (1 << W::FIRST); e++) x[e] = (uint32_t)sp_red((int)x[e], G);
      }
#pragma unroll
      for (int k = 0; k < 16; k++) wk_a[base + W::b_off(k)] = x[k];
    }
    __syncwarp();
    OIO *cp = gc + (tile << L);
#pragma unroll 1
    for (int half = 0; half < W::HALVES; half++) {         /* layout A, the last four inverse stages */
      uint32_t x[16];
      const int v = (half << 5) | lane;
      const int base = W::a_base(half, lane);
#pragma unroll
      for (int k = 0; k < 16; k++) x[k] = wk_a[base + W::a_off(k)];
      constexpr int b_in = (worst > mixed) ? (mixed > 2 ? mixed : 2) : worst;
      sp_inv_cols<8, b_in>(x, P, G);
#pragma unroll
      for (int k = 0; k < 16; k++) cp[W::a_index(k, v)] = (OIO)x[k];
    }
    __syncwarp();                                          /* shared memory reuse by the next tile */
    if (grab) next2 = dyn_base + __shfl_sync(0xffffffffu, pend, 0);
  }
  if (dyn) plant_sched_done(P.sched, lane, wstride);
  if (nowait) asm volatile("griddepcontrol.wait;" ::: "memory");
}

}  // namespace nttb200
