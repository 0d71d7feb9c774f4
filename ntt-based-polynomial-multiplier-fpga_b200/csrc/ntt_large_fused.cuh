/*
 * ntt_large_fused.cuh -- ONE persistent kernel for the whole large-n product (n = 2^15, 2^16):
 *
 *   a thread-block CLUSTER of 8 CTAs owns one polynomial product at a time and walks it through
 *   the three passes of ntt_large.cuh without leaving the kernel:
 *
 *     pass 1  columns forward  CTA r transforms the 32-column tile r of a and of b   (reads a, b from
 *                              HBM with 128-byte row segments, writes a', b' to the cluster's scratch)
 *     -- cluster barrier (release / acquire) --
 *     pass 2  rows             CTA r takes rows [r R/8, (r+1) R/8): forward rows of a' and b', pointwise
 *                              product, inverse rows; c' overwrites a' row by row
 *     -- cluster barrier --
 *     pass 3  columns inverse  CTA r transforms tile r of c' with n^-1 folded into the last stage and
 *                              writes c to HBM
 *     -- cluster barrier, split: arrive as soon as c' has been read, wait before the next
 *        polynomial's first scratch store --
 *
 * The dataflow is exactly the one of the three kernels it replaces (CT std->rev of
 * R/NTT/ntt.C:342-371 split at the row/column boundary, mul_array ntt.C:131-137, GS rev->std
 * ntt.C:428-451); what changes is where the intermediate a', b', c' live and who waits for whom:
 *
 *   - the scratch is (resident clusters) x 2n words -- 28 to 55 MB at n = 2^16 -- instead of a
 *     chunk of the batch: it stays in the 126 MB L2, so HBM sees the algorithmic 12 n bytes per
 *     product (a, b in, c out) and the 24 n bytes of intermediate traffic are L2 hits;
 *   - one launch for the whole batch: no grid-wide boundary between the passes, no ramp and tail
 *     per chunk; clusters drift apart, so column passes (memory-latency heavy) of some overlap
 *     row passes (arithmetic heavy) of others on the same SM -- the three-launch pipeline ran
 *     with the SMs idle 20-36 % of each launch and long_scoreboard as its top stall
 *     (profiles/r1_c5_large_v2_ncu_full.txt);
 *   - dependencies are cluster-local hardware barriers (barrier.cluster, ~380 cycles) instead of
 *     kernel boundaries.
 *
 * Scratch reads use ld.global.cg (L2 only): the lines are rewritten by other SMs of the cluster
 * during the kernel, so the non-coherent path of the three-launch kernels (__ldg) would be wrong here.
 */
#pragma once
#include <stdint.h>
#include "ntt_large.cuh"

namespace nttb200 {

constexpr int FUSED_CLUSTER = 8;          /* CTAs per polynomial = 32-column tiles per 256-word row */
constexpr int FUSED_WARPS = 8;            /* 256 threads per CTA                                    */

struct FusedParams {
  const uint32_t *a, *b;
  uint32_t *c;
  uint32_t *scratch;            /* [clusters][2][n]: a' (then c') and b' of the cluster's polynomial */
  const uint2 *tab;             /* forward level table, n entries (w, floor(w 2^32/q))               */
  const uint2 *tab_inv;
  unsigned long long batch;
  ModQ m;
  uint2 last_x, last_y;         /* multipliers of the very last inverse stage (n^-1 2^32 folded in)  */
  uint2 one;
  uint32_t zero;
  uint2 ufwd[32];               /* entries [1, 32) of the forward / inverse level tables: the         */
  uint2 uinv[32];               /* twiddles of the register phase on the high row bits (uniform)      */
};

template <int K1>
struct FusedGeom {
  static constexpr int RB = 3;                       /* low row bits: one per warp of the CTA          */
  static constexpr int RA = K1 - RB;                 /* high row bits: register bits of the first phase */
  static constexpr int NV = 1 << RA;
  static constexpr int GB = 1 << (RA - RB);
  static constexpr int ROWS = 1 << K1;
  static constexpr int ROWS_PER_CTA = ROWS / FUSED_CLUSTER;
  static constexpr int COL_SMEM_WORDS = ROWS * 32;
  static_assert(RA >= RB && RA <= 5, "fused kernel: 2^6 <= rows <= 2^8");
  static_assert(ROWS_PER_CTA % (2 * FUSED_WARPS) == 0, "every half-warp takes whole rows");
};

__device__ __forceinline__ void cluster_arrive() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait() {
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t ld_cg(const uint32_t *p) { return __ldcg(p); }

/* ---- pass 1: one 32-column tile, all rows, stages 0 .. K1-1 of the CT dataflow ---------------- */
template <int K1, int ARITH>
__device__ __forceinline__ void fused_cols_fwd(const uint32_t *src, uint32_t *dst, uint32_t *sm, int w,
                                               const FusedParams &P, const ModQ &m) {
  using G = FusedGeom<K1>;
  constexpr int lr = LARGE_LR;
  uint32_t x[G::NV];
#pragma unroll
  for (int k = 0; k < G::NV; k++) x[k] = __ldcg(src + ((size_t)((k << G::RB) | w) << lr));
#pragma unroll
  for (int s = 0; s < G::RA; s++) {
    const int bit = G::RA - 1 - s;
#pragma unroll
    for (int k = 0; k < G::NV; k++) {
      if (k & (1 << bit)) continue;
      const uint2 tw = P.ufwd[(1 << s) + (k >> (bit + 1))];
      if (ARITH == ARITH_CANON && bit > 0 && ((k >> (bit - 1)) & 1))
        ct_bfly<ARITH, true>(x[k], x[k | (1 << bit)], tw.x, tw.y, m);
      else
        ct_bfly<ARITH>(x[k], x[k | (1 << bit)], tw.x, tw.y, m);
    }
  }
#pragma unroll
  for (int k = 0; k < G::NV; k++) sm[((k << G::RB) | w) * 32] = x[k];
  __syncthreads();
#pragma unroll
  for (int g = 0; g < G::GB; g++) {
    const int hfix = w * G::GB + g;
#pragma unroll
    for (int kk = 0; kk < (1 << G::RB); kk++) x[(g << G::RB) + kk] = sm[((hfix << G::RB) | kk) * 32];
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
        if (ARITH == ARITH_CANON && bit > 0 && ((kk >> (bit - 1)) & 1))
          ct_bfly<ARITH, true>(x[(g << G::RB) + kk], x[(g << G::RB) + (kk | (1 << bit))], tw.x, tw.y, m);
        else
          ct_bfly<ARITH>(x[(g << G::RB) + kk], x[(g << G::RB) + (kk | (1 << bit))], tw.x, tw.y, m);
      }
    }
  }
#pragma unroll
  for (int g = 0; g < G::GB; g++) {
    const int hfix = w * G::GB + g;
#pragma unroll
    for (int kk = 0; kk < (1 << G::RB); kk++)
      dst[(size_t)((hfix << G::RB) | kk) << lr] = x[(g << G::RB) + kk];
  }
  __syncthreads();                                     /* the tile buffer is free again */
}

/* ---- pass 3: the last K1 stages of the GS dataflow on one 32-column tile ------------------------ */
template <int K1, int ARITH>
__device__ __forceinline__ void fused_cols_inv_load(uint32_t (&x)[FusedGeom<K1>::NV], const uint32_t *src, int w) {
  using G = FusedGeom<K1>;
  constexpr int lr = LARGE_LR;
#pragma unroll
  for (int g = 0; g < G::GB; g++) {
    const int hfix = w * G::GB + g;
#pragma unroll
    for (int kk = 0; kk < (1 << G::RB); kk++)
      x[(g << G::RB) + kk] = ld_cg(src + ((size_t)((hfix << G::RB) | kk) << lr));
  }
}
template <int K1, int ARITH>
__device__ __forceinline__ void fused_cols_inv(uint32_t (&x)[FusedGeom<K1>::NV], uint32_t *dst, uint32_t *sm, int w,
                                               const FusedParams &P, const ModQ &m) {
  using G = FusedGeom<K1>;
  constexpr int lr = LARGE_LR;
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
        gs_bfly<ARITH>(x[(g << G::RB) + kk], x[(g << G::RB) + (kk | (1 << bit))], tw.x, tw.y, m, yb);
      }
    }
  }
#pragma unroll
  for (int g = 0; g < G::GB; g++) {
    const int hfix = w * G::GB + g;
#pragma unroll
    for (int kk = 0; kk < (1 << G::RB); kk++) sm[((hfix << G::RB) | kk) * 32] = x[(g << G::RB) + kk];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < G::NV; k++) x[k] = sm[((k << G::RB) | w) * 32];
#pragma unroll
  for (int bit = 0; bit < G::RA; bit++) {
    const uint32_t yb = m.q2 << (G::RB + bit);
#pragma unroll
    for (int k = 0; k < G::NV; k++) {
      if (k & (1 << bit)) continue;
      if (bit < G::RA - 1) {
        const uint2 tw = P.uinv[(1 << (G::RA - 1 - bit)) + (k >> (bit + 1))];
        gs_bfly<ARITH>(x[k], x[k | (1 << bit)], tw.x, tw.y, m, yb);
      } else {
        gs_last<ARITH>(x[k], x[k | (1 << bit)], yb, P.last_x, P.last_y, m);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < G::NV; k++) __stcs(dst + ((size_t)((k << G::RB) | w) << lr), x[k]);
  __syncthreads();
}

/* ---- pass 2: row j of a' and b' -> row j of c' (in place over a'), one row pair per half-warp ---- */
template <int LR, int ARITH>
__device__ __forceinline__ void fused_row(uint32_t *ra, const uint32_t *rb, uint32_t *sm_a, uint32_t *sm_b, int l,
                                          uint32_t k1, uint32_t j, const FusedParams &P, const ModQ &m) {
  using Gm = SmallGeom<LR>;
  uint32_t xa[Gm::NV], xb[Gm::NV];
#pragma unroll
  for (int k = 0; k < Gm::NV; k++) {
    xa[k] = ld_cg(ra + (k << Gm::H) + l);
    xb[k] = ld_cg(rb + (k << Gm::H) + l);
  }
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
    uint32_t v = mont_mul(av, bv, m);                 /* (0, 2q); the 2^-32 is cancelled in pass 3 */
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
#pragma unroll
  for (int k = 0; k < Gm::NV; k++) ra[(k << Gm::H) + l] = xa[k];
  __syncwarp();                                        /* the warp's transposition buffers are free again */
}

#ifndef FUSED_MINB
#define FUSED_MINB 3
#endif
template <int K1, int ARITH>
__global__ void __launch_bounds__(FUSED_WARPS * 32, FUSED_MINB)
large_fused_polymul_kernel(const __grid_constant__ FusedParams P) {
  using G = FusedGeom<K1>;
  using Gm = SmallGeom<LARGE_LR>;
  constexpr int lr = LARGE_LR;
  constexpr int logn = K1 + lr;
  extern __shared__ __align__(16) uint32_t smem[];
  const int lane = threadIdx.x & 31;
  const int w = threadIdx.x >> 5;
  const uint32_t rank = blockIdx.x % FUSED_CLUSTER;            /* == %cluster_ctarank for (8,1,1) clusters */
  const unsigned long long cid = blockIdx.x / FUSED_CLUSTER;
  const unsigned long long nclusters = gridDim.x / FUSED_CLUSTER;
  const ModQ m = modq_regs(P.m, P.zero);
  uint32_t *sa = P.scratch + ((size_t)cid << (logn + 1));      /* a', later c' */
  uint32_t *sb = sa + ((size_t)1 << logn);                     /* b'           */
  const size_t tile_off = (size_t)rank * 32 + lane;
  uint32_t *sm_col = smem + lane;                              /* column passes: [row][32 lanes] */
  /* row pass: two transposition buffers per row pair, per half-warp */
  const int sub = lane / Gm::T, l = lane % Gm::T;
  uint32_t *sm_a = smem + (w * 2 * Gm::PPW + sub) * Gm::STRIDE;
  uint32_t *sm_b = sm_a + Gm::PPW * Gm::STRIDE;

  bool pending = false;                                        /* a split barrier is waiting for its second half */
  for (unsigned long long poly = cid; poly < P.batch; poly += nclusters) {
    const size_t po = (size_t)poly << logn;
    /* ---- pass 1 ---- */
    if (pending) {
      /* second half of the split barrier: every CTA of the cluster has read its tile of the previous
       * c', so the scratch may be overwritten (the peers' pass-3 arithmetic and stores go on) */
      cluster_wait();
      pending = false;
    }
    fused_cols_fwd<K1, ARITH>(P.a + po + tile_off, sa + tile_off, sm_col, w, P, m);
    fused_cols_fwd<K1, ARITH>(P.b + po + tile_off, sb + tile_off, sm_col, w, P, m);
    cluster_arrive();
    cluster_wait();
    /* ---- pass 2 ---- */
#pragma unroll 1
    for (int round = 0; round < G::ROWS_PER_CTA / (2 * FUSED_WARPS); round++) {
      const uint32_t j = rank * G::ROWS_PER_CTA + round * (2 * FUSED_WARPS) + w * 2 + sub;
      fused_row<lr, ARITH>(sa + ((size_t)j << lr), sb + ((size_t)j << lr), sm_a, sm_b, l, (uint32_t)K1, j, P, m);
    }
    __syncthreads();                                           /* smem changes hands: rows -> column tile */
    cluster_arrive();
    cluster_wait();
    /* ---- pass 3 ---- */
    {
      uint32_t x[G::NV];
      fused_cols_inv_load<K1, ARITH>(x, sa + tile_off, w);
      /* c' is in registers: the next polynomial's pass 1 may overwrite the scratch as far as this
       * CTA is concerned */
      cluster_arrive();
      pending = true;
      fused_cols_inv<K1, ARITH>(x, P.c + po + tile_off, sm_col, w, P, m);
    }
  }
  if (pending) cluster_wait();                                 /* nobody leaves while a peer may still wait */
}


/* =====================================================================================
 * The same three passes as a DATAFLOW: one persistent grid, no clusters, no barrier between CTAs.
 * Every CTA draws tickets from a device counter; ticket t is a fixed task
 *
 *     super-step i = t / 40:   16 x pass 1 of polynomial i        (operand, 32-column tile)
 *                              16 x pass 2 of polynomial i - D    (16 rows each)
 *                               8 x pass 3 of polynomial i - 2D   (32-column tile)
 *
 * and runs as soon as what it reads is there: pass 2 of p after the 16 pass-1 tasks of p, pass 3
 * after its 16 pass-2 tasks, pass 1 of p after pass 3 of p - W has released the scratch slot
 * p mod W (per-polynomial counters, release/acquire through L2).  A task only ever waits for tasks
 * with SMALLER tickets, and tickets are handed to running CTAs in order, so the wait ends without
 * any co-residency assumption; D super-steps (640 tickets, against < 600 CTAs in flight) between a
 * polynomial's passes make it a formality.  What this buys over the cluster version above: no CTA
 * idles at a cluster barrier for the slowest of its seven peers (19 % of all warp time there,
 * profiles/r2_c5_fused_v1_ncu_full.txt), every SM holds its full complement of CTAs (clusters of 8
 * leave 60 of 444 slots empty), and column passes of some polynomials always overlap row passes of
 * others on the same SM.  The scratch is W x 2n words (W = 48: 24 MB at n = 2^16) and lives in L2.
 * ===================================================================================== */
constexpr int FLOW_TASKS = 40;            /* per super-step: 16 + 16 + 8 */
constexpr int FLOW_DELAY = 16;            /* D */
constexpr int FLOW_SLOTS = 48;            /* W > 2 D */

struct FlowParams {
  FusedParams f;                /* f.scratch = [W][2][n] */
  unsigned *ctl;                /* [0] next ticket; then done1[batch], done2[batch], done3[batch], zero at launch */
};

__device__ __forceinline__ unsigned ld_acquire(const unsigned *p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void flow_wait(const unsigned *cnt, unsigned need) {
  if (threadIdx.x == 0) {
    unsigned ns = 32;
    while (ld_acquire(cnt) < need) {
      __nanosleep(ns);
      if (ns < 1024) ns <<= 1;
    }
  }
  __syncthreads();
}

template <int K1, int ARITH>
__global__ void __launch_bounds__(FUSED_WARPS * 32, FUSED_MINB)
large_flow_polymul_kernel(const __grid_constant__ FlowParams Q) {
  using G = FusedGeom<K1>;
  using Gm = SmallGeom<LARGE_LR>;
  constexpr int lr = LARGE_LR;
  constexpr int logn = K1 + lr;
  static_assert(G::ROWS / (2 * FUSED_WARPS) == 16, "16 row tasks per polynomial");
  const FusedParams &P = Q.f;
  extern __shared__ __align__(16) uint32_t smem[];
  __shared__ unsigned s_ticket;
  const int lane = threadIdx.x & 31;
  const int w = threadIdx.x >> 5;
  const ModQ m = modq_regs(P.m, P.zero);
  uint32_t *sm_col = smem + lane;
  const int sub = lane / Gm::T, l = lane % Gm::T;
  uint32_t *sm_a = smem + (w * 2 * Gm::PPW + sub) * Gm::STRIDE;
  uint32_t *sm_b = sm_a + Gm::PPW * Gm::STRIDE;
  const unsigned long long batch = P.batch;
  unsigned *done1 = Q.ctl + 1, *done2 = done1 + batch, *done3 = done2 + batch;
  const unsigned long long total = (batch + 2 * FLOW_DELAY) * FLOW_TASKS;

  if (threadIdx.x == 0) s_ticket = atomicAdd(Q.ctl, 1u);
  __syncthreads();
  unsigned long long t = s_ticket;
  unsigned pend = 0;
  while (t < total) {
    __syncthreads();                                             /* everyone has read s_ticket */
    if (threadIdx.x == 0) pend = atomicAdd(Q.ctl, 1u);            /* the next ticket, a whole task ahead */
    const unsigned long long step = t / FLOW_TASKS;
    const unsigned u = (unsigned)(t % FLOW_TASKS);
    unsigned *signal = nullptr;
    if (u < 16) {                                                /* pass 1 */
      const unsigned long long poly = step;
      if (poly < batch) {
        const unsigned op = u >> 3, tile = u & 7;
        if (poly >= FLOW_SLOTS) flow_wait(done3 + (poly - FLOW_SLOTS), 8);
        uint32_t *slot = P.scratch + ((size_t)(poly % FLOW_SLOTS) << (logn + 1)) + ((size_t)op << logn);
        const size_t off = (size_t)tile * 32 + lane;
        fused_cols_fwd<K1, ARITH>((op ? P.b : P.a) + ((size_t)poly << logn) + off, slot + off, sm_col, w, P, m);
        signal = done1 + poly;
      }
    } else if (u < 32) {                                         /* pass 2 */
      if (step >= FLOW_DELAY && step - FLOW_DELAY < batch) {
        const unsigned long long poly = step - FLOW_DELAY;
        flow_wait(done1 + poly, 16);
        uint32_t *sa = P.scratch + ((size_t)(poly % FLOW_SLOTS) << (logn + 1));
        const uint32_t j = (u - 16) * (2 * FUSED_WARPS) + w * 2 + sub;
        fused_row<lr, ARITH>(sa + ((size_t)j << lr), sa + ((size_t)1 << logn) + ((size_t)j << lr), sm_a, sm_b, l,
                             (uint32_t)K1, j, P, m);
        signal = done2 + poly;
      }
    } else {                                                     /* pass 3 */
      if (step >= 2 * FLOW_DELAY && step - 2 * FLOW_DELAY < batch) {
        const unsigned long long poly = step - 2 * FLOW_DELAY;
        flow_wait(done2 + poly, 16);
        const uint32_t *sa = P.scratch + ((size_t)(poly % FLOW_SLOTS) << (logn + 1));
        const size_t off = (size_t)(u - 32) * 32 + lane;
        uint32_t x[G::NV];
        fused_cols_inv_load<K1, ARITH>(x, sa + off, w);
        fused_cols_inv<K1, ARITH>(x, P.c + ((size_t)poly << logn) + off, sm_col, w, P, m);
        signal = done3 + poly;
      }
    }
    __syncthreads();                                             /* the task's stores are issued; smem is free */
    if (threadIdx.x == 0) {
      if (signal) {
        __threadfence();                                         /* cumulative: orders the CTA's stores before the count */
        atomicAdd(signal, 1u);
      }
      s_ticket = pend;
    }
    __syncthreads();
    t = s_ticket;
  }
}

}  // namespace nttb200
