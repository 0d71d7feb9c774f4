/*
 * microbench.cu -- integer-pipe peak measurement (nttb200_measure_int_peak).
 *
 * MEASURED_PEAKS.json holds only HBM and bf16 numbers; the INT32 roofline denominator for
 * the NTT kernels ("IMAD_peak must be measured") comes from these loops:
 * 8 independent dependency chains per thread, inline PTX so nothing is folded away.
 */
#include <cuda_runtime.h>
#include <stdint.h>

#include "modarith.cuh"
#include "plan.h"

namespace {

constexpr int CHAINS = 8;
constexpr int INNER = 64;

template <int WHICH>
__global__ void __launch_bounds__(256) int_peak_kernel(uint32_t *sink, uint32_t a, uint32_t b, int iters) {
  uint32_t x[CHAINS], y[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; i++) {
    x[i] = threadIdx.x * 2654435761u + i * 40503u + blockIdx.x;
    y[i] = x[i] ^ 0x9E3779B9u;
  }
  ModQ m;
  m.q = 12289u; m.nq = 0u - 12289u; m.q2 = 2u * 12289u; m.qinv = a;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < INNER; r++) {
#pragma unroll
      for (int i = 0; i < CHAINS; i++) {
        if (WHICH == 0) {
          asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
        } else if (WHICH == 1) {
          asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
        } else if (WHICH == 2) {
          asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(a));
        } else if (WHICH == 3) {
          /* the LAZY Cooley-Tukey butterfly: IMAD.HI + 2 IMAD + 2 IADD3 */
          uint32_t t, T;
          asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(t) : "r"(y[i]), "r"(b));
          asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(T) : "r"(y[i]), "r"(a));
          asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(T) : "r"(t), "r"(m.nq));
          asm volatile("sub.u32 %0, %1, %2;" : "=r"(y[i]) : "r"(x[i]), "r"(T));
          asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(T));
        } else if (WHICH == 4) {
          unsigned long long w;
          asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w) : "r"(x[i]), "r"(a));
          x[i] = (uint32_t)(w >> 32) ^ (uint32_t)w;
        } else if (WHICH == 5) {
          /* 1 IMAD : 1 IADD -- does the pair dual-issue across the fma and alu pipes? */
          asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
          asm volatile("add.u32 %0, %0, %1;" : "+r"(y[i]) : "r"(a));
        } else if (WHICH == 6) {
          asm volatile("min.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(y[i]));
          asm volatile("add.u32 %0, %0, %1;" : "+r"(y[i]) : "r"(a));
        } else if (WHICH == 7) {
          asm volatile("shfl.sync.bfly.b32 %0, %0, 1, 0x1f, 0xffffffff;" : "+r"(x[i]));
        } else if (WHICH == 8) {
          /* IMAD.WIDE with a 64-bit addend, result high word feeds the next one */
          unsigned long long w, c;
          asm volatile("mov.b64 %0, {%1, %2};" : "=l"(c) : "r"(b), "r"(y[i]));
          asm volatile("mad.wide.s32 %0, %1, %2, %3;" : "=l"(w) : "r"(x[i]), "r"(a), "l"(c));
          x[i] = (uint32_t)(w >> 32);
        } else if (WHICH == 9) {
          /* Plantard butterfly: p = Y*w~ ; T = hi(p*q + 2^31) ; X' = X+T ; Y' = X-T */
          uint32_t p;
          unsigned long long w;
          asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(p) : "r"(y[i]), "r"(a));
          unsigned long long half;
          asm volatile("mov.b64 %0, {%1, %2};" : "=l"(half) : "r"(b), "r"(0));
          asm volatile("mad.wide.s32 %0, %1, %2, %3;" : "=l"(w) : "r"(p), "r"(m.q), "l"(half));
          uint32_t T = (uint32_t)(w >> 32);
          asm volatile("sub.u32 %0, %1, %2;" : "=r"(y[i]) : "r"(x[i]), "r"(T));
          asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(T));
        } else if (WHICH == 10) {
          /* signed Shoup with the add folded: s = Y*w + X ; t = hi(Y*w') ; X' = t*(-q) + s ; Y' = 2X - X' */
          uint32_t s2, t, xn;
          asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(s2) : "r"(y[i]), "r"(a), "r"(x[i]));
          asm volatile("mul.hi.s32 %0, %1, %2;" : "=r"(t) : "r"(y[i]), "r"(b));
          asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(xn) : "r"(t), "r"(m.nq), "r"(s2));
          asm volatile("{ .reg .u32 tt; add.u32 tt, %1, %1; sub.u32 %0, tt, %2; }" : "=r"(y[i]) : "r"(x[i]), "r"(xn));
          x[i] = xn;
        } else if (WHICH == 11) {
          /* Plantard butterfly, half-word form: 2 IMAD + 2 arithmetic shifts + 2 adds */
          uint32_t p, u;
          asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(p) : "r"(y[i]), "r"(a));
          asm volatile("shr.s32 %0, %0, 16;" : "+r"(p));
          asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(u) : "r"(p), "r"(m.q), "r"(m.q2));
          asm volatile("shr.s32 %0, %0, 16;" : "+r"(u));
          asm volatile("sub.u32 %0, %1, %2;" : "=r"(y[i]) : "r"(x[i]), "r"(u));
          asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(u));
        } else if (WHICH == 13) {
          /* unsigned Plantard butterfly (q < 2^16): p = Y*w~ ; T = hi(p*q) in [0,q) ; Y' = X-T+q ; X' = X+T */
          uint32_t p, T;
          asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(p) : "r"(y[i]), "r"(a));
          asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(T) : "r"(p), "r"(m.q));
          asm volatile("{ .reg .u32 tt; sub.u32 tt, %1, %2; add.u32 %0, tt, %3; }" : "=r"(y[i]) : "r"(x[i]), "r"(T), "r"(m.q));
          asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(T));
        } else if (WHICH == 12) {
          asm volatile("shr.s32 %0, %0, 1;" : "+r"(x[i]));
          asm volatile("add.u32 %0, %0, %1;" : "+r"(y[i]) : "r"(a));
        } else if (WHICH == 14 || (WHICH == 17 && (i & 1))) {
          /* signed half-word Plantard butterfly with the last shift folded into the add:
           * IMAD, SHF.R.S32, IMAD, LEA.HI.SX32 (X' = X + (u >> 16)), IADD3 (Y' = 2X - X') */
          const int p = (int)(y[i] * a);
          const int u = (p >> 16) * (int)m.q + (int)b;
          const int xn = (int)x[i] + (u >> 16);
          y[i] = 2u * x[i] - (uint32_t)xn;
          x[i] = (uint32_t)xn;
        } else if (WHICH == 16 || WHICH == 17) {
          /* the same with a full-word second product: IMAD, IMAD.HI (q << 16), LEA.HI.SX32, IADD3 */
          const int p = (int)(y[i] * a);
          const int u = __mulhi(p, (int)m.q2 << 15) + (int)b;
          const int xn = (int)x[i] + (u >> 16);
          y[i] = 2u * x[i] - (uint32_t)xn;
          x[i] = (uint32_t)xn;
        } else if (WHICH == 18 || (WHICH == 19 && (i & 1)) || (WHICH == 23 && (i % 3) == 0)) {
          /* 14 with the second leg on the MULTIPLIER pipe: Y' = 2 X - X' as IMAD (X, 2, -X'), so that the
           * butterfly is 3 multiplier-pipe + 2 ALU instructions (14: 2 + 3); 19 alternates the two forms
           * (2.5 + 2.5: both pipes level), 23 uses this form in one butterfly of three */
          const int p = (int)(y[i] * a);
          const int u = (p >> 16) * (int)m.q + (int)b;
          const int xn = (int)x[i] + (u >> 16);
          int yn;
          asm("{ .reg .s32 t; neg.s32 t, %2; mad.lo.s32 %0, %1, 2, t; }" : "=r"(yn) : "r"((int)x[i]), "r"(xn));
          y[i] = (uint32_t)yn;
          x[i] = (uint32_t)xn;
        } else if (WHICH == 19 || WHICH == 23) {
          const int p = (int)(y[i] * a);
          const int u = (p >> 16) * (int)m.q + (int)b;
          const int xn = (int)x[i] + (u >> 16);
          y[i] = 2u * x[i] - (uint32_t)xn;
          x[i] = (uint32_t)xn;
        } else if (WHICH == 20) {
          /* signed half-word Gentleman-Sande butterfly: IADD, IADD, IMAD, SHF, IMAD, SHF */
          const int d = (int)x[i] - (int)y[i];
          x[i] = x[i] + y[i];
          const int p = (int)((uint32_t)d * a);
          const int u = (p >> 16) * (int)m.q + (int)b;
          y[i] = (uint32_t)(u >> 16);
        } else if (WHICH == 21) {
          /* 20 with the sum on the multiplier pipe (IMAD X, 1, Y): 3 + 3 */
          const int d = (int)x[i] - (int)y[i];
          int s;
          asm("mad.lo.s32 %0, %1, 1, %2;" : "=r"(s) : "r"((int)x[i]), "r"((int)y[i]));
          x[i] = (uint32_t)s;
          const int p = (int)((uint32_t)d * a);
          const int u = (p >> 16) * (int)m.q + (int)b;
          y[i] = (uint32_t)(u >> 16);
        } else if (WHICH == 22) {
          /* Gentleman-Sande on a product leg whose last shift is still pending (y holds u, the value is
           * u >> 16): S = X + (u >> 16) is LEA.HI.SX32, D = 2 X - S; then the product of D, left pending
           * again: LEA.HI, IADD3, IMAD, SHF, IMAD = 5 */
          const int s = (int)x[i] + ((int)y[i] >> 16);
          const int d = 2 * (int)x[i] - s;
          x[i] = (uint32_t)s;
          const int p = (int)((uint32_t)d * a);
          y[i] = (uint32_t)((p >> 16) * (int)m.q + (int)b);
        } else if (WHICH == 24 || WHICH == 25 || WHICH == 26) {
          /* the butterflies of the 31-bit class (q = 2013265921 > 2^30: CANON, modarith.cuh) as the large-n
           * kernels run them: 24 Cooley-Tukey with canonical results (8 instructions), 25 the same with both
           * results left in [0, 2q) for the next stage's multiplication (6), 26 Gentleman-Sande (7) */
          ModQ c;
          c.q = 2013265921u; c.nq = 0u - 2013265921u; c.q2 = 0; c.qinv = 0;
          if (WHICH == 24) ct_bfly<ARITH_CANON, false>(x[i], y[i], a, b, c);
          else if (WHICH == 25) ct_bfly<ARITH_CANON, true>(x[i], y[i], a, b, c);
          else gs_bfly<ARITH_CANON>(x[i], y[i], a, b, c, 0);
        } else if (WHICH == 15) {
          /* LEA.HI.SX32 alone: x += y >> 16 */
          x[i] = (uint32_t)((int)x[i] + ((int)y[i] >> 16));
          asm volatile("" : "+r"(x[i]));
        }
      }
    }
  }
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; i++) acc ^= x[i] ^ y[i];
  if (acc == 0x12345u) sink[0] = acc;
}

template <int WHICH>
int run(double ops_per_inner, double *rate) {
  int dev = 0, sms = 0;
  NTT_CUDA(cudaGetDevice(&dev));
  NTT_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  uint32_t *sink = nullptr;
  NTT_CUDA(cudaMalloc(&sink, 64));
  const int grid = sms * 8, threads = 256, iters = 256;
  cudaEvent_t e0, e1;
  NTT_CUDA(cudaEventCreate(&e0));
  NTT_CUDA(cudaEventCreate(&e1));
  double best = 0;
  for (int rep = 0; rep < 4; rep++) {
    NTT_CUDA(cudaEventRecord(e0));
    int_peak_kernel<WHICH><<<grid, threads>>>(sink, 0x9E3779B1u, 0x85EBCA77u, iters);
    NTT_CUDA(cudaEventRecord(e1));
    NTT_CUDA(cudaEventSynchronize(e1));
    NTT_CUDA(cudaGetLastError());
    float ms = 0;
    NTT_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    double ops = (double)grid * threads * iters * INNER * CHAINS * ops_per_inner;
    double r = ops / (ms * 1e-3);
    if (rep > 0 && r > best) best = r;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  *rate = best;
  return 0;
}
}  // namespace

extern "C" int nttb200_measure_int_peak(int which, double *lane_ops_per_s) {
  if (!lane_ops_per_s) return nttb200_fail(NTTB200_EPARAM, "NULL argument");
  switch (which) {
    case 0: return run<0>(1, lane_ops_per_s);   /* IMAD                         */
    case 1: return run<1>(1, lane_ops_per_s);   /* IMAD.HI                      */
    case 2: return run<2>(1, lane_ops_per_s);   /* IADD                         */
    case 3: return run<3>(1, lane_ops_per_s);   /* butterflies / s (5 instr each) */
    case 4: return run<4>(1, lane_ops_per_s);   /* IMAD.WIDE (+1 LOP)           */
    case 5: return run<5>(2, lane_ops_per_s);   /* IMAD + IADD pairs, counted as 2 */
    case 6: return run<6>(2, lane_ops_per_s);   /* IMNMX + IADD                 */
    case 7: return run<7>(1, lane_ops_per_s);   /* SHFL                         */
    case 8: return run<8>(1, lane_ops_per_s);   /* IMAD.WIDE.S32 with 64-bit addend */
    case 9: return run<9>(1, lane_ops_per_s);   /* Plantard butterflies / s (IMAD + IMAD.WIDE + 2 IADD) */
    case 10: return run<10>(1, lane_ops_per_s); /* signed Shoup butterflies / s, add folded        */
    case 11: return run<11>(1, lane_ops_per_s); /* half-word Plantard butterflies / s              */
    case 13: return run<13>(1, lane_ops_per_s); /* unsigned Plantard butterflies / s (IMAD + IMAD.HI + 2 IADD3) */
    case 12: return run<12>(2, lane_ops_per_s); /* SHF + IADD pairs, counted as 2                  */
    case 14: return run<14>(1, lane_ops_per_s); /* signed half-word Plantard butterflies / s, 5 instructions (LEA.HI.SX32) */
    case 15: return run<15>(1, lane_ops_per_s); /* LEA.HI.SX32                                     */
    case 16: return run<16>(1, lane_ops_per_s); /* signed Plantard butterflies / s: IMAD, IMAD.HI, LEA.HI.SX32, IADD3 */
    case 17: return run<17>(1, lane_ops_per_s); /* 14 and 16 alternating                           */
    case 18: return run<18>(1, lane_ops_per_s); /* 14 with Y' = 2X - X' as an IMAD (3 multiplier + 2 ALU)         */
    case 19: return run<19>(1, lane_ops_per_s); /* 14 and 18 alternating (2.5 + 2.5)                              */
    case 20: return run<20>(1, lane_ops_per_s); /* signed half-word GS butterfly, 6 instructions                  */
    case 21: return run<21>(1, lane_ops_per_s); /* 20 with the sum as an IMAD (3 + 3)                             */
    case 22: return run<22>(1, lane_ops_per_s); /* signed GS with the product's last shift left pending, 5        */
    case 23: return run<23>(1, lane_ops_per_s); /* 14, 14, 18 in turn                                             */
    case 24: return run<24>(1, lane_ops_per_s); /* 31-bit class: Cooley-Tukey butterflies / s, canonical results  */
    case 25: return run<25>(1, lane_ops_per_s); /* 31-bit class: Cooley-Tukey, results left in [0, 2q)            */
    case 26: return run<26>(1, lane_ops_per_s); /* 31-bit class: Gentleman-Sande butterflies / s                  */
    default: return nttb200_fail(NTTB200_EPARAM, "unknown microbenchmark %d", which);
  }
}
