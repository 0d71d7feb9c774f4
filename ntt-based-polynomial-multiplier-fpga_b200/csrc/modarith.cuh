/*
 * modarith.cuh -- integer-pipe modular arithmetic for the NTT kernels (sm_100a).
 *
 * Everything here is 32-bit IMAD / IMAD.HI / IADD3 / IMNMX work: no tensor cores, no
 * floating point.  The reference reduces with a q=12289-specific Barrett
 * (R/NTT/ntt.C:101-107, (x*178942409)>>41) or Longa-Naehrig k-red
 * (R/NTT-RED/ntt_red.c:34-46); on the GPU every multiplication by a *known* twiddle is a
 * Shoup multiplication (3 IMAD-class issues, result in [0,2q)) and the pointwise product
 * is one Montgomery REDC whose 2^-32 is cancelled by the final n^-1 constant.  All
 * variants give the same canonical residues, which is what the reference outputs.
 *
 * Arithmetic classes (picked per plan from q and log2 n):
 *   ARITH_LAZY   no correction inside the transforms: CT values grow by 2q per stage,
 *                GS values double per stage; needs q * 2^(logn+1) < 2^32 (12289, 7681, 3329 ...)
 *   ARITH_HARVEY q < 2^30: Harvey butterflies, CT values in [0,4q), GS values in [0,2q)
 *   ARITH_CANON  q < 2^31: every butterfly output canonical [0,q)
 */
#pragma once
#include <stdint.h>

enum { ARITH_LAZY = 0, ARITH_HARVEY = 1, ARITH_CANON = 2 };

struct ModQ {
  uint32_t q;      /* modulus                                  */
  uint32_t nq;     /* 2^32 - q  (so  x + t*nq == x - t*q)      */
  uint32_t q2;     /* 2q (LAZY/HARVEY only; wraps for q>=2^31) */
  uint32_t qinv;   /* q^-1 mod 2^32 (Montgomery)               */
};

/* The modulus constants arrive as kernel parameters, which ptxas keeps in UNIFORM registers
 * (LDCU -> URx operands).  Measured on B200 with the Plantard kernel, the
 * butterflies run faster when the constants they read on every instruction sit in ordinary
 * registers: adding a parameter that is always 0 makes the copy one that ptxas cannot fold
 * back into a parameter read.  NTT_MODQ_REGS=0 turns it off (tuning knob). */
#ifndef NTT_MODQ_REGS
#define NTT_MODQ_REGS 1
#endif
__device__ __forceinline__ ModQ modq_regs(const ModQ &m, uint32_t zero) {
#if NTT_MODQ_REGS
  ModQ r;
  r.q = m.q + zero;
  r.nq = m.nq + zero;
  r.q2 = m.q2 + zero;
  r.qinv = m.qinv + zero;
  return r;
#else
  (void)zero;
  return m;
#endif
}

/* x * w mod q for a twiddle w with companion wp = floor(w 2^32 / q): any 32-bit x,
 * result in [0, 2q).  IMAD.HI + 2 IMAD. */
__device__ __forceinline__ uint32_t shoup_mul(uint32_t x, uint32_t w, uint32_t wp, const ModQ &m) {
  uint32_t t = __umulhi(x, wp);
  return x * w + t * m.nq;
}

/* [0,2c) -> [0,c): x >= c ? x - c : x, as IADD + IMNMX.U32 (x - c wraps high when x < c) */
__device__ __forceinline__ uint32_t csub(uint32_t x, uint32_t c) { return min(x, x - c); }

/* Montgomery product a*b*2^-32 mod q for a*b < q*2^32; result in (0, 2q) */
__device__ __forceinline__ uint32_t mont_mul(uint32_t a, uint32_t b, const ModQ &m) {
  uint32_t lo = a * b;
  uint32_t hi = __umulhi(a, b);
  uint32_t k = lo * m.qinv;
  uint32_t t = __umulhi(k, m.q);
  return hi - t + m.q;
}

/* ---------------------------------------------------------------------------------
 * Butterflies.  `bound` arguments are compile-time multiples of q tracked by the
 * callers (LAZY); HARVEY/CANON keep fixed ranges.
 * ------------------------------------------------------------------------------- */

/* Cooley-Tukey: (X, Y) -> (X + wY, X - wY)     [reference: R/NTT/ntt.C:323-326]
 * CANON class, LAZYOUT: a Shoup multiplication takes ANY 32-bit input, so the two results of a
 * butterfly need their conditional subtractions only if the NEXT stage adds to them (its X leg);
 * when both go into the next stage's multiplication (its Y leg) they may stay in [0, 2q) --
 * 2q < 2^32 for every q < 2^31 -- which saves two of the eight instructions.  The callers know
 * at compile time which butterflies those are (the next stage's index bit of their registers). */
template <int ARITH, bool LAZYOUT = false>
__device__ __forceinline__ void ct_bfly(uint32_t &X, uint32_t &Y, uint32_t w, uint32_t wp,
                                        const ModQ &m) {
  if (ARITH == ARITH_LAZY) {
    uint32_t T = shoup_mul(Y, w, wp, m);        /* [0,2q) */
    Y = X - T + m.q2;                            /* grows by 2q per stage */
    X = X + T;
  } else if (ARITH == ARITH_HARVEY) {
    X = csub(X, m.q2);                           /* [0,4q) -> [0,2q) */
    uint32_t T = shoup_mul(Y, w, wp, m);
    Y = X - T + m.q2;                            /* [0,4q) */
    X = X + T;
  } else {
    uint32_t T = csub(shoup_mul(Y, w, wp, m), m.q);   /* [0,q) */
    if (LAZYOUT) {
      Y = X - T + m.q;                                /* (0, 2q) */
      X = X + T;                                      /* [0, 2q) */
    } else {
      uint32_t s = X + T, d = X - T;
      X = csub(s, m.q);
      Y = min(d, d + m.q);
    }
  }
}

/* Gentleman-Sande: (X, Y) -> (X + Y, (X - Y) w)     [reference: R/NTT/ntt.C:408-411]
 * ybound = compile-time multiple of q that bounds Y (LAZY), added so X - Y stays >= 0 */
template <int ARITH>
__device__ __forceinline__ void gs_bfly(uint32_t &X, uint32_t &Y, uint32_t w, uint32_t wp,
                                        const ModQ &m, uint32_t ybound_q) {
  if (ARITH == ARITH_LAZY) {
    uint32_t d = X - Y + ybound_q;
    X = X + Y;
    Y = shoup_mul(d, w, wp, m);
  } else if (ARITH == ARITH_HARVEY) {
    uint32_t d = X - Y + m.q2;                   /* X,Y in [0,2q) */
    X = csub(X + Y, m.q2);
    Y = shoup_mul(d, w, wp, m);
  } else {
    uint32_t s = X + Y, d = X - Y + m.q;         /* X,Y in [0,q); d in (0, 2q): any 32-bit value may be multiplied */
    X = csub(s, m.q);
    Y = csub(shoup_mul(d, w, wp, m), m.q);
  }
}
