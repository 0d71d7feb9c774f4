/*
 * ntt_generic.cuh -- one-butterfly-per-thread, one-stage-per-launch kernels working in
 * place in global memory.  They execute the reference dataflows literally, for any
 * power-of-two n and any table:
 *   DF_CT_STD2REV  R/NTT/ntt.C:295-371     DF_GS_REV2STD  R/NTT/ntt.C:387-451
 *   DF_CT_REV2STD  R/NTT/ntt.C:216-278     DF_GS_STD2REV  R/NTT/ntt.C:467-525
 * Values are kept canonical [0,q) after every stage (exactly the reference's invariant),
 * so these also serve as the in-GPU cross-check of the fused kernels.  They are the
 * compatibility path (caller-supplied tables, the two dataflows the fused kernels do not
 * use) -- the fast paths are ntt_small.cuh and ntt_large.cuh.
 */
#pragma once
#include <stdint.h>
#include "modarith.cuh"

namespace nttb200 {

enum { DF_CT_STD2REV = 0, DF_GS_REV2STD = 1, DF_CT_REV2STD = 2, DF_GS_STD2REV = 3 };

/* stage parameter `half`: distance between the two legs (d or t in the reference loops).
 * Block-indexed dataflows (0,1): pair index b -> block j = b / half, s = 2*half*j + b%half,
 *   twiddle p[(n/2)/half + j].
 * Offset-indexed dataflows (2,3): pair index b -> j = b % half, s = 2*half*(b/half) + j,
 *   twiddle p[half + j]. */
/* The j = 0 butterfly of the reference's UN-MERGED entry points (ntt_ct_*, ntt_gs_*; e.g.
 * R/NTT/ntt.C:313-317, 401-405): no multiplication and p[t] is never read -- the same for the
 * CT and the GS family: (X, Y) -> (X + Y, X - Y), canonical in and out. */
__device__ __forceinline__ void plain_bfly(uint32_t &X, uint32_t &Y, const ModQ &m) {
  const uint32_t s = X + Y, d = X - Y;
  X = csub(s, m.q);
  Y = min(d, d + m.q);
}

template <int DF>
__global__ void __launch_bounds__(256)
generic_stage_kernel(uint32_t *data, const uint2 *tab, uint32_t n, uint32_t logn, uint32_t half,
                     uint32_t loghalf, unsigned long long total_pairs, ModQ m, int skip0) {
  unsigned long long gid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned long long gstride = (unsigned long long)gridDim.x * blockDim.x;
  for (; gid < total_pairs; gid += gstride) {
    const unsigned long long poly = gid >> (logn - 1);
    const uint32_t b = (uint32_t)(gid & ((n >> 1) - 1));
    const uint32_t hi_part = b >> loghalf, lo_part = b & (half - 1);
    const uint32_t s = (hi_part << (loghalf + 1)) | lo_part;
    uint32_t tidx, j;
    if (DF == DF_CT_STD2REV || DF == DF_GS_REV2STD) { j = hi_part; tidx = ((n >> 1) >> loghalf) + hi_part; }
    else { j = lo_part; tidx = half + lo_part; }
    const uint2 w = __ldg(tab + tidx);
    uint32_t *px = data + (poly << logn) + s;
    uint32_t X = px[0], Y = px[half];
    if (skip0 && j == 0) plain_bfly(X, Y, m);
    else if (DF == DF_CT_STD2REV || DF == DF_CT_REV2STD) ct_bfly<ARITH_CANON>(X, Y, w.x, w.y, m);
    else gs_bfly<ARITH_CANON>(X, Y, w.x, w.y, m, 0u);
    px[0] = X;
    px[half] = Y;
  }
}

/* c[i] = a[i] * b[i] mod q  (mul_array, R/NTT/ntt.C:131-137); bw = b[i] has no Shoup
 * companion, so Montgomery twice: REDC(REDC(a*b) * r2) with r2 = 2^64 mod q */
__global__ void __launch_bounds__(256)
pointwise_kernel(uint32_t *c, const uint32_t *a, const uint32_t *b, unsigned long long count,
                 ModQ m, uint32_t r2) {
  unsigned long long gid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned long long gstride = (unsigned long long)gridDim.x * blockDim.x;
  for (; gid < count; gid += gstride) {
    uint32_t v = csub(mont_mul(a[gid], b[gid], m), m.q);
    c[gid] = csub(mont_mul(v, r2, m), m.q);
  }
}

/* a[i] = a[i] * w mod q for one scalar or a per-coefficient table (mul_array16 /
 * scalar_mul_array, R/NTT/ntt.C:119-125, 147-153).  tab == nullptr -> scalar `sc`. */
__global__ void __launch_bounds__(256)
scale_kernel(uint32_t *a, const uint2 *tab, uint2 sc, uint32_t n, unsigned long long count, ModQ m) {
  unsigned long long gid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned long long gstride = (unsigned long long)gridDim.x * blockDim.x;
  for (; gid < count; gid += gstride) {
    uint2 w = tab ? __ldg(tab + (gid & (n - 1))) : sc;
    a[gid] = csub(shoup_mul(a[gid], w.x, w.y, m), m.q);
  }
}

/* =====================================================================================
 * Whole literal transform in ONE launch for n <= 4096: one CTA per polynomial, the polynomial
 * in shared memory, one __syncthreads per stage.  Same per-stage arithmetic as
 * generic_stage_kernel (canonical after every stage) / generic_red_stage_kernel (exact
 * Longa-Naehrig values).  This is what the one-polynomial-per-call legacy functions run, on a
 * mapped pinned buffer, so a call costs one launch and one synchronisation.
 * ===================================================================================== */
__device__ __forceinline__ int32_t ln_mul_red_fwd(int32_t x, int32_t y);

template <int DF, bool RED>
__global__ void __launch_bounds__(512)
literal_cta_kernel(uint32_t *data, const uint2 *tab, const int32_t *rtab, uint32_t n, uint32_t logn,
                   unsigned long long batch, ModQ m, int skip0) {
  extern __shared__ __align__(16) uint32_t sx[];
  const bool descending = (DF == DF_CT_STD2REV || DF == DF_GS_STD2REV);
  /* the table may live in mapped host memory: read it once */
  uint2 *stab = reinterpret_cast<uint2 *>(sx + n);
  int32_t *srtab = reinterpret_cast<int32_t *>(sx + n);
  if (!RED) { for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) stab[i] = tab[i]; }
  else { for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) srtab[i] = rtab[i]; }
  for (unsigned long long poly = blockIdx.x; poly < batch; poly += gridDim.x) {
    uint32_t *g = data + (poly << logn);
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) sx[i] = g[i];
    __syncthreads();
    for (uint32_t s = 0; s < logn; s++) {
      const uint32_t lh = descending ? (logn - 1 - s) : s;
      const uint32_t half = 1u << lh;
      for (uint32_t b = threadIdx.x; b < (n >> 1); b += blockDim.x) {
        const uint32_t hi_part = b >> lh, lo_part = b & (half - 1);
        const uint32_t p0 = (hi_part << (lh + 1)) | lo_part;
        uint32_t tidx, j;
        if (DF == DF_CT_STD2REV || DF == DF_GS_REV2STD) { j = hi_part; tidx = ((n >> 1) >> lh) + hi_part; }
        else { j = lo_part; tidx = half + lo_part; }
        if (!RED) {
          const uint2 w = stab[tidx];
          uint32_t X = sx[p0], Y = sx[p0 + half];
          if (skip0 && j == 0) plain_bfly(X, Y, m);
          else if (DF == DF_CT_STD2REV || DF == DF_CT_REV2STD) ct_bfly<ARITH_CANON>(X, Y, w.x, w.y, m);
          else gs_bfly<ARITH_CANON>(X, Y, w.x, w.y, m, 0u);
          sx[p0] = X;
          sx[p0 + half] = Y;
        } else {
          const int32_t w = srtab[tidx];
          const bool plain = skip0 && j == 0;
          const int32_t X = (int32_t)sx[p0], Y = (int32_t)sx[p0 + half];
          if (DF == DF_CT_STD2REV || DF == DF_CT_REV2STD) {
            const int32_t x = plain ? Y : ln_mul_red_fwd(Y, w);
            sx[p0 + half] = (uint32_t)X - (uint32_t)x;
            sx[p0] = (uint32_t)X + (uint32_t)x;
          } else {
            const int32_t d = (int32_t)((uint32_t)X - (uint32_t)Y);
            sx[p0 + half] = (uint32_t)(plain ? d : ln_mul_red_fwd(d, w));
            sx[p0] = (uint32_t)X + (uint32_t)Y;
          }
        }
      }
      __syncthreads();
    }
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) g[i] = sx[i];
    __syncthreads();
  }
}

/* =====================================================================================
 * Exact emulation of the reference's Longa-Naehrig ("RED") functions, q = 12289 hard-wired
 * as in R/NTT-RED/ntt_red.c:24-46: signed 32-bit values that are NOT reduced mod q
 * (R/NTT-RED/ntt_red256.h:18), red(x) = 3 (x & 4095) - (x >> 12), tables pre-scaled by 1/3.
 * Bit-exact outputs require the identical butterfly network and the identical formula, so
 * these are literal one-stage-per-launch kernels (compatibility surface, not a fast path).
 * ===================================================================================== */
__device__ __forceinline__ int32_t ln_red(int32_t x) { return 3 * (x & 4095) - (x >> 12); }   /* ntt_red.c:34-36 */
__device__ __forceinline__ int32_t ln_mul_red(int32_t x, int32_t y) {                         /* ntt_red.c:39-46 */
  const long long z = (long long)x * y;
  const uint32_t lo = (uint32_t)(z & 4095);
  const uint32_t hi = (uint32_t)(z >> 12);         /* the reference truncates z >> 12 to int32 */
  return (int32_t)(3u * lo - hi);
}
__device__ __forceinline__ int32_t ln_mul_red_fwd(int32_t x, int32_t y) { return ln_mul_red(x, y); }

/* skip0: the un-merged entry points (ntt_red_ct_*, ntt_red_gs_*) do the j = 0 butterflies
 * without a multiplication (e.g. ntt_red.c:339-343); the psi-merged ones multiply every j. */
template <int DF>
__global__ void __launch_bounds__(256)
generic_red_stage_kernel(int32_t *data, const int32_t *tab, uint32_t n, uint32_t logn, uint32_t half,
                         uint32_t loghalf, unsigned long long total_pairs, int skip0) {
  unsigned long long gid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned long long gstride = (unsigned long long)gridDim.x * blockDim.x;
  for (; gid < total_pairs; gid += gstride) {
    const unsigned long long poly = gid >> (logn - 1);
    const uint32_t b = (uint32_t)(gid & ((n >> 1) - 1));
    const uint32_t hi_part = b >> loghalf, lo_part = b & (half - 1);
    const uint32_t s = (hi_part << (loghalf + 1)) | lo_part;
    uint32_t tidx, j;
    if (DF == DF_CT_STD2REV || DF == DF_GS_REV2STD) { j = hi_part; tidx = ((n >> 1) >> loghalf) + hi_part; }
    else { j = lo_part; tidx = half + lo_part; }
    const int32_t w = tab[tidx];
    const bool plain = skip0 && j == 0;
    int32_t *px = data + (poly << logn) + s;
    const int32_t X = px[0], Y = px[half];
    if (DF == DF_CT_STD2REV || DF == DF_CT_REV2STD) {
      const int32_t x = plain ? Y : ln_mul_red(Y, w);
      px[half] = (int32_t)((uint32_t)X - (uint32_t)x);
      px[0] = (int32_t)((uint32_t)X + (uint32_t)x);
    } else {
      const int32_t d = (int32_t)((uint32_t)X - (uint32_t)Y);
      px[half] = plain ? d : ln_mul_red(d, w);
      px[0] = (int32_t)((uint32_t)X + (uint32_t)Y);
    }
  }
}

enum {
  RED_OP_NORMALIZE = 0,        /* ntt_red.c:72-82    */
  RED_OP_NORMALIZE_INV3 = 1,   /* ntt_red.c:87-97    */
  RED_OP_SHIFT = 2,            /* ntt_red.c:103-111  */
  RED_OP_REDUCE = 3,           /* ntt_red.c:124-130  */
  RED_OP_REDUCE_TWICE = 4,     /* ntt_red.c:138-144  */
  RED_OP_CORRECT = 5,          /* ntt_red.c:150-169  */
  RED_OP_MUL_RED = 6,          /* ntt_red.c:197-211: c[i] = mul_red(a[i], b[i])  */
  RED_OP_SCALAR_MUL_RED = 7    /* ntt_red.c:217-223  */
};

__global__ void __launch_bounds__(256)
red_elementwise_kernel(int op, int32_t *c, const int32_t *a, const int32_t *b, int32_t sc,
                       unsigned long long count) {
  const int32_t Q = 12289;
  unsigned long long gid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned long long gstride = (unsigned long long)gridDim.x * blockDim.x;
  for (; gid < count; gid += gstride) {
    int32_t x = a[gid];
    switch (op) {
      case RED_OP_NORMALIZE: x = x % Q; if (x < 0) x += Q; break;
      case RED_OP_NORMALIZE_INV3: x = (int32_t)(((long long)x * 8193) % Q); if (x < 0) x += Q; break;
      case RED_OP_SHIFT: x = (x > (Q - 1) / 2) ? x - Q : x; break;
      case RED_OP_REDUCE: x = ln_red(x); break;
      case RED_OP_REDUCE_TWICE: x = ln_red(ln_red(x)); break;
      case RED_OP_CORRECT: x += ((x >> 16) & Q); x -= Q; x += ((x >> 16) & Q); break;
      case RED_OP_MUL_RED: x = ln_mul_red(x, b[gid]); break;
      case RED_OP_SCALAR_MUL_RED: x = ln_mul_red(x, sc); break;
      default: break;
    }
    c[gid] = x;
  }
}

/* in-place bit-reversal permutation of every row (bitrev_shuffle, R/NTT/ntt.C:27-44) */
__global__ void __launch_bounds__(256)
bitrev_shuffle_kernel(uint32_t *a, uint32_t logn, unsigned long long total) {
  unsigned long long gid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned long long gstride = (unsigned long long)gridDim.x * blockDim.x;
  for (; gid < total; gid += gstride) {
    const uint32_t i = (uint32_t)(gid & ((1u << logn) - 1));
    const uint32_t j = __brev(i) >> (32 - logn);
    if (i < j) {
      uint32_t *row = a + (gid - i);
      const uint32_t x = row[i];
      row[i] = row[j];
      row[j] = x;
    }
  }
}
/* shuffle_with_table (R/NTT/ntt.C:50-59): the caller's swap list, applied in order */
__global__ void shuffle_table_kernel(uint32_t *a, const uint16_t *pairs, uint32_t npairs) {
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (uint32_t i = 0; i < npairs; i++) {
      const uint32_t j = pairs[2 * i], k = pairs[2 * i + 1];
      const uint32_t x = a[j];
      a[j] = a[k];
      a[k] = x;
    }
}

}  // namespace nttb200
