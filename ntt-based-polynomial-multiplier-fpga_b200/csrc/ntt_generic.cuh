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
template <int DF>
__global__ void __launch_bounds__(256)
generic_stage_kernel(uint32_t *data, const uint2 *tab, uint32_t n, uint32_t logn, uint32_t half,
                     uint32_t loghalf, unsigned long long total_pairs, ModQ m) {
  unsigned long long gid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned long long gstride = (unsigned long long)gridDim.x * blockDim.x;
  for (; gid < total_pairs; gid += gstride) {
    const unsigned long long poly = gid >> (logn - 1);
    const uint32_t b = (uint32_t)(gid & ((n >> 1) - 1));
    const uint32_t hi_part = b >> loghalf, lo_part = b & (half - 1);
    const uint32_t s = (hi_part << (loghalf + 1)) | lo_part;
    uint32_t tidx;
    if (DF == DF_CT_STD2REV || DF == DF_GS_REV2STD) tidx = ((n >> 1) >> loghalf) + hi_part;
    else tidx = half + lo_part;
    const uint2 w = __ldg(tab + tidx);
    uint32_t *px = data + (poly << logn) + s;
    uint32_t X = px[0], Y = px[half];
    if (DF == DF_CT_STD2REV || DF == DF_CT_REV2STD) ct_bfly<ARITH_CANON>(X, Y, w.x, w.y, m);
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

}  // namespace nttb200
