/* multi-pass large-n kernels, arithmetic class ARITH_CANON (see modarith.cuh) */
#define LARGE_ARITH ARITH_CANON
#define LARGE_NAME canon
#include "large_dispatch.inl"
