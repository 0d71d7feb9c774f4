/* n <= 1024 kernels, arithmetic class ARITH_CANON (see modarith.cuh) */
#define SMALL_ARITH ARITH_CANON
#define SMALL_NAME canon
#include "small_dispatch.inl"
