/* n <= 1024 kernels, arithmetic class ARITH_HARVEY (see modarith.cuh) */
#define SMALL_ARITH ARITH_HARVEY
#define SMALL_NAME harvey
#include "small_dispatch.inl"
