/* multi-pass large-n kernels, arithmetic class ARITH_HARVEY (see modarith.cuh) */
#define LARGE_ARITH ARITH_HARVEY
#define LARGE_NAME harvey
#include "large_dispatch.inl"
