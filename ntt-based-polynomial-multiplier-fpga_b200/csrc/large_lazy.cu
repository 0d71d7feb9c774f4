/* multi-pass large-n kernels, arithmetic class ARITH_LAZY (see modarith.cuh) */
#define LARGE_ARITH ARITH_LAZY
#define LARGE_NAME lazy
#include "large_dispatch.inl"
