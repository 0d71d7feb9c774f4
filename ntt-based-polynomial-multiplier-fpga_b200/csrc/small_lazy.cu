/* n <= 1024 kernels, arithmetic class ARITH_LAZY (see modarith.cuh) */
#define SMALL_ARITH ARITH_LAZY
#define SMALL_NAME lazy
#include "small_dispatch.inl"
