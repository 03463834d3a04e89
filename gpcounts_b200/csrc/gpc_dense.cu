// Second instantiation of the persistent kernels: dense evaluators only, two CTAs per SM (see kernels.cuh).
#define GPC_DENSE_ONLY 1
#define GPC_MIN_BLOCKS 2
#define GPC_VARIANT d128
#include "kernels.cuh"
