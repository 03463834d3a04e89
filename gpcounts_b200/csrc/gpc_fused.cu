// Instantiation of the persistent kernels with the fused on-chip SVGP evaluator only (see kernels.cuh).
#define GPC_FUSED_ONLY 1
#define GPC_VARIANT f256
#include "kernels.cuh"
