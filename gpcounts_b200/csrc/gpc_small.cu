// Small-N instantiation of the per-gene kernels (kernels.cuh): 64-thread CTAs, on-chip full-VGP evaluator (vgp_small.cuh),
// 8-9 gene-models per SM - the latency-bound regime of the reference's fission-yeast configuration (N = 18 / 36).
#define GPC_VARIANT s64
#define GPC_THREADS 64
#define GPC_SMALL_ONLY
#define GPC_MIN_BLOCKS 6      // shared memory allows six 37 KB CTAs per SM at N = 18: 170 registers per thread
#include "kernels.cuh"
