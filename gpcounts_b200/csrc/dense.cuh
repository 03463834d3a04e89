// CTA-level FP64 dense linear algebra for one gene-model per thread block (sm_100a).
//
// Every routine is cooperative over the whole CTA (GPC_THREADS threads), takes row-major matrices
// through generic pointers (the operands live in the per-slot HBM/L2 workspace; tiles are staged
// through shared memory) and leaves the CTA synchronised on return.  Control flow is CTA-uniform.
//
// The work horse is cta_gemm: a 64x64x16-tiled, 4x4 register-blocked DFMA kernel with operand
// transposition, triangular operand masks (which also trim the k-range per tile) and a lower-only
// output mode.  Cholesky, the triangular solves and the Cholesky reverse pass are blocked at 64 and
// expressed through it; only the 64x64 diagonal blocks are factored / inverted in shared memory.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#ifndef GPC_THREADS
#define GPC_THREADS 256
#endif
#define GPC_BLK 64                 // block size of the blocked factorisations (= GEMM tile edge)
#define GPC_BK 16                 // GEMM k-panel
#define GPC_LDS 68                // padded leading dimension of the staged GEMM panels (== 4 mod 16: conflict-free DMMA fragments)
#define GPC_LDD 65                // padded leading dimension of the 64x64 diagonal block in smem

enum { F_A_LOWER = 1, F_A_UPPER = 2, F_B_LOWER = 4, F_B_UPPER = 8, F_C_LOWER = 16 };

// Shared-memory arena of the CTA (carved by the kernels in this library).
struct CtaSmem {
    double panelA[GPC_BK * GPC_LDS];
    double panelB[GPC_BK * GPC_LDS];
    double diag[GPC_BLK * GPC_LDD];      // diagonal block being factored / its inverse
    double red[64];                     // reduction scratch
    int flag;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum over the CTA; every thread receives the bit-identical result.
__device__ __forceinline__ double cta_sum(double v, double* red) {
    v = warp_sum(v);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < GPC_THREADS / 32; ++w) t += red[w];
    __syncthreads();
    return t;
}

template <int K>
__device__ __forceinline__ void cta_sum_n(double (&v)[K], double* red) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < K; ++i) {
        v[i] = warp_sum(v[i]);
        if (lane == 0) red[i * 8 + warp] = v[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < K; ++i) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < GPC_THREADS / 32; ++w) t += red[i * 8 + w];
        v[i] = t;
    }
    __syncthreads();
}

// C[m x n] = alpha * opA(A)[m x k] * opB(B)[k x n] + beta * C      (row-major everywhere)
//   TA = false: A stored m x k;  TA = true: A stored k x m  (opA = transpose)
//   TB = false: B stored k x n;  TB = true: B stored n x k
// flags: F_A_LOWER / F_A_UPPER  opA(A)(i,kk) is zero above / below its diagonal (stored garbage ignored)
//        F_B_LOWER / F_B_UPPER  same for opB(B)(kk,j)
//        F_C_LOWER              only entries j <= i are computed and stored
// In-place use (C aliasing an operand) is safe when the aliased operand region of every output tile is
// that tile itself (single tile along the shared dimension), because all loads of a tile precede its stores.
__device__ __forceinline__ void dense_dmma884(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// Inner product of a 64 x 64 output tile on the FP64 tensor pipe: warp w owns rows 8w..8w+7 (eight 8 x 8 DMMA tiles);
// the k-panels (16 deep) are staged k-major in shared memory and the global loads of the next panel are issued
// before the current one is multiplied (register prefetch).
template <bool TA, bool TB>
__device__ __noinline__ void cta_gemm(int m, int n, int k, double alpha, const double* __restrict__ A, int lda,
                         const double* __restrict__ B, int ldb, double beta, double* C, int ldc, int flags,
                         CtaSmem* sm) {
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int ar = lane >> 2, ak = lane & 3;
    double* As = sm->panelA;
    double* Bs = sm->panelB;
    for (int i0 = 0; i0 < m; i0 += 64) {
        for (int j0 = 0; j0 < n; j0 += 64) {
            if ((flags & F_C_LOWER) && j0 > i0 + 63) continue;
            int kmin = 0, kmax = k;
            if (flags & F_A_UPPER) kmin = max(kmin, i0);
            if (flags & F_B_LOWER) kmin = max(kmin, j0);
            if (flags & F_A_LOWER) kmax = min(kmax, i0 + 64);
            if (flags & F_B_UPPER) kmax = min(kmax, j0 + 64);
            kmin &= ~(GPC_BK - 1);
            double acc[8][2];
#pragma unroll
            for (int t = 0; t < 8; ++t) { acc[t][0] = 0.0; acc[t][1] = 0.0; }
            double ra0, ra1, ra2, ra3, rb0, rb1, rb2, rb3;
            const bool warp_rows_valid = i0 + warp * 8 < m && (!(flags & F_C_LOWER) || j0 <= i0 + warp * 8 + 7);
            const int nct = min(8, (n - j0 + 7) >> 3);
#define GPC_LOAD_A(dst, e, K0)                                                                         \
            {                                                                                          \
                const int idx = tid + (e) * GPC_THREADS;                                               \
                int ii, kk;                                                                            \
                if (!TA) { kk = idx & 15; ii = idx >> 4; } else { ii = idx & 63; kk = idx >> 6; }      \
                const int gi = i0 + ii, gk = (K0) + kk;                                                \
                double v = 0.0;                                                                        \
                if (gi < m && gk < kmax) {                                                             \
                    bool ok = true;                                                                    \
                    if ((flags & F_A_LOWER) && gk > gi) ok = false;                                    \
                    if ((flags & F_A_UPPER) && gk < gi) ok = false;                                    \
                    if (ok) v = TA ? A[(size_t)gk * lda + gi] : A[(size_t)gi * lda + gk];              \
                }                                                                                      \
                dst = v;                                                                               \
            }
#define GPC_LOAD_B(dst, e, K0)                                                                         \
            {                                                                                          \
                const int idx = tid + (e) * GPC_THREADS;                                               \
                int jj, kk;                                                                            \
                if (!TB) { jj = idx & 63; kk = idx >> 6; } else { kk = idx & 15; jj = idx >> 4; }      \
                const int gj = j0 + jj, gk = (K0) + kk;                                                \
                double v = 0.0;                                                                        \
                if (gj < n && gk < kmax) {                                                             \
                    bool ok = true;                                                                    \
                    if ((flags & F_B_LOWER) && gj > gk) ok = false;                                    \
                    if ((flags & F_B_UPPER) && gj < gk) ok = false;                                    \
                    if (ok) v = TB ? B[(size_t)gj * ldb + gk] : B[(size_t)gk * ldb + gj];              \
                }                                                                                      \
                dst = v;                                                                               \
            }
#define GPC_STORE_AB(e, va, vb)                                                                        \
            {                                                                                          \
                const int idx = tid + (e) * GPC_THREADS;                                               \
                int ii, kk, jj, kb;                                                                    \
                if (!TA) { kk = idx & 15; ii = idx >> 4; } else { ii = idx & 63; kk = idx >> 6; }      \
                if (!TB) { jj = idx & 63; kb = idx >> 6; } else { kb = idx & 15; jj = idx >> 4; }      \
                As[kk * GPC_LDS + ii] = va;                                                            \
                Bs[kb * GPC_LDS + jj] = vb;                                                            \
            }
            if (kmin < kmax) {
                GPC_LOAD_A(ra0, 0, kmin) GPC_LOAD_A(ra1, 1, kmin) GPC_LOAD_A(ra2, 2, kmin) GPC_LOAD_A(ra3, 3, kmin)
                GPC_LOAD_B(rb0, 0, kmin) GPC_LOAD_B(rb1, 1, kmin) GPC_LOAD_B(rb2, 2, kmin) GPC_LOAD_B(rb3, 3, kmin)
            }
            for (int k0 = kmin; k0 < kmax; k0 += GPC_BK) {
                GPC_STORE_AB(0, ra0, rb0) GPC_STORE_AB(1, ra1, rb1) GPC_STORE_AB(2, ra2, rb2) GPC_STORE_AB(3, ra3, rb3)
                __syncthreads();
                if (k0 + GPC_BK < kmax) {
                    const int kn = k0 + GPC_BK;
                    GPC_LOAD_A(ra0, 0, kn) GPC_LOAD_A(ra1, 1, kn) GPC_LOAD_A(ra2, 2, kn) GPC_LOAD_A(ra3, 3, kn)
                    GPC_LOAD_B(rb0, 0, kn) GPC_LOAD_B(rb1, 1, kn) GPC_LOAD_B(rb2, 2, kn) GPC_LOAD_B(rb3, 3, kn)
                }
                if (warp_rows_valid) {             // ragged edge tiles: warps / column tiles beyond m, n do no math
#pragma unroll
                    for (int kq = 0; kq < GPC_BK; kq += 4) {
                        const double a = As[(kq + ak) * GPC_LDS + warp * 8 + ar];
#pragma unroll
                        for (int t = 0; t < 8; ++t) {
                            if (t < nct) {
                                const double b = Bs[(kq + ak) * GPC_LDS + t * 8 + ar];
                                dense_dmma884(acc[t], a, b);
                            }
                        }
                    }
                }
                __syncthreads();
            }
#undef GPC_LOAD_A
#undef GPC_LOAD_B
#undef GPC_STORE_AB
            const int gi = i0 + warp * 8 + ar;
            if (gi < m) {
#pragma unroll
                for (int t = 0; t < 8; ++t) {
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const int gj = j0 + t * 8 + 2 * ak + c;
                        if (gj >= n) continue;
                        if ((flags & F_C_LOWER) && gj > gi) continue;
                        double* p = C + (size_t)gi * ldc + gj;
                        double v = alpha * acc[t][c];
                        if (beta != 0.0) v += beta * (*p);
                        *p = v;
                    }
                }
            }
        }
    }
    __syncthreads();
}

// Factor the nb x nb (nb <= 64) SPD block held in sm->diag (lower part used) in place: diag = chol.
// Returns false (CTA-uniform) on a non-positive or NaN pivot.
__device__ bool smem_chol(CtaSmem* sm, int nb) {
    double* D = sm->diag;
    const int tid = threadIdx.x;
    for (int j = 0; j < nb; ++j) {
        __syncthreads();
        const double djj = D[j * GPC_LDD + j];
        if (!(djj > 0.0)) return false;           // NaN or non-positive pivot (uniform: all read the same value)
        const double r = sqrt(djj);
        __syncthreads();
        if (tid == j) D[j * GPC_LDD + j] = r;
        if (tid > j && tid < nb) D[tid * GPC_LDD + j] /= r;
        __syncthreads();
        const int rem = nb - j - 1;
        for (int idx = tid; idx < rem * rem; idx += GPC_THREADS) {
            const int i = j + 1 + idx / rem, c = j + 1 + idx % rem;
            if (c <= i) D[i * GPC_LDD + c] = fma(-D[i * GPC_LDD + j], D[c * GPC_LDD + j], D[i * GPC_LDD + c]);
        }
    }
    __syncthreads();
    return true;
}

// inv (global, nb x nb row-major, ld = GPC_BLK) <- inverse of the lower-triangular block in sm->diag.
// Four threads per column split the substitution dot products.
__device__ void smem_tri_inverse(CtaSmem* sm, int nb, double* inv) {
    const double* D = sm->diag;
    const int tid = threadIdx.x;
    const int c = tid >> 2, q = tid & 3;
    // each thread keeps its strided share of column c in registers: entries k = q, q+4, ... (16 values)
    double xk[16];
#pragma unroll
    for (int t = 0; t < 16; ++t) xk[t] = 0.0;
    for (int i = 0; i < nb; ++i) {
        double part = 0.0;
#pragma unroll
        for (int t = 0; t < 16; ++t) {
            const int kk = q + 4 * t;
            if (kk < i) part = fma(D[i * GPC_LDD + kk], xk[t], part);
        }
        part += __shfl_xor_sync(0xffffffffu, part, 1);
        part += __shfl_xor_sync(0xffffffffu, part, 2);
        const double xi = (c < nb && i >= c) ? ((i == c ? 1.0 : 0.0) - part) / D[i * GPC_LDD + i] : 0.0;
#pragma unroll
        for (int t = 0; t < 16; ++t)
            if (q + 4 * t == i) xk[t] = xi;
        if (q == 0 && c < nb) inv[(size_t)i * GPC_BLK + c] = xi;
    }
    __syncthreads();
}

// Blocked left-looking Cholesky of the n x n matrix A (lower part referenced, overwritten by L).
// dinv receives the inverses of the 64x64 diagonal blocks of L (ceil(n/64) blocks of GPC_BLK*GPC_BLK doubles).
// Returns false on failure (CTA-uniform).
__device__ bool cta_potrf(int n, double* A, int lda, double* dinv, CtaSmem* sm) {
    const int tid = threadIdx.x;
    for (int j0 = 0, blk = 0; j0 < n; j0 += GPC_BLK, ++blk) {
        const int jb = min(GPC_BLK, n - j0);
        // panel update: A[j0:n, j0:j0+jb] -= L[j0:n, 0:j0] * L[j0:j0+jb, 0:j0]^T
        if (j0 > 0)
            cta_gemm<false, true>(n - j0, jb, j0, -1.0, A + (size_t)j0 * lda, lda, A + (size_t)j0 * lda, lda, 1.0,
                                  A + (size_t)j0 * lda + j0, lda, 0, sm);
        // diagonal block -> smem, factor, write back, invert
        for (int idx = tid; idx < jb * jb; idx += GPC_THREADS) {
            const int i = idx / jb, c = idx % jb;
            sm->diag[i * GPC_LDD + c] = (c <= i) ? A[(size_t)(j0 + i) * lda + j0 + c] : 0.0;
        }
        if (!smem_chol(sm, jb)) return false;
        for (int idx = tid; idx < jb * jb; idx += GPC_THREADS) {
            const int i = idx / jb, c = idx % jb;
            if (c <= i) A[(size_t)(j0 + i) * lda + j0 + c] = sm->diag[i * GPC_LDD + c];
        }
        double* inv = dinv + (size_t)blk * GPC_BLK * GPC_BLK;
        smem_tri_inverse(sm, jb, inv);
        // panel solve: L21 = A21 * L11^-T   (in place; single column tile)
        const int mrem = n - j0 - jb;
        if (mrem > 0)
            cta_gemm<false, true>(mrem, jb, jb, 1.0, A + (size_t)(j0 + jb) * lda + j0, lda, inv, GPC_BLK, 0.0,
                                  A + (size_t)(j0 + jb) * lda + j0, lda, F_B_UPPER, sm);
    }
    return true;
}

// B (n x nrhs) <- L^-1 B
__device__ void cta_trsm_left(int n, int nrhs, const double* L, int ldl, const double* dinv, double* B, int ldb,
                              CtaSmem* sm) {
    for (int i0 = 0, blk = 0; i0 < n; i0 += GPC_BLK, ++blk) {
        const int ib = min(GPC_BLK, n - i0);
        if (i0 > 0)
            cta_gemm<false, false>(ib, nrhs, i0, -1.0, L + (size_t)i0 * ldl, ldl, B, ldb, 1.0, B + (size_t)i0 * ldb,
                                   ldb, 0, sm);
        cta_gemm<false, false>(ib, nrhs, ib, 1.0, dinv + (size_t)blk * GPC_BLK * GPC_BLK, GPC_BLK, B + (size_t)i0 * ldb,
                               ldb, 0.0, B + (size_t)i0 * ldb, ldb, F_A_LOWER, sm);
    }
}

// B (n x nrhs) <- L^-T B
__device__ void cta_trsm_left_t(int n, int nrhs, const double* L, int ldl, const double* dinv, double* B, int ldb,
                                CtaSmem* sm) {
    const int nblk = (n + GPC_BLK - 1) / GPC_BLK;
    for (int blk = nblk - 1; blk >= 0; --blk) {
        const int i0 = blk * GPC_BLK, ib = min(GPC_BLK, n - i0), i1 = i0 + ib;
        if (i1 < n)   // B_i -= L[i1:n, i0:i1]^T * X[i1:n, :]
            cta_gemm<true, false>(ib, nrhs, n - i1, -1.0, L + (size_t)i1 * ldl + i0, ldl, B + (size_t)i1 * ldb, ldb,
                                  1.0, B + (size_t)i0 * ldb, ldb, 0, sm);
        cta_gemm<true, false>(ib, nrhs, ib, 1.0, dinv + (size_t)blk * GPC_BLK * GPC_BLK, GPC_BLK, B + (size_t)i0 * ldb,
                              ldb, 0.0, B + (size_t)i0 * ldb, ldb, F_A_UPPER, sm);
    }
}

// B (mrows x n) <- B L^-1
__device__ void cta_trsm_right(int mrows, int n, const double* L, int ldl, const double* dinv, double* B, int ldb,
                               CtaSmem* sm) {
    const int nblk = (n + GPC_BLK - 1) / GPC_BLK;
    for (int blk = nblk - 1; blk >= 0; --blk) {
        const int j0 = blk * GPC_BLK, jb = min(GPC_BLK, n - j0), j1 = j0 + jb;
        if (j1 < n)   // B_j -= X[:, j1:n] * L[j1:n, j0:j1]
            cta_gemm<false, false>(mrows, jb, n - j1, -1.0, B + j1, ldb, L + (size_t)j1 * ldl + j0, ldl, 1.0, B + j0,
                                   ldb, 0, sm);
        cta_gemm<false, false>(mrows, jb, jb, 1.0, B + j0, ldb, dinv + (size_t)blk * GPC_BLK * GPC_BLK, GPC_BLK, 0.0,
                               B + j0, ldb, F_B_LOWER, sm);
    }
}

// Reverse pass of the Cholesky factorisation (unsymmetrised):
//   G (n x n) holds Lbar in its lower triangle on entry; on exit G = L^-T Phi(L^T Lbar) L^-1 with
//   Phi = lower triangle with halved diagonal.  <sym(G), dK> = <G, dK> for symmetric dK.
// T is an n x n scratch matrix.
__device__ void cta_chol_backward(int n, const double* L, int ldl, const double* dinv, double* G, int ldg, double* T,
                                  int ldt, CtaSmem* sm) {
    // T = tril(L^T * Lbar)
    cta_gemm<true, false>(n, n, n, 1.0, L, ldl, G, ldg, 0.0, T, ldt, F_A_UPPER | F_B_LOWER | F_C_LOWER, sm);
    for (int idx = threadIdx.x; idx < n * n; idx += GPC_THREADS) {
        const int i = idx / n, j = idx % n;
        const double v = T[(size_t)i * ldt + j];
        G[(size_t)i * ldg + j] = (j < i) ? v : (j == i ? 0.5 * v : 0.0);
    }
    __syncthreads();
    cta_trsm_left_t(n, n, L, ldl, dinv, G, ldg, sm);
    cta_trsm_right(n, n, L, ldl, dinv, G, ldg, sm);
}
