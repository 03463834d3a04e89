// CTA-level FP64 dense linear algebra for one gene-model per thread block (sm_100a).
//
// Every routine is cooperative over the whole CTA (GPC_THREADS threads), takes row-major matrices
// through generic pointers (the operands live in the per-slot HBM/L2 workspace; tiles are staged
// through shared memory) and leaves the CTA synchronised on return.  Control flow is CTA-uniform.
//
// The work horse is cta_gemm: a 64x64x16-tiled DMMA kernel fed by a cp.async panel ring, with operand
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
#define GPC_LDR 20                // leading dimension of a row-major staged panel (64 rows x 16 k; == 4 mod 16 too)
#define GPC_PANEL (64 * GPC_LDR)  // doubles per staged operand panel (>= GPC_BK * GPC_LDS)
#ifndef GPC_STAGES
#define GPC_STAGES 3              // depth of the cp.async panel ring
#endif
#define GPC_LDD 65                // padded leading dimension of the 64x64 diagonal block in smem

// Visit every (i, j) of an R x C row-major array: warp w takes rows w, w + 8, ..., the lanes stride the columns
// (coalesced); idx = i * C + j.  No integer division, unlike a flat index loop.
#define GPC_FOR_2D(R, C, i, j)                                                  \
    for (int i = threadIdx.x >> 5; i < (R); i += GPC_THREADS / 32)              \
        for (int j = threadIdx.x & 31, idx = i * (C) + j; j < (C); j += 32, idx += 32)

enum { F_A_LOWER = 1, F_A_UPPER = 2, F_B_LOWER = 4, F_B_UPPER = 8, F_C_LOWER = 16 };

extern __shared__ __align__(128) unsigned char g_smem[];   // 128 B: TMA destinations live in it

__device__ __forceinline__ struct CtaSmem* cta_arena();

// Shared-memory arena of the CTA (carved by the kernels in this library at offset 0 of the dynamic shared memory).
struct CtaSmem {
    double panel[GPC_STAGES][2][GPC_PANEL];   // cp.async ring of (A, B) k-panels
    double diag[GPC_BLK * GPC_LDD];      // diagonal block being factored / its inverse
    double red[64];                     // reduction scratch
    int flag;
};

// The arena through the shared window (see model.cuh: evaluators are reached through a function pointer, so a
// `CtaSmem*` argument is a generic pointer to the compiler; this one is known to be shared memory).
__device__ __forceinline__ CtaSmem* cta_arena() { return reinterpret_cast<CtaSmem*>(g_smem); }

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

// 8-byte asynchronous global -> shared copy; `ok == false` writes a zero without touching global memory.
__device__ __forceinline__ void cp_async8(double* dst, const double* src, bool ok) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    const int nbytes = ok ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" :: "r"(d), "l"(src), "r"(nbytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" :: "n"(N) : "memory"); }

// Inner product of a 64 x 64 output tile on the FP64 tensor pipe: warp w owns rows 8w..8w+7 (eight 8 x 8 DMMA tiles).
// The 16-deep k-panels travel global -> shared with cp.async through a GPC_STAGES-deep ring (the operands of the
// dense path live in the slot's HBM workspace, far larger than L2 over all resident CTAs: two to three panels in
// flight per CTA cover the DRAM latency); they land in their global layout (row tiles ld 20, k-major tiles ld 68,
// both == 4 mod 16: conflict-free DMMA fragment loads), triangular masks and ragged edges are zero-filled by the copy.
template <bool TA, bool TB>
__device__ __noinline__ void cta_gemm(int m, int n, int k, double alpha, const double* __restrict__ A, int lda,
                         const double* __restrict__ B, int ldb, double beta, double* C, int ldc, int flags,
                         CtaSmem* sm) {
    sm = cta_arena();
    // the arena sits at the start of the dynamic shared memory in every kernel of this library: address it through the
    // shared window (no generic-pointer loads, and `sm`, which arrives on the stack, is never read here)
    CtaSmem* ring = reinterpret_cast<CtaSmem*>(g_smem);
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int ar = lane >> 2, ak = lane & 3;
    // copy geometry: a row-major panel (64 x 16) gives thread t the element (t >> 4, t & 15) and every 16th row after
    // it, a k-major panel (16 x 64) the element (t >> 6, t & 63) and every 4th k after it
    constexpr int NE = (64 * GPC_BK) / GPC_THREADS;
    const int a_i = TA ? (tid & 63) : (tid >> 4), a_k = TA ? (tid >> 6) : (tid & 15);
    const int b_j = TB ? (tid >> 4) : (tid & 63), b_k = TB ? (tid & 15) : (tid >> 6);
    const int a_dst = TA ? a_k * GPC_LDS + a_i : a_i * GPC_LDR + a_k;
    const int b_dst = TB ? b_j * GPC_LDR + b_k : b_k * GPC_LDS + b_j;
    const size_t a_step = TA ? (size_t)4 * lda : (size_t)16 * lda, b_step = TB ? (size_t)16 * ldb : (size_t)4 * ldb;
    for (int i0 = 0; i0 < m; i0 += 64) {
        for (int j0 = 0; j0 < n; j0 += 64) {
            if ((flags & F_C_LOWER) && j0 > i0 + 63) continue;
            int kmin = 0, kmax = k;
            if (flags & F_A_UPPER) kmin = max(kmin, i0);
            if (flags & F_B_LOWER) kmin = max(kmin, j0);
            if (flags & F_A_LOWER) kmax = min(kmax, i0 + 64);
            if (flags & F_B_UPPER) kmax = min(kmax, j0 + 64);
            kmin &= ~(GPC_BK - 1);
            const int npan = kmax > kmin ? (kmax - kmin + GPC_BK - 1) / GPC_BK : 0;
            double acc[8][2];
#pragma unroll
            for (int t = 0; t < 8; ++t) { acc[t][0] = 0.0; acc[t][1] = 0.0; }
            const bool warp_rows_valid = i0 + warp * 8 < m && (!(flags & F_C_LOWER) || j0 <= i0 + warp * 8 + 7);
            const int nct = min(8, (n - j0 + 7) >> 3);
            // running source pointers of this thread's first element; element e sits a fixed stride further on
            const double* pa = TA ? A + (size_t)(kmin + a_k) * lda + i0 + a_i : A + (size_t)(i0 + a_i) * lda + kmin + a_k;
            const double* pb = TB ? B + (size_t)(j0 + b_j) * ldb + kmin + b_k : B + (size_t)(kmin + b_k) * ldb + j0 + b_j;
            auto issue = [&](int pan) {
                const int k0 = kmin + pan * GPC_BK;
                double* As = ring->panel[pan % GPC_STAGES][0] + a_dst;
                double* Bs = ring->panel[pan % GPC_STAGES][1] + b_dst;
#pragma unroll
                for (int e = 0; e < NE; ++e) {
                    {
                        const int gi = i0 + a_i + (TA ? 0 : 16 * e), gk = k0 + a_k + (TA ? 4 * e : 0);
                        bool ok = gi < m && gk < kmax;
                        if ((flags & F_A_LOWER) && gk > gi) ok = false;
                        if ((flags & F_A_UPPER) && gk < gi) ok = false;
                        const double* src = pa + (size_t)e * a_step;
                        cp_async8(As + e * (TA ? 4 * GPC_LDS : 16 * GPC_LDR), ok ? src : A, ok);
                    }
                    {
                        const int gj = j0 + b_j + (TB ? 16 * e : 0), gk = k0 + b_k + (TB ? 0 : 4 * e);
                        bool ok = gj < n && gk < kmax;
                        if ((flags & F_B_LOWER) && gj > gk) ok = false;
                        if ((flags & F_B_UPPER) && gj < gk) ok = false;
                        const double* src = pb + (size_t)e * b_step;
                        cp_async8(Bs + e * (TB ? 16 * GPC_LDR : 4 * GPC_LDS), ok ? src : B, ok);
                    }
                }
                pa += TA ? (size_t)GPC_BK * lda : (size_t)GPC_BK;
                pb += TB ? (size_t)GPC_BK : (size_t)GPC_BK * ldb;
            };
#pragma unroll
            for (int s = 0; s < GPC_STAGES - 1; ++s) {
                if (s < npan) issue(s);
                cp_async_commit();
            }
            for (int pan = 0; pan < npan; ++pan) {
                cp_async_wait<GPC_STAGES - 2>();   // this thread's share of panel `pan` has landed ...
                __syncthreads();                   // ... and everyone's; everyone is also done with panel pan - 1
                if (pan + GPC_STAGES - 1 < npan) issue(pan + GPC_STAGES - 1);   // refills the buffer of panel pan - 1
                cp_async_commit();
                if (warp_rows_valid) {             // ragged edge tiles: warps / column tiles beyond m, n do no math
                    const double* As = ring->panel[pan % GPC_STAGES][0];
                    const double* Bs = ring->panel[pan % GPC_STAGES][1];
#pragma unroll
                    for (int kq = 0; kq < GPC_BK; kq += 4) {
                        const double a = TA ? As[(kq + ak) * GPC_LDS + warp * 8 + ar] : As[(warp * 8 + ar) * GPC_LDR + kq + ak];
#pragma unroll
                        for (int t = 0; t < 8; ++t) {
                            if (t < nct) {
                                const double b = TB ? Bs[(t * 8 + ar) * GPC_LDR + kq + ak] : Bs[(kq + ak) * GPC_LDS + t * 8 + ar];
                                dense_dmma884(acc[t], a, b);
                            }
                        }
                    }
                }
            }
            // the old C values (beta != 0) are fetched while the last DMMAs drain, ahead of the barrier
            const int gi = i0 + warp * 8 + ar;
            double cold[8][2];
#pragma unroll
            for (int t = 0; t < 8; ++t) {
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int gj = j0 + t * 8 + 2 * ak + c;
                    const bool live = gi < m && gj < n && !((flags & F_C_LOWER) && gj > gi);
                    cold[t][c] = (beta != 0.0 && live) ? C[(size_t)gi * ldc + gj] : 0.0;
                }
            }
            cp_async_wait<0>();
            __syncthreads();       // all operand reads of this tile precede its stores (in-place use) and the next tile's copies
            if (gi < m) {
#pragma unroll
                for (int t = 0; t < 8; ++t) {
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const int gj = j0 + t * 8 + 2 * ak + c;
                        if (gj >= n) continue;
                        if ((flags & F_C_LOWER) && gj > gi) continue;
                        C[(size_t)gi * ldc + gj] = fma(beta, cold[t][c], alpha * acc[t][c]);
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
    sm = cta_arena();
    // Left-looking, four lanes per row: lane q of row i keeps l_ik, k = q (mod 4), in registers; an element is one dot
    // product and every row group recomputes the pivot from row j in shared memory: ONE barrier per column.
    double* D = sm->diag;
    const int tid = threadIdx.x;
    const int i = tid >> 2, q = tid & 3;
    double lk[16];
#pragma unroll
    for (int t = 0; t < 16; ++t) lk[t] = 0.0;
    bool ok = true;
    double lii = 0.0;
    for (int j = 0; j < nb; ++j) {
        __syncthreads();
        double pj4[4] = {0.0, 0.0, 0.0, 0.0}, pi4[4] = {0.0, 0.0, 0.0, 0.0};   // 4-deep instead of 16-deep FMA chains
#pragma unroll
        for (int t = 0; t < 16; ++t) {
            const int k = q + 4 * t;
            if (k < j) {
                const double ljk = D[j * GPC_LDD + k];
                pj4[t & 3] = fma(ljk, ljk, pj4[t & 3]);
                pi4[t & 3] = fma(lk[t], ljk, pi4[t & 3]);
            }
        }
        double pj = (pj4[0] + pj4[1]) + (pj4[2] + pj4[3]), pi = (pi4[0] + pi4[1]) + (pi4[2] + pi4[3]);
        pj += __shfl_xor_sync(0xffffffffu, pj, 1); pj += __shfl_xor_sync(0xffffffffu, pj, 2);
        pi += __shfl_xor_sync(0xffffffffu, pi, 1); pi += __shfl_xor_sync(0xffffffffu, pi, 2);
        const double djj = D[j * GPC_LDD + j] - pj;
        if (!(djj > 0.0)) { ok = false; break; }   // NaN or non-positive pivot (uniform: every thread sees the same value)
        const double ljj = sqrt(djj);
        if (i < nb && i >= j) {
            const double lij = (i == j) ? ljj : (D[i * GPC_LDD + j] - pi) / ljj;
#pragma unroll
            for (int t = 0; t < 16; ++t)
                if (q + 4 * t == j) lk[t] = lij;
            // the diagonal entry still holds a_jj, which slower warps read for this column's pivot: it is replaced
            // by l_jj only after the loop
            if (i == j) lii = lij;
            else if (q == 0) D[i * GPC_LDD + j] = lij;
        }
    }
    __syncthreads();
    if (ok && q == 0 && i < nb) D[i * GPC_LDD + i] = lii;
    __syncthreads();
    return ok;
}

// inv (global, nb x nb row-major, ld = GPC_BLK) <- inverse of the lower-triangular block in sm->diag.
// Four threads per column split the substitution dot products.
__device__ void smem_tri_inverse(CtaSmem* sm, int nb, double* inv) {
    sm = cta_arena();
    const double* D = sm->diag;
    const int tid = threadIdx.x;
    const int c = tid >> 2, q = tid & 3;
    // each thread keeps its strided share of column c in registers: entries k = q, q+4, ... (16 values)
    double xk[16];
#pragma unroll
    for (int t = 0; t < 16; ++t) xk[t] = 0.0;
    for (int i = 0; i < nb; ++i) {
        double p4[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int t = 0; t < 16; ++t) {
            const int kk = q + 4 * t;
            if (kk < i) p4[t & 3] = fma(D[i * GPC_LDD + kk], xk[t], p4[t & 3]);
        }
        double part = (p4[0] + p4[1]) + (p4[2] + p4[3]);
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
    sm = cta_arena();
    const int tid = threadIdx.x;
    for (int j0 = 0, blk = 0; j0 < n; j0 += GPC_BLK, ++blk) {
        const int jb = min(GPC_BLK, n - j0);
        // panel update: A[j0:n, j0:j0+jb] -= L[j0:n, 0:j0] * L[j0:j0+jb, 0:j0]^T
        if (j0 > 0)
            cta_gemm<false, true>(n - j0, jb, j0, -1.0, A + (size_t)j0 * lda, lda, A + (size_t)j0 * lda, lda, 1.0,
                                  A + (size_t)j0 * lda + j0, lda, 0, sm);
        // diagonal block -> smem, factor, write back, invert
        GPC_FOR_2D(jb, jb, i, c) {
            sm->diag[i * GPC_LDD + c] = (c <= i) ? A[(size_t)(j0 + i) * lda + j0 + c] : 0.0;
        }
        if (!smem_chol(sm, jb)) return false;
        GPC_FOR_2D(jb, jb, i, c) {
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
    sm = cta_arena();
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
    sm = cta_arena();
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
    sm = cta_arena();
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
    sm = cta_arena();
    // T = tril(L^T * Lbar)
    cta_gemm<true, false>(n, n, n, 1.0, L, ldl, G, ldg, 0.0, T, ldt, F_A_UPPER | F_B_LOWER | F_C_LOWER, sm);
    GPC_FOR_2D(n, n, i, j) {
        const double v = T[(size_t)i * ldt + j];
        G[(size_t)i * ldg + j] = (j < i) ? v : (j == i ? 0.5 * v : 0.0);
    }
    __syncthreads();
    cta_trsm_left_t(n, n, L, ldl, dinv, G, ldg, sm);
    cta_trsm_right(n, n, L, ldl, dinv, G, ldg, sm);
}
