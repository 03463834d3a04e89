// Fused on-chip SVGP evaluation for M <= 64 inducing points (the headline configuration: 20k genes x 1k spots,
// M = 64): one CTA evaluates ELBO + gradient of one gene with every M x M factor and every M x 64 column tile of
// the N data points resident in shared memory; HBM sees only y / scale / X / Z / theta in and the gradient out.
//
// Math: SURVEY.md appendix B.2 (reverse pass of gpflow SVGP.elbo -> base_conditional + gauss_kl, whitened;
// reference call site GPcounts_Module.py:679-680).  Per 64-column tile of the data:
//     Kuf -> A = Lm^-1 Kuf -> B = Lq^T A -> (fmean, fvar) -> 20-node quadrature -> (g_m, g_v)
//     Bbar = 2 B diag(g_v);  Abar = q_mu g_m^T - 2 A diag(g_v) + Lq Bbar;  S = Lm^-T Abar
//     GLq += A Bbar^T (lower);  GLm += S A^T (lower);  gq += A g_m;  <S o Kuf, R^2>, <Abar, A>
// The triangular solves are products with the explicit inverse of the 64 x 64 factor (formed once per
// evaluation in shared memory); all six products per tile run on the FP64 tensor pipe (DMMA m8n8k4), one
// 8-row block of the output per warp so that the triangular k-ranges are warp-uniform.  The two rank updates
// keep their accumulators in registers across tiles and are paired (row block w with row block nrb-1-w) so
// every warp does the same amount of work.
#pragma once
#include "model.cuh"

#define FZ_LD 68                    // leading dimension of every smem matrix (== 4 mod 16: conflict-free DMMA fragments)
#define FZ_MAT (64 * FZ_LD)         // doubles per 64 x 64 smem matrix

struct FusedSmem {
    double T0[FZ_MAT];              // Kuf tile            (end phase: Kuu0 / scratch)
    double T1[FZ_MAT];              // A tile              (end phase: GLq / P / Kbar)
    double T2[FZ_MAT];              // B, Bbar, then S     (end phase: Lm)
    double T3[FZ_MAT];              // Abar                (end phase: Lm-bar / T)
    double Linv[FZ_MAT];            // Lm^-1 (lower)
    double Lq[FZ_MAT];              // tril(q_sqrt)
    double qmu[64];
    double gq[64];                  // A g_m accumulator
    double part[4][64][4];          // per column partial sums (4 parts x 64 columns x 4 values)
    double col[8][64];              // per column: fmean, fvar, gm, gv, y, scale, 2*gv, valid
    double red[64];
    double Zs[64 * 2];              // inducing inputs (d <= 2)
};

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// Warp-level product on smem operands: for column tiles ct in [ct0, ct1):
//   acc[ct] (8 x 8) += sum_{k in [kbeg, kend)} Aop(r0 + 0..7, k) * Bop(k, 8 ct + 0..7),   kbeg, kend multiples of 4
// AT: A stored [k][row] (operand is a transpose) else [row][k];  BT: B stored [col][k] else [k][col].
// KLO != 0: additionally start each column tile at k >= 8 ct (operand B lower-triangular in (k, col)).
template <bool AT, bool BT, bool KLO>
__device__ __forceinline__ void warp_mma(double (&acc)[8][2], const double* __restrict__ A, int r0,
                                         const double* __restrict__ B, int kbeg, int kend, int ct0, int ct1) {
    const int lane = threadIdx.x & 31, ar = lane >> 2, ak = lane & 3;
    for (int k0 = kbeg; k0 < kend; k0 += 4) {
        const double a = AT ? A[(k0 + ak) * FZ_LD + r0 + ar] : A[(r0 + ar) * FZ_LD + k0 + ak];
#pragma unroll
        for (int ct = 0; ct < 8; ++ct) {
            if (ct >= ct0 && ct < ct1 && (!KLO || k0 + 4 > 8 * ct)) {
                const double b = BT ? B[(ct * 8 + ar) * FZ_LD + k0 + ak] : B[(k0 + ak) * FZ_LD + ct * 8 + ar];
                dmma884(acc[ct], a, b);
            }
        }
    }
}

__device__ __forceinline__ void acc_zero(double (&acc)[8][2]) {
#pragma unroll
    for (int t = 0; t < 8; ++t) { acc[t][0] = 0.0; acc[t][1] = 0.0; }
}

// Store the warp's 8 x 64 accumulator block to smem matrix C (rows r0..r0+7).
__device__ __forceinline__ void acc_store(const double (&acc)[8][2], double* C, int r0, int ct0, int ct1) {
    const int lane = threadIdx.x & 31, ar = lane >> 2, ac = 2 * (lane & 3);
#pragma unroll
    for (int ct = 0; ct < 8; ++ct)
        if (ct >= ct0 && ct < ct1)
            *reinterpret_cast<double2*>(&C[(r0 + ar) * FZ_LD + ct * 8 + ac]) = make_double2(acc[ct][0], acc[ct][1]);
}

// In-place Cholesky of the n x n SPD matrix in smem D (ld FZ_LD, lower part used); false on a bad pivot.
__device__ bool fz_chol(double* D, int n) {
    const int tid = threadIdx.x;
    for (int j = 0; j < n; ++j) {
        __syncthreads();
        const double djj = D[j * FZ_LD + j];
        if (!(djj > 0.0)) return false;
        const double r = sqrt(djj);
        __syncthreads();
        if (tid == j) D[j * FZ_LD + j] = r;
        if (tid > j && tid < n) D[tid * FZ_LD + j] /= r;
        __syncthreads();
        const int rem = n - j - 1;
        for (int idx = tid; idx < rem * rem; idx += GPC_THREADS) {
            const int i = j + 1 + idx / rem, c = j + 1 + idx % rem;
            if (c <= i) D[i * FZ_LD + c] = fma(-D[i * FZ_LD + j], D[c * FZ_LD + j], D[i * FZ_LD + c]);
        }
    }
    __syncthreads();
    return true;
}

// X (smem, 64 x 64, ld FZ_LD) <- inverse of the lower-triangular n x n matrix L (smem), identity-padded to 64.
__device__ void fz_tri_inverse(const double* L, int n, double* X) {
    const int tid = threadIdx.x;
    const int c = tid >> 2, q = tid & 3;
    double xk[16];
#pragma unroll
    for (int t = 0; t < 16; ++t) xk[t] = 0.0;
    for (int i = 0; i < 64; ++i) {
        double xi;
        if (i < n) {
            double part = 0.0;
#pragma unroll
            for (int t = 0; t < 16; ++t) {
                const int kk = q + 4 * t;
                if (kk < i) part = fma(L[i * FZ_LD + kk], xk[t], part);
            }
            part += __shfl_xor_sync(0xffffffffu, part, 1);
            part += __shfl_xor_sync(0xffffffffu, part, 2);
            xi = (c < n && i >= c) ? ((i == c ? 1.0 : 0.0) - part) / L[i * FZ_LD + i] : 0.0;
        } else {
            xi = (i == c) ? 1.0 : 0.0;
        }
#pragma unroll
        for (int t = 0; t < 16; ++t)
            if (q + 4 * t == i) xk[t] = xi;
        if (q == 0) X[i * FZ_LD + c] = xi;
    }
    __syncthreads();
}

// Quadrature of one data point restricted to nodes [i0, i1): partial (ve, gm, gz, dalpha, dkm) sums.
// Mirrors lik_point() in lik.cuh (same formulas, same order inside a node).
struct LikPart { double ve, gm, gz, da, dk; bool poison; };

__device__ __forceinline__ LikPart lik_nodes(const LikParams& lp, double fm, double sd, double y, double logs, int i0,
                                             int i1) {
    LikPart o{0.0, 0.0, 0.0, 0.0, 0.0, false};
    const double alpha = lp.alpha, k = lp.k, ia = k, ia2 = k * k, km = lp.km;
    for (int i = i0; i < i1; ++i) {
        const double z = c_gh_z[i], w = c_gh_w[i];
        if (lp.lik == GPC_NB) {
            const double logm = fm + sd * z + logs;
            const double m = exp(logm);
            const double t = 1.0 + m * alpha;
            const double lt = log(t);
            const double l = y * (logm + lp.log_alpha - lt) - k * lt;
            const double dl = (y - m) / t;
            const double dal = y * ia + lt * ia2 - (y + k) * m / t;
            o.ve = fma(w, l, o.ve); o.gm = fma(w, dl, o.gm); o.gz = fma(w * z, dl, o.gz); o.da = fma(w, dal, o.da);
        } else {   // ZINB, see lik.cuh for the reference-faithful psi arithmetic
            const double F = fm + sd * z;
            const double m = exp(F);
            const double t = 1.0 + m * alpha;
            const double lt = log(t);
            const double kmm = km + m;
            const double q = m / kmm;
            const double psi = 1.0 - q;
            const double omp = 1.0 - psi;
            o.poison = o.poison || (psi == 0.0);
            if (y == 0.0) {
                const double a = log(psi);
                const double b = log(omp) - lt * ia;
                const double mx = fmax(a, b);
                const double ea = exp(a - mx), eb = exp(b - mx);
                const double se = ea + eb;
                const double l = mx + log(se);
                const double ra = ea / se, rb = eb / se;
                const double dq = q * (km / kmm);
                const double dqk = m / (kmm * kmm);
                const double dl = -ra * dq / psi + rb * (dq / omp - m / t);
                o.ve = fma(w, l, o.ve); o.gm = fma(w, dl, o.gm); o.gz = fma(w * z, dl, o.gz);
                o.dk = fma(w, ra * dqk / psi - rb * dqk / omp, o.dk);
                o.da = fma(w, rb * (lt * ia2 - m * ia / t), o.da);
            } else {
                const double l = log(omp) + y * (F + lp.log_alpha - lt) - k * lt;
                const double dl = q * (km / kmm) / omp + (y - m) / t;
                o.ve = fma(w, l, o.ve); o.gm = fma(w, dl, o.gm); o.gz = fma(w * z, dl, o.gz);
                o.dk = fma(w, -(m / (kmm * kmm)) / omp, o.dk);
                o.da = fma(w, y * ia + lt * ia2 - (y + k) * m / t, o.da);
            }
        }
    }
    return o;
}

// ELBO + gradient of one SVGP gene-model, M <= 64.  Same contract as eval_svgp().
__device__ __noinline__ int eval_svgp_fused(const ModelDev& md, const double* Z, const double* y, const double* scale,
                               const double* theta, double* grad, double* f_out, const SlotWs& ws, FusedSmem* fs) {
    const int N = md.N, M = md.M, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nrb = (M + 7) >> 3, Mp = nrb * 8;          // 8-row blocks in use
    const Hyper h = hyper_from_theta(md, theta);
    KernParams kp;
    kern_setup(kp, md, h.var, h.ls);
    LikParams lp;
    lik_setup(lp, md.lik, h.alpha, h.km);
    const double* q_mu = theta + GPC_NHYP;
    const double* lq_packed = theta + GPC_NHYP + M;
    const double shift = md.jitter + (md.kernel == GPC_BRANCH ? GPC_BRANCH_NOISE : 0.0);
    const double kdiag = h.var + (md.kernel == GPC_BRANCH ? GPC_BRANCH_NOISE : 0.0);
    const bool need_kg = (md.train_mask & 3) != 0;
    const int dd = md.d;

    // ---- setup: Z, q_mu, Lq, Kuu -> Lm -> Linv ----------------------------------------------------------------
    for (int i = tid; i < 64 * 2; i += GPC_THREADS) fs->Zs[i] = (i < M * dd) ? Z[i] : 0.0;
    if (tid < 64) { fs->qmu[tid] = tid < M ? q_mu[tid] : 0.0; fs->gq[tid] = 0.0; }
    __syncthreads();
    for (int idx = tid; idx < 64 * 64; idx += GPC_THREADS) {
        const int i = idx >> 6, j = idx & 63;
        double lq = (i == j) ? 1.0 : 0.0, kv = 0.0;
        if (i < M && j < M) {
            lq = (j <= i) ? lq_packed[tri_index(i, j)] : 0.0;
            if (j <= i) {
                bool cross;
                const double r2 = kern_r2(kp, fs->Zs + i * dd, fs->Zs + j * dd, cross);
                kv = kern_val(kp, r2, cross) + (i == j ? shift : 0.0);
            }
        }
        fs->Lq[i * FZ_LD + j] = lq;
        fs->T2[i * FZ_LD + j] = kv;
    }
    if (!fz_chol(fs->T2, M)) return GPC_ST_CHOL;
    fz_tri_inverse(fs->T2, M, fs->Linv);
    for (int idx = tid; idx < 64 * 64; idx += GPC_THREADS) {       // keep Lm for the end phase
        const int i = idx >> 6, j = idx & 63;
        ws.L[idx] = (i < M && j <= i) ? fs->T2[i * FZ_LD + j] : 0.0;
    }

    // ---- column tiles -------------------------------------------------------------------------------------------
    double accLq[8][2], accLm[8][2];
    acc_zero(accLq);
    acc_zero(accLm);
    const int rbq = warp, rbm = nrb - 1 - warp;          // row blocks of the two rank updates owned by this warp
    double sum_ve = 0.0, sum_da = 0.0, sum_dk = 0.0, sum_gv = 0.0, sum_sk = 0.0, sum_skr = 0.0;
    bool poison_any = false;
    const int r0 = warp * 8;
    const bool active = warp < nrb;
    for (int n0 = 0; n0 < N; n0 += 64) {
        const int nt = min(64, N - n0);
        __syncthreads();
        // Kuf tile (zero beyond M rows / nt columns); per-column inputs
        for (int idx = tid; idx < 64 * 64; idx += GPC_THREADS) {
            const int i = idx >> 6, c = idx & 63;
            double kv = 0.0;
            if (i < M && c < nt) {
                bool cross;
                const double r2 = kern_r2(kp, fs->Zs + i * dd, md.X + (size_t)(n0 + c) * dd, cross);
                kv = kern_val(kp, r2, cross);
            }
            fs->T0[i * FZ_LD + c] = kv;
        }
        if (tid < 64) {
            const bool v = tid < nt;
            fs->col[4][tid] = v ? y[n0 + tid] : 0.0;
            fs->col[5][tid] = (v && scale) ? scale[n0 + tid] : 1.0;
        }
        __syncthreads();
        double acc[8][2];
        // A = Linv * Kuf      (rows r0.., k <= r0 + 7)
        if (active) {
            acc_zero(acc);
            warp_mma<false, false, false>(acc, fs->Linv, r0, fs->T0, 0, r0 + 8, 0, 8);
            acc_store(acc, fs->T1, r0, 0, 8);
        }
        __syncthreads();
        // B = Lq^T * A        (k >= r0)
        if (active) {
            acc_zero(acc);
            warp_mma<true, false, false>(acc, fs->Lq, r0, fs->T1, r0, Mp, 0, 8);
            acc_store(acc, fs->T2, r0, 0, 8);
        }
        __syncthreads();
        // per-column statistics: 4 threads per column over the rows
        {
            const int c = tid & 63, part = tid >> 6;
            double s1 = 0.0, s2 = 0.0, s3 = 0.0;
            for (int i = part; i < Mp; i += 4) {
                const double a = fs->T1[i * FZ_LD + c], b = fs->T2[i * FZ_LD + c];
                s1 = fma(a, fs->qmu[i], s1);
                s2 = fma(a, a, s2);
                s3 = fma(b, b, s3);
            }
            fs->part[part][c][0] = s1; fs->part[part][c][1] = s2; fs->part[part][c][2] = s3;
        }
        __syncthreads();
        if (tid < 64) {
            double s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
            for (int p = 0; p < 4; ++p) { s1 += fs->part[p][tid][0]; s2 += fs->part[p][tid][1]; s3 += fs->part[p][tid][2]; }
            fs->col[0][tid] = s1;
            fs->col[1][tid] = (kdiag - s2) + s3;
        }
        __syncthreads();
        // quadrature: 4 threads per column, 5 nodes each (Poisson: closed form by part 0)
        {
            const int c = tid & 63, part = tid >> 6;
            const double fm = fs->col[0][c], fv = fs->col[1][c], yy = fs->col[4][c];
            double ve = 0.0, gm = 0.0, gz = 0.0, da = 0.0, dk = 0.0;
            bool poison = false;
            if (c < nt) {
                if (md.lik == GPC_POISSON) {
                    if (part == 0) {
                        const double e = exp(fm + 0.5 * fv);
                        ve = yy * fm - e - lgamma(yy + 1.0);
                        gm = yy - e;
                        gz = -0.5 * e;            // holds g_v directly for Poisson
                    }
                } else {
                    const double sd = sqrt(fv);
                    const LikPart o = lik_nodes(lp, fm, sd, yy, log(fs->col[5][c]), part * 5, part * 5 + 5);
                    ve = o.ve; gm = o.gm; gz = o.gz; da = o.da; dk = o.dk; poison = o.poison;
                    const bool has_const = (md.lik == GPC_NB) || (yy != 0.0);
                    if (part == 1 && has_const) ve += lgamma(lp.k + yy) - lgamma(yy + 1.0) - lp.lgamma_k;
                    if (part == 2 && has_const) da += -(lp.k * lp.k) * (digamma_pos(lp.k + yy) - lp.digamma_k);
                }
            }
            if (poison) { gm = NAN; gz = NAN; dk = NAN; }
            fs->part[part][c][0] = gm; fs->part[part][c][1] = gz;
            sum_ve += ve; sum_da += da; sum_dk += dk;
        }
        __syncthreads();
        if (tid < 64) {
            double gm = 0.0, gz = 0.0;
#pragma unroll
            for (int p = 0; p < 4; ++p) { gm += fs->part[p][tid][0]; gz += fs->part[p][tid][1]; }
            double gv = 0.0;
            if (tid < nt) gv = (md.lik == GPC_POISSON) ? gz : gz / (2.0 * sqrt(fs->col[1][tid]));
            else gm = 0.0;
            fs->col[2][tid] = gm;
            fs->col[3][tid] = gv;
            fs->col[6][tid] = 2.0 * gv;
            sum_gv += gv;
        }
        __syncthreads();
        // Bbar = B * diag(2 gv) in place;  gq += A gm
        for (int idx = tid; idx < Mp * 64; idx += GPC_THREADS) {
            const int i = idx >> 6, c = idx & 63;
            fs->T2[i * FZ_LD + c] *= fs->col[6][c];
        }
        for (int row = warp; row < Mp; row += 8) {
            double s = fs->T1[row * FZ_LD + lane] * fs->col[2][lane] + fs->T1[row * FZ_LD + lane + 32] * fs->col[2][lane + 32];
            s = warp_sum(s);
            if (lane == 0) fs->gq[row] += s;
        }
        __syncthreads();
        // Abar = q_mu gm^T - A diag(2 gv) + Lq * Bbar      (k <= r0 + 7)
        if (active) {
            acc_zero(acc);
            warp_mma<false, false, false>(acc, fs->Lq, r0, fs->T2, 0, r0 + 8, 0, 8);
            const int ar = lane >> 2, ac = 2 * (lane & 3);
            const double qm = fs->qmu[r0 + ar];
            double sab = 0.0;
#pragma unroll
            for (int ct = 0; ct < 8; ++ct) {
                const int c = ct * 8 + ac;
                const double2 a = *reinterpret_cast<const double2*>(&fs->T1[(r0 + ar) * FZ_LD + c]);
                const double v0 = acc[ct][0] + qm * fs->col[2][c] - a.x * fs->col[6][c];
                const double v1 = acc[ct][1] + qm * fs->col[2][c + 1] - a.y * fs->col[6][c + 1];
                sab = fma(v0, a.x, sab);
                sab = fma(v1, a.y, sab);
                *reinterpret_cast<double2*>(&fs->T3[(r0 + ar) * FZ_LD + c]) = make_double2(v0, v1);
            }
            sum_sk += sab;                                   // <Abar, A> = <S, Kuf>
        }
        // GLq += A * Bbar^T   (row block rbq, column tiles 0..rbq)
        if (rbq < nrb) warp_mma<false, true, false>(accLq, fs->T1, rbq * 8, fs->T2, 0, 64, 0, rbq + 1);
        __syncthreads();
        if (need_kg) {
            // S = Linv^T * Abar   (k >= r0) -> T2
            if (active) {
                acc_zero(acc);
                warp_mma<true, false, false>(acc, fs->Linv, r0, fs->T3, r0, Mp, 0, 8);
                acc_store(acc, fs->T2, r0, 0, 8);
            }
            __syncthreads();
            // GLm += S * A^T      (row block rbm, column tiles 0..rbm)
            if (rbm >= 0 && rbm < nrb) warp_mma<false, true, false>(accLm, fs->T2, rbm * 8, fs->T1, 0, 64, 0, rbm + 1);
            // <S o Kuf, R^2>
            if (md.kernel != GPC_CONST) {
                for (int idx = tid; idx < M * 64; idx += GPC_THREADS) {
                    const int i = idx >> 6, c = idx & 63;
                    if (c < nt) {
                        bool cross;
                        const double r2 = kern_r2(kp, fs->Zs + i * dd, md.X + (size_t)(n0 + c) * dd, cross);
                        sum_skr = fma(fs->T2[i * FZ_LD + c] * fs->T0[i * FZ_LD + c], r2, sum_skr);
                    }
                }
            }
        }
        poison_any = poison_any;
    }
    __syncthreads();
    // ---- end phase ------------------------------------------------------------------------------------------------
    // GLq -> T1 (lower), GLm -> T3 (lower, sign applied later)
    for (int idx = tid; idx < 64 * 64; idx += GPC_THREADS) {
        const int i = idx >> 6, j = idx & 63;
        fs->T1[i * FZ_LD + j] = 0.0;
        fs->T3[i * FZ_LD + j] = 0.0;
    }
    __syncthreads();
    if (rbq < nrb) acc_store(accLq, fs->T1, rbq * 8, 0, rbq + 1);
    if (rbm >= 0 && rbm < nrb) acc_store(accLm, fs->T3, rbm * 8, 0, rbm + 1);
    __syncthreads();
    double* glq = grad + GPC_NHYP + M;
    for (int idx = tid; idx < M * M; idx += GPC_THREADS) {
        const int i = idx / M, j = idx % M;
        if (j <= i) {
            const double l = fs->Lq[i * FZ_LD + j];
            glq[tri_index(i, j)] = fs->T1[i * FZ_LD + j] - (i == j ? l - 1.0 / l : l);
        }
    }
    if (tid < M) grad[GPC_NHYP + tid] = fs->gq[tid] - fs->qmu[tid];
    // reductions of the per-thread sums
    double acc6[6] = {sum_ve, sum_da, sum_dk, sum_gv, sum_sk, sum_skr};
    cta_sum_n<6>(acc6, fs->red);
    // KL
    double klacc[3] = {0.0, 0.0, 0.0};
    for (int idx = tid; idx < M * M; idx += GPC_THREADS) {
        const int i = idx / M, j = idx % M;
        if (j <= i) { const double l = fs->Lq[i * FZ_LD + j]; klacc[1] = fma(l, l, klacc[1]); if (i == j) klacc[2] += log(l * l); }
    }
    if (tid < M) klacc[0] = fs->qmu[tid] * fs->qmu[tid];
    cta_sum_n<3>(klacc, fs->red);
    const double kl = 0.5 * (klacc[0] - (double)M + klacc[1] - klacc[2]);
    double gvar = 0.0, gls = 0.0;
    if (need_kg) {
        // Lm -> T2;  Lm-bar = -tril(GLm) -> T3 (lower, zero upper)
        for (int idx = tid; idx < 64 * 64; idx += GPC_THREADS) {
            const int i = idx >> 6, j = idx & 63;
            fs->T2[i * FZ_LD + j] = ws.L[idx];
            const double v = fs->T3[i * FZ_LD + j];
            fs->T3[i * FZ_LD + j] = (j <= i) ? -v : 0.0;
        }
        __syncthreads();
        double acc[8][2];
        // P = Phi(Lm^T * Lmbar) -> T1 (lower):  P[i][j] = sum_{k >= i} Lm[k][i] Lmbar[k][j]
        if (active) {
            acc_zero(acc);
            warp_mma<true, false, false>(acc, fs->T2, r0, fs->T3, r0, Mp, 0, warp + 1);
            const int ar = lane >> 2, ac = 2 * (lane & 3);
#pragma unroll
            for (int ct = 0; ct < 8; ++ct) {
                const int i = r0 + ar, j = ct * 8 + ac;
                double v0 = (ct <= warp) ? acc[ct][0] : 0.0, v1 = (ct <= warp) ? acc[ct][1] : 0.0;
                v0 = (j < i) ? v0 : (j == i ? 0.5 * v0 : 0.0);
                v1 = (j + 1 < i) ? v1 : (j + 1 == i ? 0.5 * v1 : 0.0);
                *reinterpret_cast<double2*>(&fs->T1[i * FZ_LD + j]) = make_double2(v0, v1);
            }
        }
        __syncthreads();
        // T = Linv^T * P -> T3:  T[i][j] = sum_{k >= max(i, j)} Linv[k][i] P[k][j]
        if (active) {
            acc_zero(acc);
            warp_mma<true, false, true>(acc, fs->Linv, r0, fs->T1, r0, Mp, 0, nrb);
            acc_store(acc, fs->T3, r0, 0, 8);
        }
        __syncthreads();
        // Kbar = T * Linv -> T1:  Kbar[i][j] = sum_{k >= j} T[i][k] Linv[k][j]
        if (active) {
            acc_zero(acc);
            warp_mma<false, false, true>(acc, fs->T3, r0, fs->Linv, 0, Mp, 0, nrb);
            acc_store(acc, fs->T1, r0, 0, 8);
        }
        __syncthreads();
        double a2[2] = {0.0, 0.0};
        for (int idx = tid; idx < M * M; idx += GPC_THREADS) {
            const int i = idx / M, j = idx % M;
            bool cross;
            const double r2 = kern_r2(kp, fs->Zs + i * dd, fs->Zs + j * dd, cross);
            const double k0 = kern_val(kp, r2, cross);
            kern_grad_acc(kp, fs->T1[i * FZ_LD + j], k0, r2, cross, a2[0], a2[1]);
        }
        cta_sum_n<2>(a2, fs->red);
        if (md.kernel == GPC_CONST) {
            gvar = acc6[3] + acc6[4] / h.var + a2[0];
        } else {
            // <S, Kuf>/var only holds for kernels proportional to var; BranchKernel cross terms are not, so the
            // generic path (dense evaluator) handles BRANCH; here RBF: d Kuf / d var = Kuf / var
            gvar = acc6[3] + acc6[4] / h.var + a2[0];
            gls = acc6[5] * kp.inv_ls3 + a2[1];
        }
    }
    store_hyper_grad(md, h, gvar, gls, acc6[1], acc6[2], grad);
    *f_out = acc6[0] - kl;
    __syncthreads();
    return GPC_ST_OK;
}
