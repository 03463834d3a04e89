"""Oracle: FP64 CPU restatement of one ELBO evaluation of the GPcounts hot path (test infrastructure).

Follows the GPflow model code the reference calls (third-party, un-vendored):
  * ``gpflow.models.VGP.elbo``            (reference call site ``GPcounts/GPcounts_Module.py:685-686``)
  * ``gpflow.models.SVGP.elbo`` -> ``conditionals.base_conditional`` + ``kullback_leiblers.gauss_kl``
    with ``whiten=True``                  (``GPcounts_Module.py:679-680``)
  * ``gpflow.models.GPR`` / ``SGPR``      (``GPcounts_Module.py:661-667``)
  * ``gpflow.kernels.RBF`` / ``Constant`` (``GPcounts_Module.py:574, 601-604``)
  * GPcounts' ``BranchKernel``            (``GPcounts/branchingKernel.py:31-74``)
  * GPcounts' NB / ZINB log-densities     (``GPcounts/NegativeBinomialLikelihood.py:23-25, 37-48, 63-77``)
  * ``gpflow.likelihoods.Poisson`` closed-form variational expectation (``GPcounts_Module.py:622-623``)
  * 20-point Gauss-Hermite quadrature of ``ScalarLikelihood``.

Gradients come from torch autograd, which makes this an independent check of the hand-derived
reverse pass in the CUDA kernels.

Parameter vector layout (shared with the CUDA path so gradients compare element-wise):
  theta[0] = softplus^-1(kernel variance)      theta[1] = softplus^-1(lengthscale)
  theta[2] = softplus^-1(alpha)  (NB/ZINB)  or softplus^-1(noise variance - 1e-6)  (Gaussian)
  theta[3] = softplus^-1(km)     (ZINB)
  theta[4 : 4+Mv]          q_mu         (Mv = N for VGP, M for SVGP; absent for GPR/SGPR)
  theta[4+Mv : ]           tril(q_sqrt) row-major packed ((0,0),(1,0),(1,1),(2,0),...)
GPflow's own ordering of the triangle (TFP FillTriangular) differs, which is irrelevant to L-BFGS.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import numpy as np
import torch

torch.set_default_dtype(torch.float64)

N_GH = 20
_gh_x, _gh_w = np.polynomial.hermite.hermgauss(N_GH)
GH_Z = torch.tensor(_gh_x * np.sqrt(2.0))          # nodes   z_i = sqrt(2) x_i
GH_W = torch.tensor(_gh_w / np.sqrt(np.pi))        # weights w_i / sqrt(pi)
JITTER = 1e-6                                      # gpflow.config.default_jitter()
BRANCH_NOISE = 1e-6                                # BranchKernel(noise_level=1e-6), branchingKernel.py:23
GAUSS_VAR_LOWER = 1e-6                             # gpflow.likelihoods.Gaussian DEFAULT_VARIANCE_LOWER_BOUND
NHYP = 4


def softplus(x):
    return torch.logaddexp(x, torch.zeros_like(x))


def softplus_inv(p):
    """tfp.math.softplus_inverse: x = p + log(-expm1(-p))."""
    p = np.asarray(p, dtype=np.float64)
    return p + np.log(-np.expm1(-p))


@dataclass
class ModelSpec:
    """One (model, kernel, likelihood) variant of the hot path."""
    kind: str = "VGP"            # VGP | SVGP | GPR | SGPR
    kernel: str = "RBF"          # RBF | CONST | BRANCH
    lik: str = "NB"              # NB | ZINB | POISSON | GAUSS
    X: Optional[np.ndarray] = None      # (N, d); BRANCH: (N, 2) = [t, label]
    Z: Optional[np.ndarray] = None      # (M, d) inducing points (SVGP/SGPR)
    xp: float = -1000.0                 # branching point
    train: tuple = (True, True, True, True)   # trainable mask for (var, ls, alpha|noise, km)
    jitter: float = JITTER

    @property
    def N(self):
        return self.X.shape[0]

    @property
    def Mv(self):
        if self.kind == "VGP":
            return self.X.shape[0]
        if self.kind == "SVGP":
            return self.Z.shape[0]
        return 0

    @property
    def P(self):
        mv = self.Mv
        return NHYP + mv + mv * (mv + 1) // 2

    def trainable_index(self):
        """Indices of theta that the optimiser sees (gpflow ``trainable_variables``)."""
        used = [True, self.kernel != "CONST", self.lik in ("NB", "ZINB", "GAUSS"), self.lik == "ZINB"]
        idx = [i for i in range(NHYP) if used[i] and self.train[i]]
        return np.array(idx + list(range(NHYP, self.P)), dtype=np.int64)


# ----------------------------------------------------------------------------- kernels
def square_distance(X, X2):
    """gpflow.utilities.ops.square_distance (expansion form)."""
    Xs = (X * X).sum(-1)
    X2s = (X2 * X2).sum(-1)
    return -2.0 * X @ X2.T + Xs[:, None] + X2s[None, :]


def rbf_K(X, X2, var, ls):
    """gpflow.kernels.SquaredExponential.K: var * exp(-r2/2) on inputs scaled by 1/ls."""
    return var * torch.exp(-0.5 * square_distance(X / ls, X2 / ls))


def kernel_K(spec, X, X2, var, ls, square=False):
    if spec.kernel == "CONST":                       # gpflow.kernels.Constant
        return var * torch.ones(X.shape[0], X2.shape[0])
    if spec.kernel == "RBF":
        return rbf_K(X, X2, var, ls)
    # BranchKernel.K, branchingKernel.py:31-70
    t1, t2 = X[:, 0:1], X2[:, 0:1]
    same = X[:, 1:2] == X2[:, 1:2].T                  # :41-45
    Ktt = rbf_K(t1, t2, var, ls)                      # :48
    Bs = torch.full((1, 1), float(spec.xp))           # :50
    kbb = rbf_K(Bs, Bs, var, ls) + JITTER             # :51-55
    Kb1 = rbf_K(t1, Bs, var, ls)                      # :59
    Kb2 = rbf_K(t2, Bs, var, ls)                      # :60
    Kcross = (Kb1 / kbb) @ Kb2.T                      # :57-62
    K = torch.where(same, Ktt, Kcross)                # :63
    if square:
        K = K + BRANCH_NOISE * torch.eye(K.shape[0])  # :65-68
    return K


def kernel_Kdiag(spec, X, var):
    if spec.kernel == "BRANCH":                       # branchingKernel.py:72-74
        return var + BRANCH_NOISE + torch.zeros(X.shape[0])
    return var + torch.zeros(X.shape[0])


# ----------------------------------------------------------------------------- likelihoods
def nb_logp(m, Y, alpha):
    """negative_binomial(), NegativeBinomialLikelihood.py:37-48."""
    k = 1.0 / alpha
    return (torch.lgamma(k + Y) - torch.lgamma(Y + 1.0) - torch.lgamma(k)
            + Y * torch.log(m / (m + k)) - k * torch.log(1.0 + m * alpha))


def zinb_logp(F, Y, alpha, km):
    """ZeroInflatedNegativeBinomial._scalar_log_prob, NegativeBinomialLikelihood.py:63-77."""
    m = torch.exp(F)
    psi = 1.0 - (m / (km + m))
    nb_zero = -torch.log(1.0 + m * alpha) / alpha
    log_p_zero = torch.logsumexp(torch.stack([torch.log(psi), torch.log(1.0 - psi) + nb_zero]), dim=0)
    # guard the unused branch so autograd does not propagate NaNs from it (tf.where semantics differ
    # only in the gradient of the masked-out branch, which is multiplied by zero in both frameworks)
    Ysafe = torch.where(Y == 0, torch.ones_like(Y), Y)
    log_p_nonzero = torch.log(1.0 - psi) + nb_logp(m, Ysafe, alpha)
    return torch.where(Y == 0, log_p_zero, log_p_nonzero)


def variational_expectations(spec, fmean, fvar, y, scale, alpha, km):
    """Sum_n E_{q(f_n)}[log p(y_n|f_n)] - ScalarLikelihood GH-20 quadrature / Poisson closed form."""
    if spec.lik == "POISSON":
        # gpflow.likelihoods.Poisson._variational_expectations, invlink = exp, binsize = 1
        return (y * fmean - torch.exp(fmean + fvar / 2.0) - torch.lgamma(y + 1.0)).sum()
    sd = torch.sqrt(fvar)
    F = fmean[None, :] + sd[None, :] * GH_Z[:, None]            # [Q, N]
    Y = y[None, :].expand_as(F)
    if spec.lik == "NB":
        s = scale[None, :] if scale is not None else 1.0        # intent of GPcounts_Module.py:629-636
        logp = nb_logp(torch.exp(F) * s, Y, alpha)
    elif spec.lik == "ZINB":
        logp = zinb_logp(F, Y, alpha, km)
    else:
        raise ValueError(spec.lik)
    return (GH_W[:, None] * logp).sum()


# ----------------------------------------------------------------------------- models
def unpack(theta, spec):
    mv = spec.Mv
    hyp = theta[:NHYP]
    q_mu = theta[NHYP:NHYP + mv]
    Lq = torch.zeros(mv, mv)
    ii = torch.tril_indices(mv, mv)
    Lq[ii[0], ii[1]] = theta[NHYP + mv:]
    return hyp, q_mu, Lq


def gauss_kl_white(q_mu, Lq):
    """gpflow.kullback_leiblers.gauss_kl with K=None (whitened prior)."""
    mahal = (q_mu * q_mu).sum()
    logdet = torch.log(torch.diagonal(Lq) ** 2).sum()
    trace = (torch.tril(Lq) ** 2).sum()
    return 0.5 * (mahal - q_mu.shape[0] - logdet + trace)


def elbo(theta, spec, y, scale=None):
    """ELBO (VGP/SVGP) or log marginal / collapsed bound (GPR/SGPR) as a torch scalar."""
    X = torch.as_tensor(spec.X)
    y = torch.as_tensor(y)
    scale = None if scale is None else torch.as_tensor(scale)
    hyp, q_mu, Lq = unpack(theta, spec)
    var, ls, alpha, km = (softplus(hyp[i]) for i in range(NHYP))
    N = X.shape[0]
    if spec.kind == "VGP":
        K = kernel_K(spec, X, X, var, ls, square=True) + spec.jitter * torch.eye(N)
        L = torch.linalg.cholesky(K)
        fmean = L @ q_mu
        fvar = ((L @ torch.tril(Lq)) ** 2).sum(1)
        ve = variational_expectations(spec, fmean, fvar, y, scale, alpha, km)
        return ve - gauss_kl_white(q_mu, Lq)
    if spec.kind == "SVGP":
        Z = torch.as_tensor(spec.Z)
        M = Z.shape[0]
        Kmm = kernel_K(spec, Z, Z, var, ls, square=True) + spec.jitter * torch.eye(M)
        Kmn = kernel_K(spec, Z, X, var, ls)
        Knn = kernel_Kdiag(spec, X, var)
        Lm = torch.linalg.cholesky(Kmm)
        A = torch.linalg.solve_triangular(Lm, Kmn, upper=False)
        fvar = Knn - (A * A).sum(0)
        fmean = A.T @ q_mu
        LTA = torch.tril(Lq).T @ A
        fvar = fvar + (LTA * LTA).sum(0)
        ve = variational_expectations(spec, fmean, fvar, y, scale, alpha, km)
        return ve - gauss_kl_white(q_mu, Lq)
    noise = GAUSS_VAR_LOWER + alpha                  # slot 2 holds the Gaussian noise variance
    if spec.kind == "GPR":
        K = kernel_K(spec, X, X, var, ls, square=True) + noise * torch.eye(N)
        L = torch.linalg.cholesky(K)
        v = torch.linalg.solve_triangular(L, y[:, None], upper=False)
        return (-0.5 * N * np.log(2 * np.pi) - torch.log(torch.diagonal(L)).sum() - 0.5 * (v * v).sum())
    if spec.kind == "SGPR":
        Z = torch.as_tensor(spec.Z)
        M = Z.shape[0]
        kuf = kernel_K(spec, Z, X, var, ls)
        kuu = kernel_K(spec, Z, Z, var, ls, square=True) + spec.jitter * torch.eye(M)
        kdiag = kernel_Kdiag(spec, X, var)
        sigma = torch.sqrt(noise)
        L = torch.linalg.cholesky(kuu)
        A = torch.linalg.solve_triangular(L, kuf, upper=False) / sigma
        AAT = A @ A.T
        B = AAT + torch.eye(M)
        LB = torch.linalg.cholesky(B)
        Aerr = A @ (y / sigma)[:, None]
        c = torch.linalg.solve_triangular(LB, Aerr, upper=False)
        trace = (kdiag / noise).sum() - torch.diagonal(AAT).sum()
        logdet = -(torch.log(torch.diagonal(LB)).sum() + 0.5 * N * torch.log(noise)) - 0.5 * trace
        quad = -0.5 * ((y * y / noise).sum() - (c * c).sum())
        return -0.5 * N * np.log(2 * np.pi) + logdet + quad
    raise ValueError(spec.kind)


class CholeskyFailure(Exception):
    """Stands for tf.errors.InvalidArgumentError ("Cholesky decomposition was not successful")."""


def elbo_and_grad(theta_np, spec, y, scale=None):
    """(ELBO, dELBO/dtheta) with frozen / unused hyper-parameter slots zeroed in the gradient."""
    theta = torch.tensor(np.asarray(theta_np, dtype=np.float64), requires_grad=True)
    try:
        val = elbo(theta, spec, y, scale)
    except torch.linalg.LinAlgError as e:            # non-PD Gram
        raise CholeskyFailure(str(e))
    (g,) = torch.autograd.grad(val, theta)
    g = g.numpy().copy()
    mask = np.zeros(theta_np.shape[0], dtype=bool)
    mask[spec.trainable_index()] = True
    g[~mask] = 0.0
    return float(val.detach()), g


def initial_theta(spec, var0, ls0, alpha0, km0=35.0):
    """q_mu = 0, q_sqrt = I (gpflow VGP/SVGP defaults) + softplus^-1 of the hyper-parameters."""
    th = np.zeros(spec.P)
    ls_used = ls0 if spec.kernel != "CONST" else 1.0
    th[0] = softplus_inv(var0)
    th[1] = softplus_inv(ls_used)
    th[2] = softplus_inv(alpha0 - (GAUSS_VAR_LOWER if spec.lik == "GAUSS" else 0.0))
    th[3] = softplus_inv(km0)
    mv = spec.Mv
    if mv:
        diag = NHYP + mv + np.array([i * (i + 1) // 2 + i for i in range(mv)])
        th[diag] = 1.0
    return th


def predict_f(theta_np, spec, Xnew, y=None, full_cov=False):
    """gpflow ``model.predict_f(Xnew, full_cov)`` of the fitted model (reference call sites GPcounts_Module.py:930-945 via
    predict_y / predict_f_samples, :305-317): VGP / SVGP through ``conditionals.base_conditional`` (whitened, q_sqrt),
    GPR through the exact Gaussian conditional.  Returns (mean (T,), var (T,) | cov (T, T)) as NumPy arrays."""
    theta = torch.as_tensor(np.asarray(theta_np, dtype=np.float64))
    X = torch.as_tensor(spec.X)
    Xn = torch.as_tensor(np.asarray(Xnew, dtype=np.float64).reshape(len(Xnew), -1))
    hyp, q_mu, Lq = unpack(theta, spec)
    var, ls, alpha, km = (softplus(hyp[i]) for i in range(NHYP))
    T = Xn.shape[0]
    Kss = kernel_K(spec, Xn, Xn, var, ls, square=True) if full_cov else kernel_Kdiag(spec, Xn, var)
    if spec.kind in ("VGP", "SVGP"):
        Zr = X if spec.kind == "VGP" else torch.as_tensor(spec.Z)
        Kmm = kernel_K(spec, Zr, Zr, var, ls, square=True) + spec.jitter * torch.eye(Zr.shape[0])
        Kmn = kernel_K(spec, Zr, Xn, var, ls)
        Lm = torch.linalg.cholesky(Kmm)
        A = torch.linalg.solve_triangular(Lm, Kmn, upper=False)
        mean = A.T @ q_mu
        LTA = torch.tril(Lq).T @ A
        if full_cov:
            cov = Kss - A.T @ A + LTA.T @ LTA
        else:
            cov = Kss - (A * A).sum(0) + (LTA * LTA).sum(0)
        return mean.numpy(), cov.numpy()
    if spec.kind == "GPR":
        noise = GAUSS_VAR_LOWER + alpha
        K = kernel_K(spec, X, X, var, ls, square=True) + noise * torch.eye(X.shape[0])
        L = torch.linalg.cholesky(K)
        A = torch.linalg.solve_triangular(L, kernel_K(spec, X, Xn, var, ls), upper=False)
        v = torch.linalg.solve_triangular(L, torch.as_tensor(y)[:, None], upper=False)
        mean = (A.T @ v)[:, 0]
        cov = Kss - A.T @ A if full_cov else Kss - (A * A).sum(0)
        return mean.numpy(), cov.numpy()
    raise ValueError(spec.kind)
