"""Torch-tensor front end of the C ABI: device buffers, streams and workspace live in PyTorch, every
number is produced by the sm_100a kernels behind ``libgpcounts_b200.so`` (no CPU or eager fallback)."""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np
import torch

from . import _lib as L

KIND = {"VGP": L.GPC_VGP, "SVGP": L.GPC_SVGP, "GPR": L.GPC_GPR, "SGPR": L.GPC_SGPR}
KERNEL = {"RBF": L.GPC_RBF, "CONST": L.GPC_CONST, "BRANCH": L.GPC_BRANCH}
LIK = {"NB": L.GPC_NB, "ZINB": L.GPC_ZINB, "POISSON": L.GPC_POISSON, "GAUSS": L.GPC_GAUSS}


_pinned = {"buf": None, "event": None}


def _pinned_stage(nbytes):
    """Grow-only pinned staging buffer, reused across calls (a fresh cudaHostAlloc per call costs milliseconds);
    the previous asynchronous copy out of it is awaited before it is overwritten."""
    if _pinned["event"] is not None:
        _pinned["event"].synchronize()
        _pinned["event"] = None
    buf = _pinned["buf"]
    if buf is None or buf.numel() < nbytes:
        _pinned["buf"] = buf = torch.empty(int(nbytes), dtype=torch.uint8).pin_memory()
    return buf


def _dev_f64(a, device):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        a = np.ascontiguousarray(a, dtype=np.float64)
        if torch.device(device).type == "cuda" and a.size >= (1 << 16):
            stage = _pinned_stage(a.nbytes)[: a.nbytes].view(torch.float64).view(a.shape)
            stage.copy_(torch.from_numpy(a))        # H2D from pinned host memory
            out = stage.to(device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(device))
            _pinned["event"] = ev
            return out
        return torch.from_numpy(a).to(device)
    t = torch.as_tensor(a, dtype=torch.float64)
    return t.to(device).contiguous()


class Model:
    """One model of the per-gene chain; owns the device copies of X / Z / restart table it points to."""

    def __init__(self, kind, kernel, lik, X, Z=None, n_off=0, train=(True, True, True, True), handoff_alpha=False,
                 xp=-1000.0, jitter=1e-6, restart=None, Zgene=None, device="cuda"):
        self.kind, self.kernel, self.lik = kind, kernel, lik
        X = np.asarray(X, dtype=np.float64)
        self.X = _dev_f64(X.reshape(X.shape[0], -1), device)
        self.N, self.d = self.X.shape
        self.Z = None
        self.M, self.n_zsets = 0, 0
        if Z is not None:
            Z = np.asarray(Z, dtype=np.float64)
            if Z.ndim == 2:
                Z = Z[None]
            self.Z = _dev_f64(Z, device)            # (n_zsets, M, d)
            self.n_zsets, self.M = Z.shape[0], Z.shape[1]
        self.restart = _dev_f64(restart, device)    # (10, 4) = (ls, var, alpha, km) or None
        self.Zgene = _dev_f64(Zgene, device)        # (D, M, d) per-gene inducing points of attempt 0, or None
        self.n_off = int(n_off)
        self.train_mask = sum(1 << i for i, t in enumerate(train) if t)
        self.handoff_alpha = bool(handoff_alpha)
        self.xp, self.jitter = float(xp), float(jitter)

    @property
    def Mv(self):
        return self.N if self.kind == "VGP" else (self.M if self.kind == "SVGP" else 0)

    @property
    def P(self):
        mv = self.Mv
        return L.GPC_NHYP + mv + mv * (mv + 1) // 2

    def c_struct(self):
        m = L.gpc_model_t()
        m.kind, m.kernel, m.lik = KIND[self.kind], KERNEL[self.kernel], LIK[self.lik]
        m.N, m.M, m.d, m.n_off = self.N, self.M, self.d, self.n_off
        m.train_mask, m.n_zsets, m.handoff_alpha = self.train_mask, max(self.n_zsets, 1), int(self.handoff_alpha)
        m.xp, m.jitter = self.xp, self.jitter
        m.X = self.X.data_ptr()
        m.Z = self.Z.data_ptr() if self.Z is not None else None
        m.Zgene = self.Zgene.data_ptr() if self.Zgene is not None else None
        m.restart = self.restart.data_ptr() if self.restart is not None else None
        return m


def default_opt(**over):
    lib = L.load()
    o = L.gpc_opt_t()
    lib.gpc_default_opt(C.byref(o))
    for k, v in over.items():
        setattr(o, k, v)
    return o


def _ld(t):
    """Leading dimension of a 2-D/3-D device tensor; size-1 leading dims can carry arbitrary strides in torch."""
    return int(t.stride(0)) if t.shape[0] > 1 else int(np.prod(t.shape[1:]))


def _stream_ptr(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _models_array(models: Sequence[Model]):
    arr = (L.gpc_model_t * len(models))()
    for i, m in enumerate(models):
        arr[i] = m.c_struct()
    return arr


_ws_cache = {}


def _workspace(nbytes, device):
    """Grow-only per-device workspace (caller-owned memory of the C ABI)."""
    key = torch.device(device).index or 0
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = None
        _ws_cache[key] = None
        buf = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


def default_slots(models):
    lib = L.load()
    arr = _models_array(models)
    return int(lib.gpc_default_slots(arr, len(models)))


def elbo_grad(model: Model, Y, scale, theta, zset=0, n_slots=None):
    """Batched ELBO and gradient at fixed theta.  Y, scale: (D, ldy) device tensors; theta (D, P)."""
    lib = L.load()
    dev = model.X.device
    Y = _dev_f64(Y, dev)
    scale = _dev_f64(scale, dev)
    theta = _dev_f64(theta, dev)
    D = Y.shape[0]
    assert theta.shape == (D, model.P), (theta.shape, model.P)
    arr = _models_array([model])
    n_slots = n_slots or min(default_slots([model]), max(D, 1))
    nbytes = lib.gpc_workspace_bytes(arr, 1, None, n_slots)
    ws = _workspace(nbytes, dev)
    elbo = torch.empty(D, dtype=torch.float64, device=dev)
    grad = torch.empty(D, model.P, dtype=torch.float64, device=dev)
    status = torch.empty(D, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        L.check(lib.gpc_elbo_grad(arr, D, Y.data_ptr(), scale.data_ptr() if scale is not None else None, _ld(Y),
                                  theta.data_ptr(), zset, elbo.data_ptr(), grad.data_ptr(), status.data_ptr(),
                                  ws.data_ptr(), ws.numel(), n_slots, _stream_ptr(dev)))
    return elbo, grad, status


def fit(models: Sequence[Model], Y, scale, hyp0, chain=True, opt=None, return_theta=True, n_slots=None,
        max_ws_bytes=None):
    """Batched fit of the chain of models for every gene.  Returns (stats (D,K,16), theta (D,K,Pmax) | None)."""
    lib = L.load()
    dev = models[0].X.device
    Y = _dev_f64(Y, dev)
    scale = _dev_f64(scale, dev)
    hyp0 = _dev_f64(hyp0, dev)
    D, K = Y.shape[0], len(models)
    assert hyp0.shape == (D, K, 4), hyp0.shape
    opt = opt or default_opt()
    arr = _models_array(models)
    n_jobs = D if chain else D * K
    n_slots = n_slots or min(default_slots(models), max(n_jobs, 1))
    if max_ws_bytes is None:        # 80 % of what is free now, plus what the cached workspace already holds
        free, _total = torch.cuda.mem_get_info(dev)
        cached = _ws_cache.get(torch.device(dev).index or 0)
        max_ws_bytes = int(free * 0.8) + (cached.numel() if cached is not None else 0)
    while n_slots > 1 and lib.gpc_workspace_bytes(arr, K, C.byref(opt), n_slots) > max_ws_bytes:
        n_slots = max(1, n_slots // 2)
    nbytes = lib.gpc_workspace_bytes(arr, K, C.byref(opt), n_slots)
    ws = _workspace(nbytes, dev)
    Pmax = max(m.P for m in models)
    stats = torch.empty(D, K, L.GPC_NSTAT, dtype=torch.float64, device=dev)
    theta = torch.empty(D, K, Pmax, dtype=torch.float64, device=dev) if return_theta else None
    with torch.cuda.device(dev):
        L.check(lib.gpc_fit(arr, K, int(chain), C.byref(opt), D, Y.data_ptr(),
                            scale.data_ptr() if scale is not None else None, _ld(Y), hyp0.data_ptr(),
                            stats.data_ptr(), theta.data_ptr() if theta is not None else None, Pmax,
                            ws.data_ptr(), ws.numel(), n_slots, _stream_ptr(dev)))
    return stats, theta


def predict(model: Model, theta, Xnew, Y=None, zset=0, full_cov=False, n_slots=None):
    """Posterior mean / variance (and optionally covariance) of f at Xnew for D fitted parameter vectors theta (D, >= P)."""
    lib = L.load()
    dev = model.X.device
    theta = _dev_f64(theta, dev).contiguous()
    Xnew = _dev_f64(np.asarray(Xnew, dtype=np.float64).reshape(len(Xnew), -1), dev)
    Y = _dev_f64(Y, dev)
    D, T = theta.shape[0], Xnew.shape[0]
    assert Xnew.shape[1] == model.d and theta.shape[1] >= model.P
    arr = _models_array([model])
    n_slots = n_slots or min(296, max(D, 1))
    nbytes = lib.gpc_predict_workspace_bytes(arr, T, n_slots)
    ws = _workspace(nbytes, dev)
    mean = torch.empty(D, T, dtype=torch.float64, device=dev)
    var = torch.empty(D, T, dtype=torch.float64, device=dev)
    cov = torch.empty(D, T, T, dtype=torch.float64, device=dev) if full_cov else None
    with torch.cuda.device(dev):
        L.check(lib.gpc_predict(arr, D, Y.data_ptr() if Y is not None else None, _ld(Y) if Y is not None else 0,
                                theta.data_ptr(), _ld(theta), zset, Xnew.data_ptr(), T, mean.data_ptr(), var.data_ptr(),
                                cov.data_ptr() if cov is not None else None, ws.data_ptr(), ws.numel(), n_slots,
                                _stream_ptr(dev)))
    return mean, var, cov


def dbg_linalg(op, A, B=None, nrhs=0):
    """Test hook for the CTA-level dense primitives (see include/gpcounts_b200.h)."""
    lib = L.load()
    dev = A.device
    batch, n = A.shape[0], A.shape[1]
    status = torch.empty(batch, dtype=torch.int32, device=dev)
    scratch = torch.empty(int(lib.gpc_dbg_linalg_scratch_bytes(batch, n)), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        L.check(lib.gpc_dbg_linalg(op, batch, n, nrhs, A.data_ptr(), B.data_ptr() if B is not None else None,
                                   status.data_ptr(), scratch.data_ptr(), scratch.numel(), _stream_ptr(dev)))
        torch.cuda.synchronize(dev)         # the scratch buffer must outlive the (stream-ordered) kernel
    return status


def launch_count():
    return int(L.load().gpc_launch_count())
