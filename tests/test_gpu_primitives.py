"""CTA-level dense primitives (blocked Cholesky, triangular solves, Cholesky reverse pass) vs torch.linalg FP64."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _spd(batch, n, seed):
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(batch, n, n, generator=g, dtype=torch.float64)
    return (A @ A.transpose(1, 2) / n + 0.5 * torch.eye(n, dtype=torch.float64)).cuda()


@pytest.mark.parametrize("n", [1, 7, 18, 64, 65, 100, 130, 260])
def test_potrf(n):
    from gpcounts_b200 import engine
    A = _spd(5, n, n)
    ref = torch.linalg.cholesky(A)
    out = A.clone()
    st = engine.dbg_linalg(0, out)
    assert int(st.abs().sum()) == 0
    err = (torch.tril(out) - ref).abs().max() / ref.abs().max()
    assert err < 1e-13, err


def test_potrf_not_pd():
    from gpcounts_b200 import engine
    A = _spd(3, 70, 1)
    A[1, 40, 40] = -5.0
    st = engine.dbg_linalg(0, A.clone())
    assert st.tolist() == [0, 1, 0]


@pytest.mark.parametrize("n,nrhs", [(18, 5), (64, 100), (100, 37), (133, 260)])
def test_trsm(n, nrhs):
    from gpcounts_b200 import engine
    A = _spd(4, n, n + 1)
    Lr = torch.linalg.cholesky(A)
    g = torch.Generator().manual_seed(n)
    B = torch.randn(4, n, nrhs, generator=g, dtype=torch.float64).cuda()
    for op, ref in [(1, torch.linalg.solve_triangular(Lr, B, upper=False)),
                    (2, torch.linalg.solve_triangular(Lr.transpose(1, 2), B, upper=True))]:
        out = B.clone()
        engine.dbg_linalg(op, A.clone(), out, nrhs)
        assert ((out - ref).abs().max() / ref.abs().max()) < 1e-12, op
    Bt = B.transpose(1, 2).contiguous()             # (nrhs x n) <- Bt L^-1
    ref = torch.linalg.solve_triangular(Lr, Bt, upper=False, left=False)
    out = Bt.clone()
    engine.dbg_linalg(3, A.clone(), out, nrhs)
    assert ((out - ref).abs().max() / ref.abs().max()) < 1e-12


@pytest.mark.parametrize("n", [18, 64, 100, 150])
def test_chol_backward(n):
    from gpcounts_b200 import engine
    A = _spd(3, n, n + 2).requires_grad_(True)
    Lr = torch.linalg.cholesky(A)
    g = torch.Generator().manual_seed(n)
    Lbar = torch.tril(torch.randn(3, n, n, generator=g, dtype=torch.float64)).cuda()
    (Abar,) = torch.autograd.grad((Lr * Lbar).sum(), A)
    out = Lbar.clone()
    engine.dbg_linalg(4, A.detach().clone(), out)
    sym = 0.5 * (out + out.transpose(1, 2))
    refsym = 0.5 * (Abar + Abar.transpose(1, 2))
    assert ((sym - refsym).abs().max() / refsym.abs().max()) < 1e-11
