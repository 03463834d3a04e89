"""Oracle: NumPy restatement of L-BFGS-B for the unbounded case (test infrastructure).

The reference optimises every model with ``gpflow.optimizers.Scipy().minimize`` ->
``scipy.optimize.minimize(method="L-BFGS-B", jac=True, options=dict(maxiter=5000))``
(``GPcounts/GPcounts_Module.py:692-697``).  SciPy (unpinned in ``requirements.txt``) ships L-BFGS-B
v3.0 (Zhu, Byrd, Lu, Nocedal 1997; Morales & Nocedal 2011) as compiled code that is not in
``/root/reference``; this file restates that published algorithm for problems without bounds:

  mainlb   - iteration, convergence tests (``projgr <= pgtol``; ``(fold-f) <= factr*epsmch*max(|fold|,|f|,1)``),
             BFGS pair acceptance ``dr > epsmch*ddum``, memory refresh on failure
  cauchy   - with no bounds and an empty memory the generalised Cauchy point is ``x - g/theta``
  formk    - LEL^T factorisation of the 2col x 2col middle matrix (all variables free)
  subsm    - subspace minimisation = quasi-Newton direction through the compact representation
  matupd / formt - limited-memory matrices WS, WY, SY, SS and T = theta*SS + L D^-1 L^T
  lnsrlb   - first step ``min(1/||d||, stpmx)``, then 1; ``maxls = 20`` trial points
  dcsrch / dcstep - More'-Thuente line search (MINPACK-2), ftol=1e-3, gtol=0.9, xtol=0.1

It exists so that the device optimiser (``gpcounts_b200/csrc/lbfgsb.cuh``) has a CPU twin that is
itself checked, trajectory for trajectory, against the installed SciPy in ``tests/test_oracle_lbfgsb.py``.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

EPSMCH = np.finfo(np.float64).eps
FTOL_LS, GTOL_LS, XTOL_LS = 1e-3, 0.9, 0.1
STPMAX_BIG = 1e10


def dcstep(stx, fx, dx, sty, fy, dy, stp, fp, dp, brackt, stpmin, stpmax):
    sgnd = dp * (dx / abs(dx))
    if fp > fx:                                           # case 1
        theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp
        s = max(abs(theta), abs(dx), abs(dp))
        gamma = s * np.sqrt((theta / s) ** 2 - (dx / s) * (dp / s))
        if stp < stx:
            gamma = -gamma
        p = (gamma - dx) + theta
        q = ((gamma - dx) + gamma) + dp
        r = p / q
        stpc = stx + r * (stp - stx)
        stpq = stx + ((dx / ((fx - fp) / (stp - stx) + dx)) / 2.0) * (stp - stx)
        if abs(stpc - stx) < abs(stpq - stx):
            stpf = stpc
        else:
            stpf = stpc + (stpq - stpc) / 2.0
        brackt = True
    elif sgnd < 0.0:                                      # case 2
        theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp
        s = max(abs(theta), abs(dx), abs(dp))
        gamma = s * np.sqrt((theta / s) ** 2 - (dx / s) * (dp / s))
        if stp > stx:
            gamma = -gamma
        p = (gamma - dp) + theta
        q = ((gamma - dp) + gamma) + dx
        r = p / q
        stpc = stp + r * (stx - stp)
        stpq = stp + (dp / (dp - dx)) * (stx - stp)
        if abs(stpc - stp) > abs(stpq - stp):
            stpf = stpc
        else:
            stpf = stpq
        brackt = True
    elif abs(dp) < abs(dx):                               # case 3
        theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp
        s = max(abs(theta), abs(dx), abs(dp))
        gamma = s * np.sqrt(max(0.0, (theta / s) ** 2 - (dx / s) * (dp / s)))
        if stp > stx:
            gamma = -gamma
        p = (gamma - dp) + theta
        q = (gamma + (dx - dp)) + gamma
        r = p / q
        if r < 0.0 and gamma != 0.0:
            stpc = stp + r * (stx - stp)
        elif stp > stx:
            stpc = stpmax
        else:
            stpc = stpmin
        stpq = stp + (dp / (dp - dx)) * (stx - stp)
        if brackt:
            if abs(stpc - stp) < abs(stpq - stp):
                stpf = stpc
            else:
                stpf = stpq
            if stp > stx:
                stpf = min(stp + 0.66 * (sty - stp), stpf)
            else:
                stpf = max(stp + 0.66 * (sty - stp), stpf)
        else:
            if abs(stpc - stp) > abs(stpq - stp):
                stpf = stpc
            else:
                stpf = stpq
            stpf = min(stpmax, stpf)
            stpf = max(stpmin, stpf)
    else:                                                 # case 4
        if brackt:
            theta = 3.0 * (fp - fy) / (sty - stp) + dy + dp
            s = max(abs(theta), abs(dy), abs(dp))
            gamma = s * np.sqrt((theta / s) ** 2 - (dy / s) * (dp / s))
            if stp > sty:
                gamma = -gamma
            p = (gamma - dp) + theta
            q = ((gamma - dp) + gamma) + dy
            r = p / q
            stpc = stp + r * (sty - stp)
            stpf = stpc
        elif stp > stx:
            stpf = stpmax
        else:
            stpf = stpmin
    if fp > fx:
        sty, fy, dy = stp, fp, dp
    else:
        if sgnd < 0.0:
            sty, fy, dy = stx, fx, dx
        stx, fx, dx = stp, fp, dp
    return stx, fx, dx, sty, fy, dy, stpf, brackt


class DCSRCH:
    """More'-Thuente search as a small state machine: start(f0, g0, stp0) then step(f, g) -> task."""

    def __init__(self, ftol=FTOL_LS, gtol=GTOL_LS, xtol=XTOL_LS, stpmin=0.0, stpmax=STPMAX_BIG):
        self.ftol, self.gtol, self.xtol, self.stpmin, self.stpmax = ftol, gtol, xtol, stpmin, stpmax

    def start(self, f, g, stp):
        if stp < self.stpmin or stp > self.stpmax or g >= 0.0:
            return "ERROR"
        self.brackt = False
        self.stage = 1
        self.finit, self.ginit = f, g
        self.gtest = self.ftol * g
        self.width = self.stpmax - self.stpmin
        self.width1 = self.width / 0.5
        self.stx, self.fx, self.gx = 0.0, f, g
        self.sty, self.fy, self.gy = 0.0, f, g
        self.stmin = 0.0
        self.stmax = stp + 4.0 * stp
        self.stp = stp
        return "FG"

    def step(self, f, g):
        stp = self.stp
        ftest = self.finit + stp * self.gtest
        if self.stage == 1 and f <= ftest and g >= 0.0:
            self.stage = 2
        task = "FG"
        if self.brackt and (stp <= self.stmin or stp >= self.stmax):
            task = "WARNING: ROUNDING ERRORS PREVENT PROGRESS"
        if self.brackt and self.stmax - self.stmin <= self.xtol * self.stmax:
            task = "WARNING: XTOL TEST SATISFIED"
        if stp == self.stpmax and f <= ftest and g <= self.gtest:
            task = "WARNING: STP = STPMAX"
        if stp == self.stpmin and (f > ftest or g >= self.gtest):
            task = "WARNING: STP = STPMIN"
        if f <= ftest and abs(g) <= self.gtol * (-self.ginit):
            task = "CONVERGENCE"
        if task != "FG":
            return task
        if self.stage == 1 and f <= self.fx and f > ftest:
            gt = self.gtest
            fm, fxm, fym = f - stp * gt, self.fx - self.stx * gt, self.fy - self.sty * gt
            gm, gxm, gym = g - gt, self.gx - gt, self.gy - gt
            (self.stx, fxm, gxm, self.sty, fym, gym, stp, self.brackt) = dcstep(
                self.stx, fxm, gxm, self.sty, fym, gym, stp, fm, gm, self.brackt, self.stmin, self.stmax)
            self.fx, self.fy = fxm + self.stx * gt, fym + self.sty * gt
            self.gx, self.gy = gxm + gt, gym + gt
        else:
            (self.stx, self.fx, self.gx, self.sty, self.fy, self.gy, stp, self.brackt) = dcstep(
                self.stx, self.fx, self.gx, self.sty, self.fy, self.gy, stp, f, g, self.brackt,
                self.stmin, self.stmax)
        if self.brackt:
            if abs(self.sty - self.stx) >= 0.66 * self.width1:
                stp = self.stx + 0.5 * (self.sty - self.stx)
            self.width1 = self.width
            self.width = abs(self.sty - self.stx)
        if self.brackt:
            self.stmin = min(self.stx, self.sty)
            self.stmax = max(self.stx, self.sty)
        else:
            self.stmin = stp + 1.1 * (stp - self.stx)
            self.stmax = stp + 4.0 * (stp - self.stx)
        stp = max(stp, self.stpmin)
        stp = min(stp, self.stpmax)
        if (self.brackt and (stp <= self.stmin or stp >= self.stmax)) or \
           (self.brackt and self.stmax - self.stmin <= self.xtol * self.stmax):
            stp = self.stx
        self.stp = stp
        return "FG"


def _chol_upper(A):
    """LINPACK dpofa: A = R^T R, R upper; returns (R, info) with info != 0 on a non-positive pivot."""
    n = A.shape[0]
    R = np.array(A, dtype=np.float64)
    for j in range(n):
        s = 0.0
        for k in range(j):
            t = R[k, j] - np.dot(R[:k, k], R[:k, j])
            t = t / R[k, k]
            R[k, j] = t
            s += t * t
        s = R[j, j] - s
        if not (s > 0.0):
            return R, j + 1
        R[j, j] = np.sqrt(s)
    return np.triu(R), 0


@dataclass
class LbfgsbResult:
    x: np.ndarray
    fun: float
    nit: int
    nfev: int
    success: bool
    status: int
    message: str
    trace: list


def minimize_lbfgsb_unbounded(fun, x0, m=10, ftol=2.220446049250313e-09, gtol=1e-5,
                              maxiter=15000, maxfun=15000, maxls=20, trace=False):
    """fun(x) -> (f, g).  Mirrors scipy.optimize.minimize(method='L-BFGS-B') without bounds."""
    x = np.array(x0, dtype=np.float64)
    n = x.size
    S, Y = [], []                       # chronological correction pairs (WS, WY columns)
    SS = np.zeros((0, 0)); SY = np.zeros((0, 0)); YY = np.zeros((0, 0)); RZ = np.zeros((0, 0))
    theta = 1.0
    f, g = fun(x)
    nfev, nit = 1, 0
    tr = []
    if np.max(np.abs(g)) <= gtol:
        return LbfgsbResult(x, f, 0, nfev, True, 0, "CONVERGENCE: NORM OF PROJECTED GRADIENT <= PGTOL", tr)
    message, success, status = "", False, 2
    while True:
        col = len(S)
        # ---- search direction ------------------------------------------------------------------
        refresh = False
        if col == 0:
            z = x - g / theta                       # cauchy: GCP of the model B = theta*I
        else:
            # formk (all variables free): upper triangle of WN, LEL^T factorisation
            WN = np.zeros((2 * col, 2 * col))
            WN[:col, :col] = np.triu(YY / theta) + np.diag(np.diag(SY))
            WN[:col, col:] = np.tril(RZ.T)         # R_z^T:  WN[jy, col+iy] = s_iy.y_jy for jy >= iy
            R1, info = _chol_upper(WN[:col, :col] + np.triu(WN[:col, :col], 1).T)
            if info == 0:
                X12 = np.zeros((col, col))
                for js in range(col):               # dtrsl job 11: solve R1^T x = b
                    X12[:, js] = _solve_lower(R1.T, WN[:col, col + js])
                B22 = X12.T @ X12                   # + theta * S'AA'S (= 0, no active bounds)
                R2, info = _chol_upper(B22)
            if info != 0:
                refresh = True
            else:
                U = np.zeros((2 * col, 2 * col))
                U[:col, :col], U[:col, col:], U[col:, col:] = R1, X12, R2
                r = -g
                wv = np.concatenate([np.array([np.dot(yk, r) for yk in Y]),
                                     theta * np.array([np.dot(sk, r) for sk in S])])
                wv = _solve_lower(U.T, wv)
                wv[:col] = -wv[:col]
                wv = _solve_upper(U, wv)
                d = r.copy()
                for j in range(col):
                    d += Y[j] * (wv[j] / theta) + S[j] * wv[col + j]
                d *= 1.0 / theta
                z = x + d
        if refresh:
            S, Y = [], []
            SS = np.zeros((0, 0)); SY = np.zeros((0, 0)); YY = np.zeros((0, 0)); RZ = np.zeros((0, 0))
            theta = 1.0
            continue
        d = z - x
        # ---- line search (lnsrlb) -----------------------------------------------------------------
        dtd = np.dot(d, d)
        dnorm = np.sqrt(dtd)
        stp = min(1.0 / dnorm, STPMAX_BIG) if nit == 0 else 1.0
        xold, gold, fold = x.copy(), g.copy(), f
        gd = np.dot(g, d)
        gdold = gd
        ls_failed = False
        ifun = 0
        if gd >= 0.0:
            ls_failed = True                        # "ascent direction in projection", info = -4
        else:
            ls = DCSRCH()
            task = ls.start(f, gd, stp)
            while True:
                stp = ls.stp
                ifun += 1
                nfev += 1
                iback = ifun - 1
                x = z.copy() if stp == 1.0 else stp * d + xold
                f, g = fun(x)
                gd = np.dot(g, d)
                task = ls.step(f, gd)
                if task != "FG":
                    break
                if iback + 1 >= maxls:              # iback >= 20 on the next entry
                    ls_failed = True
                    break
                if nfev > maxfun:
                    break
            if not ls_failed and task.startswith("ERROR"):
                ls_failed = True
        if ls_failed:
            x, g, f = xold, gold, fold
            if col == 0:
                message, status = "ABNORMAL_TERMINATION_IN_LNSRCH", 2
                nit += 1
                break
            S, Y = [], []
            SS = np.zeros((0, 0)); SY = np.zeros((0, 0)); YY = np.zeros((0, 0)); RZ = np.zeros((0, 0))
            theta = 1.0
            continue
        stp = ls.stp
        nit += 1
        if trace:
            tr.append((nit, f, float(np.max(np.abs(g))), stp, ifun))
        # ---- SciPy wrapper limits, then convergence tests ---------------------------------------------
        if nit >= maxiter:
            message, status = "STOP: TOTAL NO. OF ITERATIONS REACHED LIMIT", 1
            break
        if nfev > maxfun:
            message, status = "STOP: TOTAL NO. OF F,G EVALUATIONS EXCEEDS LIMIT", 1
            break
        if np.max(np.abs(g)) <= gtol:
            message, status, success = "CONVERGENCE: NORM OF PROJECTED GRADIENT <= PGTOL", 0, True
            break
        if (fold - f) <= ftol * max(abs(fold), abs(f), 1.0):
            message, status, success = "CONVERGENCE: RELATIVE REDUCTION OF F <= FACTR*EPSMCH", 0, True
            break
        # ---- BFGS update ---------------------------------------------------------------------------------
        r = g - gold
        rr = np.dot(r, r)
        if stp == 1.0:
            dr = gd - gdold
            ddum = -gdold
        else:
            dr = (gd - gdold) * stp
            d = d * stp
            ddum = -gdold * stp
        if dr <= EPSMCH * ddum:
            continue                                # skip the update, keep the memory
        if len(S) == m:
            S.pop(0); Y.pop(0)
            SS, SY, YY, RZ = SS[1:, 1:], SY[1:, 1:], YY[1:, 1:], RZ[1:, 1:]
        S.append(d.copy()); Y.append(r.copy())
        c = len(S)
        nSS = np.zeros((c, c)); nSY = np.zeros((c, c)); nYY = np.zeros((c, c)); nRZ = np.zeros((c, c))
        nSS[:c - 1, :c - 1], nSY[:c - 1, :c - 1], nYY[:c - 1, :c - 1], nRZ[:c - 1, :c - 1] = SS, SY, YY, RZ
        for j in range(c):
            nSS[j, c - 1] = nSS[c - 1, j] = np.dot(S[j], d)
            nYY[j, c - 1] = nYY[c - 1, j] = np.dot(Y[j], r)
            nSY[c - 1, j] = np.dot(d, Y[j])          # s_new . y_j   (lower part of SY, matupd)
            nRZ[j, c - 1] = np.dot(S[j], r)          # s_j . y_new   (new R_z column, formk; diagonal included)
        nSS[c - 1, c - 1] = dtd if stp == 1.0 else stp * stp * dtd
        nSY[c - 1, c - 1] = dr
        SS, SY, YY, RZ = nSS, nSY, nYY, nRZ
        theta = rr / dr
        # formt: T = theta*SS + L D^-1 L^T must be positive definite
        Lm = np.tril(SY, -1)
        T = theta * SS + Lm @ np.diag(1.0 / np.diag(SY)) @ Lm.T
        _, info = _chol_upper(T)
        if info != 0:
            S, Y = [], []
            SS = np.zeros((0, 0)); SY = np.zeros((0, 0)); YY = np.zeros((0, 0)); RZ = np.zeros((0, 0))
            theta = 1.0
    return LbfgsbResult(x, f, nit, nfev, success, status, message, tr)


def _solve_lower(L, b):
    n = L.shape[0]
    x = np.array(b, dtype=np.float64)
    for i in range(n):
        x[i] = (x[i] - np.dot(L[i, :i], x[:i])) / L[i, i]
    return x


def _solve_upper(U, b):
    n = U.shape[0]
    x = np.array(b, dtype=np.float64)
    for i in range(n - 1, -1, -1):
        x[i] = (x[i] - np.dot(U[i, i + 1:], x[i + 1:])) / U[i, i]
    return x
