"""Oracle: p-values and Storey q-values (test infrastructure).

Follows ``Fit_GPcounts.calculate_FDR`` (``GPcounts_Module.py:368-374``) and ``Fit_GPcounts.qvalue``
(``GPcounts_Module.py:1008-1063``).  The reference calls the NumPy aliases ``sp.arange / sp.array /
sp.argsort / sp.zeros_like`` that current SciPy no longer exports; they are
restated with NumPy, which is what they were.
"""
import numpy as np
import scipy.stats as ss
from scipy import interpolate


def qvalue(pv, pi0=None):
    pv = np.asarray(pv, dtype=np.float64)
    assert pv.min() >= 0 and pv.max() <= 1, "p-values should be between 0 and 1"
    original_shape = pv.shape
    pv = pv.ravel()
    m = float(len(pv))
    if len(pv) < 100 and pi0 is None:                      # :1024-1025
        pi0 = 1.0
    elif pi0 is not None:
        pi0 = pi0
    else:                                                  # :1029-1043
        pi0 = []
        lam = np.arange(0, 0.90, 0.01)
        counts = np.array([(pv > i).sum() for i in np.arange(0, 0.9, 0.01)])
        for l in range(len(lam)):
            pi0.append(counts[l] / (m * (1 - lam[l])))
        pi0 = np.array(pi0)
        tck = interpolate.splrep(lam, pi0, k=3)
        pi0 = interpolate.splev(lam[-1], tck)
        if pi0 > 1:
            pi0 = 1.0
    assert pi0 >= 0 and pi0 <= 1, "pi0 is not between 0 and 1: %f" % pi0
    p_ordered = np.argsort(pv)                             # :1047
    pv = pv[p_ordered]
    qv = pi0 * m / len(pv) * pv
    qv[-1] = min(qv[-1], 1.0)
    for i in range(len(pv) - 2, -1, -1):                   # :1052-1053
        qv[i] = min(pi0 * m * pv[i] / (i + 1.0), qv[i + 1])
    qv_temp = qv.copy()
    qv = np.zeros_like(qv)
    qv[p_ordered] = qv_temp
    return qv.reshape(original_shape)


def calculate_FDR(llr):
    """p = 1 - chi2_1.cdf(2 LLR) (:369-371), q = qvalue(p) (:372)."""
    p = 1 - ss.chi2.cdf(2 * np.asarray(llr, dtype=np.float64), df=1)
    return p, qvalue(p)
