"""CPU: what the inducing-point policy can change, measured on 2000 genes of the headline workload.

The reference selects RobustGP inducing points per gene under that gene's initial variance (GPcounts_Module.py:673-677);
the product computes ONE set per (X, M, lengthscale) under unit variance (exact arithmetic gives the same points - the
greedy selection ties to the last bit and the winner depends on rounding).  The two oracle runs below (tests/golden/
make_large_fixtures.py: c5_2000 = shared set, c5_2000_refpolicy = the reference's policy) differ from each other exactly
as much as the shared-policy oracle differs from ITSELF under a 1e-12 perturbation of the start: the policy is inside the
round-off chaos of the optimiser, so the shared set stays the default and `inducing_per_gene=True` remains the opt-in."""
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _de(llr):
    from gpcounts_b200 import host
    return host.qvalue(host.p_values(np.maximum(llr, 0))) < 0.05


def test_reference_inducing_policy_is_within_the_oracles_own_noise():
    a = np.load(os.path.join(GOLD, "c5_2000.npz"))
    r = np.load(os.path.join(GOLD, "c5_2000_refpolicy.npz"))
    llr, llr_pert, llr_ref = a["ll"][:, 2], a["ll_pert"][:, 2], r["ll"][:, 2]
    assert (_de(llr) == _de(llr_ref)).mean() >= 0.999          # 1 call of 2000 differs
    assert (_de(llr) == _de(llr_pert)).mean() >= 0.999
    e_policy, e_self = np.abs(llr - llr_ref), np.abs(llr - llr_pert)
    for q in (0.5, 0.9, 0.99):
        assert np.quantile(e_policy, q) <= 1.5 * np.quantile(e_self, q) + 1e-3, (q, np.quantile(e_policy, q), np.quantile(e_self, q))
    tol = 1e-4 * np.abs(llr) + 2e-3
    assert abs((e_policy <= tol).mean() - (e_self <= tol).mean()) <= 0.03


def test_large_fixture_shapes():
    m = np.load(os.path.join(GOLD, "mouseob_455.npz"), allow_pickle=True)
    d = np.load(os.path.join(GOLD, "data_mouseob.npz"), allow_pickle=True)
    assert m["ll"].shape == (455, 3) and d["Y"].shape == (455, 260) and list(m["genes"]) == list(d["genes"])
    # the oracle reproduces the reference's published LLRs on the well-posed class: 325 of 329 to 1e-3 absolute
    llr, tight, paper = m["ll"][:, 2], m["llr_tight"], d["paper_llr"]
    well = np.abs(llr - tight) <= 1e-4 * np.abs(llr)
    assert well.sum() == 329 and (np.abs(llr[well] - paper[well]) <= 1e-3).sum() >= 325
    # ... and the plateau-stopped rest is matched by the tight-tolerance behaviour for most genes
    assert ((np.abs(llr - paper) <= 1e-3) | (np.abs(tight - paper) <= 1e-3)).sum() >= 400
