"""CPU oracle for the GPcounts per-gene GP fitting hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``gpcounts_b200/`` may import this package; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs
do, and there only as the checker or the timed CPU baseline - never as part of the product path.

The reference (``/root/reference``, ManchesterBioinference/GPcounts) is ~2.5 kLoC of Python glue over
GPflow / TensorFlow / SciPy / RobustGP, none of which is vendored, pinned or importable in the build
container.  The arithmetic of the hot path therefore lives in third-party
packages that are absent; this oracle *restates* their published algorithms:

  * GPflow (unpinned, >=2.6 implied by ``check_shapes``): ``VGP.elbo``, ``SVGP.elbo`` with
    ``base_conditional`` + ``gauss_kl`` (whitened), ``GPR.log_marginal_likelihood``, ``SGPR.elbo``,
    RBF / Constant kernels, 20-point Gauss-Hermite ``ScalarLikelihood`` quadrature, ``Poisson``
    closed form, softplus / fill-triangular parameter transforms.
  * GPcounts' own pieces: ``NegativeBinomialLikelihood.py:11-89``, ``branchingKernel.py:31-74`` and the
    orchestration in ``GPcounts_Module.py:393-806, 243-366, 1008-1063``.
  * SciPy L-BFGS-B (``gpflow.optimizers.Scipy`` default) - the installed SciPy is used directly as the
    optimiser of record; ``oracle/lbfgsb.py`` additionally restates the unbounded L-BFGS-B iteration
    (Byrd-Lu-Nocedal-Zhu / Morales-Nocedal v3.0) so the device optimiser has a line-by-line CPU twin.
  * RobustGP ``ConditionalVariance`` greedy inducing-point selection (git master, unpinned).

Parity pins (the reference has no tests; these are its notebook outputs / result tables, see
``tests/golden/``): four VGP log-likelihoods for Slc1a3 / 2010300C02Rik
(``demo_notebooks/GPcounts_spatial.ipynb:1080,1086``), the 20-row p/q table (``:1334-1376``),
``Z[0]=0.71943149`` for RobustGP on ``data/alpha_time_points.csv``
(``demo_notebooks/scRNA-Seq_time_series.ipynb:670``) and the ``SE_genes_comparison.csv`` LLRs.
Sparse SVGP numbers, Poisson, ZINB, Gaussian, two-sample and branching results are NOT pinned by the
reference ("parity unpinned" for those): there this restatement is the sole authority.
"""
