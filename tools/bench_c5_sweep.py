"""C5 sweep shapes of BASELINE.json configs[4] beyond the headline (N in {1000, 10000}, M in {64, 256}):

    python tools/bench_c5_sweep.py N M D [maxiter]

Synthetic genes from bench.synth_batch, sparse NB One_sample_test through Fit_GPcounts; prints genes/s,
ELBO+gradient evaluations/s and algorithmic FP64 TFLOP/s (6 M^2 N + 4/3 M^3 + 12 M N per evaluation)."""
import sys, time, numpy as np, pandas as pd
sys.path.insert(0, '/root/repo')
from bench import synth_batch
from gpcounts_b200 import Fit_GPcounts
N, M, D = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
maxiter = int(sys.argv[4]) if len(sys.argv) > 4 else 0
X, Y = synth_batch(D, N, seed=1234)
cells = ["c%d" % i for i in range(N)]
gp = Fit_GPcounts(pd.DataFrame(X, index=cells), pd.DataFrame(Y, index=["g%d" % i for i in range(D)], columns=cells),
                  sparse=True, M=M)
if maxiter: gp.optimizer_options = {"maxiter": maxiter}
t = time.time(); df = gp.One_sample_test("Negative_binomial"); dt = time.time() - t
st = gp.fit_stats
nfev = st.n_feval_total.sum()
flops = nfev * (6.0 * M * M * N + 4.0 / 3.0 * M ** 3 + 12.0 * M * N)
print("C5 N=%d M=%d D=%d maxiter=%s: %.2f s wall, kernel %.2f s -> %.1f genes/s, %.0f evals/s, %.2f TFLOP/s algorithmic, ok %.3f, evals/gene %.0f"
      % (N, M, D, maxiter or "default", dt, gp.last_fit_seconds, D / gp.last_fit_seconds, nfev / gp.last_fit_seconds,
         flops / gp.last_fit_seconds * 1e-12, (st.status == 0).mean(), nfev / D))
