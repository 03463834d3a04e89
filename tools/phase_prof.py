"""Phase breakdown of the fused SVGP evaluator (C5 shape) from SM-clock timers compiled in with -DGPC_PHASE_PROF.

    python tools/phase_prof.py build      # here: compiles gpcounts_b200/lib/libgpcounts_b200_prof.so (instantiation t256 only)
    python tools/phase_prof.py            # on the GPU: runs a fit of 296 C5 genes and prints cycles per evaluation by phase

The profiling library is a separate file; the product library has no timers."""
import ctypes as C, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PROF = os.path.join(ROOT, "gpcounts_b200", "lib", "libgpcounts_b200_prof.so")
if len(sys.argv) > 1 and sys.argv[1] == "build":
    src = [os.path.join(ROOT, "gpcounts_b200", "csrc", f) for f in ("gpc_api.cu", "gpc_dense.cu", "gpc_fused.cu", "gpc_small.cu")]
    flags = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-DGPC_PHASE_PROF"]
    objs = []
    procs = []
    for s in src:
        o = s.replace(".cu", "_prof.o").replace("csrc", "lib")
        objs.append(o)
        procs.append(subprocess.Popen(["nvcc", *flags, "-c", s, "-o", o]))
    assert all(p.wait() == 0 for p in procs)
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", PROF, *objs])
    sys.exit(0)
import numpy as np, torch
from gpcounts_b200 import _lib
_lib.LIB_PATH = PROF
from bench import synth_batch
from gpcounts_b200 import engine, host
N, M, D = 1000, 64, int(sys.argv[1]) if len(sys.argv) > 1 else 296
X, Y = synth_batch(D, N, seed=1)
ls0 = host.default_lengthscale(X)
Z = host.inducing_sets(X, M, ls0); rt, _ = host.restart_table(X)
models = [engine.Model("SVGP", "RBF", "NB", X, Z=Z, restart=rt), engine.Model("SVGP", "CONST", "NB", X, Z=Z, restart=rt, handoff_alpha=True)]
h = np.zeros((D, 2, 4)); h[:, :, 0] = host.default_variance(Y, "Negative_binomial", True)[:, None]; h[:, :, 1] = ls0; h[:, :, 2] = 1; h[:, :, 3] = 35
lib = _lib.load()
lib.gpc_dbg_phase_prof.argtypes = [C.c_int, C.c_void_p]
out = (C.c_ulonglong * 24)()
names = ["setup(Z,Lq,Kuu,W)", "cholesky", "tri-inverse", "Lm save,u,lgamma pre-pass", "tile load + Kuf exp", "A = Linv Kuf", "C = W A",
         "stats + quadrature", "rank update H += A D A^T", "end phase", "S product Linv^T C", "S epilogue + gq"]
for variant in (-1, 0):
    st, _ = engine.fit(models, Y, None, h, chain=True, opt=engine.default_opt(variant=variant), return_theta=False)
    torch.cuda.synchronize()
    assert lib.gpc_dbg_phase_prof(2 if variant < 0 else 0, out) == 0
    print('instantiation', 'f256' if variant < 0 else 't256')
    n = out[15]
    tot = sum(out[i] for i in range(12))
    print("evaluations %d, cycles per evaluation %.0f (%.1f us at 1965 MHz)" % (n, tot / n, tot / n / 1965))
    for i, nm in enumerate(names):
        print("  %-32s %8.0f cycles  %5.1f %%" % (nm, out[i] / n, 100.0 * out[i] / tot))
    nit = max(out[23], 1)
    onames = ["direction pass", "trial point x = xold + stp d", "gradient pass + line-search step", "BFGS update (s, y stored)", "dot products with the memory", "small matrices (formt / formk / subsm)"]
    print("  optimiser, cycles per iteration (%d iterations):" % nit)
    for i, nm in enumerate(onames):
        print("    %-40s %8.0f" % (nm, out[16 + i] / nit))
    s = st.cpu().numpy()
    print("  fit cycles per evaluation %.0f, eval share %.3f" % (s[..., 13].sum() / s[..., 9].sum(), s[..., 12].sum() / s[..., 13].sum()))
