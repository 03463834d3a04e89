"""CPU: the C-ABI library loads and exports every symbol include/gpcounts_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    from gpcounts_b200 import _lib
    return _lib.load()


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "gpcounts_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gpc_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_are_exported(lib):
    names = _declared_functions()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), n
    from gpcounts_b200 import _lib
    assert sorted(_lib.EXPORTS) == names


def test_struct_layout_matches_header(lib):
    from gpcounts_b200 import _lib
    # 10 int32 + 2 double + 4 pointers, natural alignment
    assert ctypes.sizeof(_lib.gpc_model_t) == 10 * 4 + 2 * 8 + 4 * 8
    # 4 int32 + 2 double + 4 int32 (variant, trace_cap, trace_gene, trace_model) + 2 pointers (trace, attempt0)
    assert ctypes.sizeof(_lib.gpc_opt_t) == 4 * 4 + 2 * 8 + 4 * 4 + 2 * 8


def test_defaults_and_param_count(lib):
    from gpcounts_b200 import _lib
    o = _lib.gpc_opt_t()
    lib.gpc_default_opt(ctypes.byref(o))
    assert (o.m, o.maxiter, o.maxfun, o.maxls) == (10, 5000, 15000, 20)
    assert o.ftol == 2.220446049250313e-09 and o.gtol == 1e-5
    assert o.variant == -1 and not o.trace and o.trace_cap == 0
    m = _lib.gpc_model_t()
    m.kind, m.N, m.M = _lib.GPC_SVGP, 1000, 64
    assert lib.gpc_param_count(ctypes.byref(m)) == 4 + 64 + 64 * 65 // 2
    m.kind = _lib.GPC_VGP
    m.N = 260
    assert lib.gpc_param_count(ctypes.byref(m)) == 4 + 260 + 260 * 261 // 2
    assert b"sm_100a" in lib.gpc_version()


def test_argument_errors_are_reported_without_a_gpu(lib):
    from gpcounts_b200 import _lib
    m = _lib.gpc_model_t()
    o = _lib.gpc_opt_t()
    lib.gpc_default_opt(ctypes.byref(o))
    nbytes = lib.gpc_workspace_bytes(ctypes.byref(m), 1, ctypes.byref(o), 4)
    assert nbytes > 0


def test_library_reads_no_environment_and_keeps_no_debug_globals():
    """trace / variant overrides travel in gpc_opt_t, not in globals or environment variables."""
    for f in os.listdir(os.path.join(ROOT, "gpcounts_b200", "csrc")):
        txt = open(os.path.join(ROOT, "gpcounts_b200", "csrc", f)).read()
        assert "getenv" not in txt and "cudaMalloc(" not in txt and "cudaStreamSynchronize" not in txt, f


def test_no_cpu_fallback():
    """Without a CUDA device the product path raises instead of computing on the CPU."""
    import numpy as np
    import pandas as pd
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from gpcounts_b200 import Fit_GPcounts
    X = pd.DataFrame(np.linspace(0, 1, 10), index=["c%d" % i for i in range(10)])
    Y = pd.DataFrame(np.ones((2, 10)), index=["a", "b"], columns=X.index)
    gp = Fit_GPcounts(X, Y)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        gp.One_sample_test()


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "gpcounts_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
