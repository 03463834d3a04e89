"""CPU: accuracy of the branch-free FP64 exp / log / lgamma / digamma used in the kernels' inner loops
(gpcounts_b200/csrc/fastmath.cuh is host-compilable; tests/native/fastmath_check.cpp compares it with long-double libm)."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fastmath_accuracy(tmp_path):
    exe = str(tmp_path / "fastmath_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "native", "fastmath_check.cpp")])
    out = subprocess.check_output([exe], text=True)
    v = {k: float(x) for k, x in re.findall(r"(\w+) ([0-9.e+-]+)", out.split("special")[0])}
    assert v["exp_ulp"] <= 1.0 and v["exp_nonpos_ulp"] <= 1.0, out
    assert v["log_ulp"] <= 2.5, out
    assert v["lgamma_err"] <= 2e-14 and v["digamma_err"] <= 1e-14, out
    assert v["vec_ok"] == 1
    assert "exp(800)=inf exp(-800)=0" in out and "log(inf)=inf" in out and "exp(0)=1 log(1)=0" in out
