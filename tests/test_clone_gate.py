"""Build-time guard (no GPU needed): the register frame ptxas gives the fused evaluator under `k_fit`.

DESIGN.md section 3: the clone of `eval_svgp_fused` that `k_fit` reaches reacts to unrelated edits of the kernel (a
batched dot-product pass in the optimiser cost the evaluator 25 % of its speed).  The accepted frame consists of the
ABI entry/exit saves only.  If this test fails after a change, measure `python bench.py --steps 1 --warmup 1 --no-cpu
` (`time_split.us_per_eval_in_fit`, ~250 us for the accepted frame) and either undo the change or update
`tools/clone_gate.py::ACCEPTED` together with the measurement in DESIGN.md."""
import importlib.util
import os
import shutil

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="needs nvcc")
def test_fused_evaluator_frame_under_k_fit():
    spec = importlib.util.spec_from_file_location("clone_gate", os.path.join(ROOT, "tools", "clone_gate.py"))
    gate = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gate)
    fr = gate.frames()
    assert "k_fit" in fr, fr
    assert fr["k_fit"] == gate.ACCEPTED, (fr, gate.ACCEPTED)
