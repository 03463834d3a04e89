"""Clone gate for the fused evaluator ("How the evaluators are called").

ptxas compiles a separate clone of `eval_svgp_fused` for every kernel that reaches it, and the register allocation of
the clone under `k_fit` reacts to unrelated edits of the optimiser.  The known-good frame is the one that consists of
the ABI entry/exit saves only; this script recompiles gpc_fused.cu with `-Xptxas -v` and compares.

    python tools/clone_gate.py            # exit 0 if the k_fit clone has the accepted frame

Accepted (round 2, column-owner tile loop, 1.24-1.27 k genes/s, 248 us per evaluation inside the fit): 552 B stack, 520 B spill
stores, 520 B spill loads - the ABI saves only, nothing spilled inside the tile loop (earlier in round 2: 720 / 972 / 972 at 292 us;
round 1: 656 / 740 / 844 at 386 us).  A different frame is not necessarily slower - measure
(`python bench.py --steps 1 --warmup 1 --no-cpu`, time_split.us_per_eval_in_fit) and update ACCEPTED if the new one is kept.
"""
import os, re, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ACCEPTED = (552, 520, 520)


def frames():
    with tempfile.TemporaryDirectory() as tmp:
        cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
               "-Xptxas", "-v", "-c", os.path.join(ROOT, "gpcounts_b200", "csrc", "gpc_fused.cu"), "-o", os.path.join(tmp, "f.o")]
        out = subprocess.run(cmd, capture_output=True, text=True).stderr
    res, kernel, fn = {}, None, None
    for line in out.splitlines():
        m = re.search(r"Compiling entry function '(\S+)'", line)
        if m:
            kernel = "k_fit" if "k_fit" in m.group(1) else ("k_elbo_grad" if "k_elbo_grad" in m.group(1) else m.group(1))
            continue
        m = re.search(r"Function properties for (\S+)", line)
        if m:
            fn = m.group(1)
            continue
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m and fn and "eval_svgp_fused" in fn:
            res[kernel] = tuple(int(x) for x in m.groups())
            fn = None
    return res


if __name__ == "__main__":
    fr = frames()
    for k, v in fr.items():
        print("eval_svgp_fused under %-12s: %d B stack, %d B spill stores, %d B spill loads" % ((k,) + v))
    ok = fr.get("k_fit") == ACCEPTED
    print("k_fit clone:", "accepted frame" if ok else "DIFFERENT from the accepted frame %s - measure before keeping" % (ACCEPTED,))
    sys.exit(0 if ok else 1)
