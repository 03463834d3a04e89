"""Write a markdown summary of an .ncu-rep capture of one of our kernels: headline counters, stall reasons and
the source lines that collect the most warp-stall samples (SASS samples mapped to CUDA lines through nvdisasm).

    python tools/ncu_summary.py <rep> <out.md> "<title / command>"
"""
import csv, io, os, re, subprocess, sys, collections, tempfile

rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "gpcounts_b200", "lib", "libgpcounts_b200.so")

def run(cmd, **kw):
    return subprocess.run(cmd, capture_output=True, text=True, **kw).stdout

raw = list(csv.reader(io.StringIO(run(["ncu", "-i", rep, "--page", "raw", "--csv"]))))
hdr, units, vals = raw[0], raw[1], raw[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum"]
tmp = tempfile.mkdtemp()
run(["cuobjdump", "-xelf", "all", lib], cwd=tmp)
dis = ""
for f in sorted(os.listdir(tmp)):
    if f.endswith(".cubin"):
        dis += run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)])
src = list(csv.reader(io.StringIO(run(["ncu", "-i", rep, "--page", "source", "--csv"]))))
kname = src[0][1] if src and len(src[0]) > 1 else "?"
h2 = src[1]
reasons = ["stall_wait", "stall_short_sb", "stall_barrier", "stall_math", "stall_long_sb", "stall_branch_resolving",
           "stall_selected", "stall_not_selected", "stall_no_inst", "stall_mio"]
ia, ismp = h2.index("Address"), h2.index("# Samples")
ridx = [h2.index(r) for r in reasons]
ins = [(int(r[ismp] or 0), [int(r[i] or 0) for i in ridx]) for r in src[2:] if len(r) > max(ridx)]
byfn, cur, line = collections.defaultdict(list), None, None
for l in dis.splitlines():
    mm = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if mm: line = (mm.group(1).split("/")[-1], int(mm.group(2))); continue
    mm = re.match(r"\s*\.text\.(\S+):", l)
    if mm: cur = mm.group(1); continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+.*?;", l): byfn[cur].append(line)
fn = [f for f, v in byfn.items() if len(v) == len(ins)]
with open(out, "w") as fo:
    fo.write("# %s\n\nkernel `%s`\n\n| counter | value |\n|---|---|\n" % (title, kname))
    for k in want:
        if k in m: fo.write("| %s | %s %s |\n" % (k, m[k][0], m[k][1]))
    tot = [sum(x[1][i] for x in ins) for i in range(len(reasons))]
    T = float(sum(tot)) or 1.0
    fo.write("\nWarp-stall samples by reason: " + ", ".join("%s %.1f%%" % (r[6:], 100 * v / T) for r, v in zip(reasons, tot)) + "\n")
    if fn:
        per = collections.Counter()
        for (n, rs), ln in zip(ins, byfn[fn[0]]): per[ln] += n
        S = float(sum(per.values())) or 1.0
        fo.write("\nSource lines with the most stall samples (file, line -> share of all samples):\n\n")
        for ln, n in per.most_common(25):
            text = ""
            if ln:
                try:
                    p = os.path.join(ROOT, "gpcounts_b200", "csrc", ln[0])
                    text = open(p).read().split("\n")[ln[1] - 1].strip()[:90] if os.path.exists(p) else ""
                except Exception:
                    pass
            fo.write("* %5.2f%%  %s:%s  `%s`\n" % (100 * n / S, ln[0] if ln else "?", ln[1] if ln else "?", text))
    else:
        fo.write("\n(source mapping unavailable: the library was rebuilt after this capture)\n")
print("wrote", out)
