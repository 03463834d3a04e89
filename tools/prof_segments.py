"""Instruction mix of the profiling build between consecutive phase-timer reads (SR_CLOCK) of a function.

usage: python tools/prof_segments.py <nvdisasm -g -c listing of the -DGPC_PHASE_PROF object> [function substring]
"""
import re, sys, collections
dis = sys.argv[1]; fsel = sys.argv[2] if len(sys.argv) > 2 else "k_fit"
name, line, seg, cur = None, None, [], []
for l in open(dis):
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m: name = m.group(1); continue
    if fsel not in (name or ""): continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m: line = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if not m: continue
    cur.append((m.group(1), m.group(2), line))
    if "SR_CLOCK" in m.group(2):
        seg.append(cur); cur = []
for sg in seg:
    lines = [x[2][1] for x in sg if x[2] and x[2][0] == "svgp_fused.cuh"]
    if not lines: continue
    c = collections.Counter()
    for x in sg:
        t = x[1]; c[(t.split()[1] if t.startswith("@") else t.split()[0]).split(".")[0]] += 1
    fp = c["DFMA"] + c["DMUL"] + c["DADD"] + c["DSETP"]
    print("end@%s n=%5d lines %4d..%4d mark@%s DMMA %3d FP64 %4d LDS %3d STS %3d LDL %2d STL %2d SHFL %2d BAR %d LD %d MUFU %d" % (
        sg[-1][0], len(sg), min(lines), max(lines), sg[-1][2][1] if sg[-1][2] else "?", c["DMMA"], fp, c["LDS"], c["STS"], c["LDL"], c["STL"], c["SHFL"],
        c["BAR"], c["LDG"] + c["LD"], c["MUFU"]))
