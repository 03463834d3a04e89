"""Inventory of the innermost SASS loops (label .. backward branch to it) of a function: instruction mix per loop.

usage: python tools/sass_loops.py <nvdisasm -g -c listing> [function substring] [min DMMA]
Shows, per loop, the source lines it comes from and counts of DMMA / LDS / LDL / STL / SEL / WARPSYNC / total.
"""
import re, sys, collections
dis = sys.argv[1]; fsel = sys.argv[2] if len(sys.argv) > 2 else ""; mind = int(sys.argv[3]) if len(sys.argv) > 3 else 1
name, line = None, None
ins = []   # (label or None, text, line)
for l in open(dis):
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m: name = m.group(1); continue
    if fsel not in (name or ""): continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m: line = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s*(\.L_x_\d+):", l)
    if m: ins.append((m.group(1), None, None)); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m: ins.append((None, m.group(2).strip(), line, m.group(1)))
pos = {x[0]: i for i, x in enumerate(ins) if x[0]}
loops = []
for i, x in enumerate(ins):
    if x[0] is None and "BRA" in x[1]:
        m = re.search(r"`\((\.L_x_\d+)\)", x[1])
        if m and m.group(1) in pos and pos[m.group(1)] < i:
            loops.append((pos[m.group(1)], i))
# innermost only
inner = [lp for lp in loops if not any(o != lp and lp[0] <= o[0] and o[1] <= lp[1] for o in loops)]
for a, b in inner:
    body = [x for x in ins[a:b + 1] if x[0] is None]
    c = collections.Counter()
    for x in body:
        op = x[1].split()[1] if x[1].startswith("@") else x[1].split()[0]
        c[op.split(".")[0]] += 1
    if c["DMMA"] < mind: continue
    lines = collections.Counter(x[2] for x in body if x[2] and x[2][0] == "svgp_fused.cuh")
    top = sorted(lines)[:1] + sorted(lines)[-1:]
    print("loop @%s..%s n=%4d DMMA %3d LDS %3d LDL %2d STL %2d SEL %2d WARPSYNC %2d DMUL %2d | lines %s" % (
        body[0][3], body[-1][3], len(body), c["DMMA"], c["LDS"], c["LDL"], c["STL"], c["SEL"] + c["FSEL"], c["WARPSYNC"], c["DMUL"],
        ",".join(str(l[1]) for l in top)))
