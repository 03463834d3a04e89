"""Instruction pattern of the basic blocks that hold many DMMAs: M = DMMA, f = FP64 arithmetic, l = LDS, s = STS, . = other.

usage: python tools/sass_blocks.py <nvdisasm -g -c listing> [function substring] [min DMMA per block]
Shows how the scheduler interleaved tensor-pipe and scalar FP64 work inside a block.
"""
import re, sys
dis = sys.argv[1]; fsel = sys.argv[2] if len(sys.argv) > 2 else ""; mind = int(sys.argv[3]) if len(sys.argv) > 3 else 40
name, blocks, cur = None, [], []
for l in open(dis):
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m: name = m.group(1); continue
    if fsel not in (name or ""): continue
    if re.match(r"\s*\.L_x_\d+:", l):
        blocks.append(cur); cur = []; continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if not m: continue
    t = m.group(2); op = (t.split()[1] if t.startswith("@") else t.split()[0])
    cur.append((m.group(1), op))
    if op.startswith("BRA") or op.startswith("EXIT") or op.startswith("RET") or op.startswith("BAR"):
        blocks.append(cur); cur = []
blocks.append(cur)
for b in blocks:
    nd = sum(1 for x in b if x[1].startswith("DMMA"))
    if nd < mind: continue
    pat = "".join("M" if o.startswith("DMMA") else "f" if o.split(".")[0] in ("DFMA", "DMUL", "DADD", "DSETP") else
                  "l" if o.startswith("LDS") else "s" if o.startswith("STS") else "." for _, o in b)
    print("block @%s n=%d DMMA=%d FP64=%d" % (b[0][0], len(b), nd, pat.count("f")))
    for i in range(0, len(pat), 120): print("   " + pat[i:i + 120])
