"""Warp-stall samples of an ncu source-page CSV in chunks of consecutive SASS instructions inside one barrier-delimited
segment (see tools/ncu_segments.py for the segment numbers): where inside a long phase the time goes.

usage: python tools/ncu_chunks.py <source.csv> <segment> [chunk]
"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1]))); segno = int(sys.argv[2]); chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 100
hdr = rows[1]
ia, isrc, ismp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
ridx = [hdr.index(r) for r in reasons]
ins = [(r[ia], r[isrc].strip(), int(r[ismp] or 0), int(r[iex] or 0), [int(r[i] or 0) for i in ridx]) for r in rows[2:] if len(r) > max(ridx)]
tot = sum(x[2] for x in ins)
seg, cur = [], []
for x in ins:
    cur.append(x)
    if "BAR.SYNC" in x[1]: seg.append(cur); cur = []
seg.append(cur)
sg = seg[segno]
for i in range(0, len(sg), chunk):
    c = sg[i:i + chunk]
    s = sum(x[2] for x in c)
    rs = sorted(((sum(x[4][k] for x in c), reasons[k][6:]) for k in range(len(reasons))), reverse=True)[:3]
    nd = sum(1 for x in c if "DMMA" in x[1]); nf = sum(1 for x in c if any(t in x[1] for t in ("DFMA", "DMUL", "DADD")))
    nl = sum(1 for x in c if "LDS" in x[1]); ex = max(x[3] for x in c)
    print("ins %5d..%5d samples %5.2f%% exec/ins %7d DMMA %3d FP64 %3d LDS %3d | %s" % (
        i, i + len(c), 100.0 * s / tot, ex, nd, nf, nl, ", ".join("%s %.0f%%" % (r, 100.0 * v / max(s, 1)) for v, r in rs)))
