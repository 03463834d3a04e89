"""Aggregate the warp-stall samples of an ncu source-page CSV (`ncu -i rep --page source --csv`) by code segment, where a
segment is the SASS between two CTA barriers (BAR.SYNC) - i.e. by phase of the evaluator.

    python tools/ncu_segments.py gpurun_out/r2_eval_source.csv [min_share_percent]
"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
minp = float(sys.argv[2]) if len(sys.argv) > 2 else 0.4
hdr = rows[1]
ia, isrc, ismp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
ridx = [hdr.index(r) for r in reasons]
ins = []
for r in rows[2:]:
    if len(r) <= max(ridx):
        continue
    ins.append((r[ia], r[isrc].strip(), int(r[ismp] or 0), int(r[iex] or 0), [int(r[i] or 0) for i in ridx]))
tot = sum(x[2] for x in ins)
print("instructions", len(ins), "samples", tot, "kernel", rows[0][1])
seg, cur = [], []
for x in ins:
    cur.append(x)
    if "BAR.SYNC" in x[1]:
        seg.append(cur); cur = []
seg.append(cur)
for k, sg in enumerate(seg):
    s = sum(x[2] for x in sg)
    if s < minp / 100 * tot:
        continue
    ex = sum(x[3] for x in sg)
    rs = [sum(x[4][i] for x in sg) for i in range(len(reasons))]
    top = sorted(zip(rs, reasons), reverse=True)[:4]
    exd = sum(x[3] for x in sg if "DMMA" in x[1])
    exf = sum(x[3] for x in sg if any(t in x[1] for t in ("DFMA", "DMUL", "DADD", "DSETP")))
    print("seg %3d @%s n_ins %5d samples %5.1f%% exec %9d (DMMA %7d FP64 %8d) | %s" % (
        k, sg[0][0], len(sg), 100 * s / tot, ex, exd, exf, ", ".join("%s %.0f%%" % (r[6:], 100 * v / max(s, 1)) for v, r in top)))
