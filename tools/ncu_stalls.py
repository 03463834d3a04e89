"""Per-source-line stall-reason breakdown from an ncu SASS source page (see tools/ncu_lines.py for the mapping)."""
import csv, re, sys, collections
sass_csv, dis, top = sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 25
rows = list(csv.reader(open(sass_csv)))
hdr = rows[1]
ia, isrc, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples")
reasons = ["stall_wait", "stall_short_sb", "stall_barrier", "stall_math", "stall_long_sb", "stall_branch_resolving",
           "stall_selected", "stall_not_selected", "stall_no_inst", "stall_mio", "stall_dispatch"]
ridx = [hdr.index(r) for r in reasons]
ins = [(int(r[ia], 16), r[isrc].strip(), int(r[ismp] or 0), [int(r[i] or 0) for i in ridx]) for r in rows[2:] if len(r) > max(ridx)]
items, line = [], None
for l in open(dis):
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m: line = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m: cur = m.group(1); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m: items.append((cur, int(m.group(1), 16), line))
# the profiled kernel = the function whose instruction count matches
byfn = collections.defaultdict(list)
for fn, off, ln in items: byfn[fn].append(ln)
fn = [f for f, v in byfn.items() if len(v) == len(ins)][0]
tot = [0] * len(reasons); per = collections.defaultdict(lambda: [0] * len(reasons))
for (a, s, n, rs), ln in zip(ins, byfn[fn]):
    for i, v in enumerate(rs): tot[i] += v; per[ln][i] += v
T = sum(tot)
print("kernel", fn[:60]); print("total", {r[6:]: "%.1f%%" % (100.0 * v / T) for r, v in zip(reasons, tot)})
for i, r in enumerate(reasons[:5]):
    print("\n== top lines for", r, "(%.1f%% of all stall samples)" % (100.0 * tot[i] / T))
    for ln, v in sorted(per.items(), key=lambda kv: -kv[1][i])[:top]:
        if v[i] * 200 > tot[i]: print("   %5.1f%%  %s" % (100.0 * v[i] / tot[i], ln))
