"""Attribute ncu warp-stall samples (SASS page) to CUDA source lines via nvdisasm line info.

usage: python tools/ncu_lines.py <sass_page.csv from `ncu -i rep --page source --csv`> <nvdisasm -g -c listing> [top]
"""
import csv, re, sys, collections

sass_csv, dis, top = sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40
rows = list(csv.reader(open(sass_csv)))
hdr = rows[1]
ia, isrc, ismp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
ins = [(int(r[ia], 16), r[isrc].strip(), int(r[ismp] or 0), int(r[iex] or 0)) for r in rows[2:] if len(r) > iex]
# split into contiguous runs (functions)
runs, cur = [], [ins[0]]
for a in ins[1:]:
    if a[0] - cur[-1][0] != 16:
        runs.append(cur); cur = []
    cur.append(a)
runs.append(cur)
# parse nvdisasm listing
funcs, name, line, items = {}, None, None, None
for l in open(dis):
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m:
        name = m.group(1); items = []; funcs[name] = items; line = None; continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        if "inlined at" in l and False:
            pass
        line = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m and items is not None:
        items.append((int(m.group(1), 16), m.group(2).strip(), line))
norm = lambda s: re.sub(r"\s+", " ", s.replace(";", "")).strip()
by_line = collections.Counter(); by_func = collections.Counter(); total = 0
for run in runs:
    key = [norm(x[1]) for x in run[:6]]
    best = None
    for fn, it in funcs.items():
        if len(it) == len(run) and [norm(x[1]) for x in it[:6]] == key:
            best = fn; break
    if best is None:
        for fn, it in funcs.items():
            if len(it) == len(run):
                best = fn; break
    smp = sum(x[2] for x in run); total += smp
    by_func[best or "?(%d instrs)" % len(run)] += smp
    if best:
        for (a, s, n, e), (off, txt, ln) in zip(run, funcs[best]):
            by_line[(best[:28], ln)] += n
print("total samples", total)
for f, n in by_func.most_common(12):
    print("%6.2f%%  %s" % (100.0 * n / total, f))
print()
for (f, ln), n in by_line.most_common(top):
    print("%6.2f%%  %-28s %s" % (100.0 * n / total, f, ln))

if len(sys.argv) > 4:
    print("\nby file/line order (>= 0.15%):")
    for (f, ln), n in sorted(by_line.items(), key=lambda kv: (str(kv[0][1]))):
        if ln and 100.0 * n / total >= 0.15:
            print("%6.2f%%  %s" % (100.0 * n / total, ln))
