"""Print the SASS (nvdisasm -g -c listing) attributed to a range of source lines of one file, in address order.

usage: python tools/sass_lines.py <listing> <file suffix> <first line> <last line> [function substring]
"""
import re, sys
dis, fsuf, l0, l1 = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
fsel = sys.argv[5] if len(sys.argv) > 5 else ""
name, line, on = None, None, False
for l in open(dis):
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m:
        name = m.group(1); on = fsel in name; line = None; continue
    if not on:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        line = (m.group(1), int(m.group(2))); continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/", l) or re.match(r"\s*\.L_", l):
        if line and line[0].endswith(fsuf) and l0 <= line[1] <= l1:
            print("%4d %s" % (line[1], l.rstrip()[:150]))
        elif re.match(r"\s*\.L_", l):
            print("     %s" % l.rstrip())
