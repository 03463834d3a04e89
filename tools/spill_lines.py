"""Local-memory traffic (STL / LDL = register spills and ABI saves) of a kernel by CUDA source line.

usage: python tools/spill_lines.py <object or cubin> [function substring] [file suffix]
"""
import collections, os, re, subprocess, sys, tempfile
obj = os.path.abspath(sys.argv[1]); fsel = sys.argv[2] if len(sys.argv) > 2 else ""; suf = sys.argv[3] if len(sys.argv) > 3 else ""
tmp = tempfile.mkdtemp()
if not obj.endswith(".cubin"):
    subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=tmp, capture_output=True)
    obj = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", obj], capture_output=True, text=True).stdout
name, line, cnt = None, None, collections.Counter()
for l in dis.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m: name = m.group(1); continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m: line = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.search(r"\b(STL|LDL)\b", l)
    if m and fsel in (name or "") and line and line[0].endswith(suf):
        cnt[(line, m.group(1))] += 1
for (ln, k), v in sorted(cnt.items(), key=lambda kv: (kv[0][0][0], kv[0][0][1])):
    print("%-22s %5d  %s %d" % (ln[0], ln[1], k, v))
