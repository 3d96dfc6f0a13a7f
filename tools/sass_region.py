#!/usr/bin/env python
"""Sequential SASS of a kernel between the first and last instruction attributed to a source-line range (all files shown).
usage: sass_region.py <lib.so> <kernel symbol> <file> <first line> <last line>"""
import os, re, subprocess, sys, tempfile
lib, kname, fname, lo, hi = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5])
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, check=True, capture_output=True)
cubin = next(os.path.join(d, f) for f in sorted(os.listdir(d)) if f.endswith(".cubin") and
             (".text." + kname) in subprocess.run(["cuobjdump", "-elf", os.path.join(d, f)], capture_output=True, text=True).stdout)
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.startswith(".text." + kname + ":"))
cur = ("?", 0); rows = []
for l in dis[start + 1:]:
    if l.startswith("//-----") or l.startswith("\t.section"): break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l): rows.append((cur, re.sub(r"/\*[0-9a-f]+\*/", "", l).strip()))
    elif re.match(r"^\.L_x_\d+:", l): rows.append((("label", 0), l.strip()))
idx = [i for i, (c, _) in enumerate(rows) if c[0] == fname and lo <= c[1] <= hi]
for c, t in rows[idx[0]:idx[-1] + 1]: print("%-22s %s" % ("%s:%d" % c, t))
