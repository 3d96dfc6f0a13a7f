#!/usr/bin/env python
"""Static SASS view of a source-line range: usage sass_lines.py <lib.so> <kernel symbol> <file> <first line> <last line> [-v]
Prints the number of SASS instructions attributed (nvdisasm -g line info) to each line of the range, -v lists them."""
import os, re, subprocess, sys, tempfile
lib, kname, fname, lo, hi = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5])
verbose = "-v" in sys.argv
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, check=True, capture_output=True)
cubin = next(os.path.join(d, f) for f in sorted(os.listdir(d)) if f.endswith(".cubin") and
             (".text." + kname) in subprocess.run(["cuobjdump", "-elf", os.path.join(d, f)], capture_output=True, text=True).stdout)
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.startswith(".text." + kname + ":"))
cur = ("?", 0); counts = {}; total = 0
for l in dis[start + 1:]:
    if l.startswith("//-----") or l.startswith("\t.section"): break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        total += 1
        if cur[0] == fname and lo <= cur[1] <= hi:
            counts[cur[1]] = counts.get(cur[1], 0) + 1
            if verbose: print(cur[1], re.sub(r"/\*[0-9a-f]+\*/", "", l).strip())
print("kernel total", total)
for k in sorted(counts): print(k, counts[k])
print("range total", sum(counts.values()))
