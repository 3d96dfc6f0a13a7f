#!/usr/bin/env python
"""Do two kernels of the library round alike?  Compares, per source line, the multiset of floating-point opcodes (with
their negation pattern) of two kernels, e.g. the device-buffer and host-buffer instantiations of a step kernel.
usage: sass_fp_diff.py <lib.so> <kernel symbol A> <kernel symbol B>"""
import os, re, subprocess, sys, tempfile
from collections import Counter
lib, ka, kb = sys.argv[1:4]
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, check=True, capture_output=True)
def fp_ops(kname):
    cubin = next(os.path.join(d, f) for f in sorted(os.listdir(d)) if f.endswith(".cubin") and
                 (".text." + kname) in subprocess.run(["cuobjdump", "-elf", os.path.join(d, f)], capture_output=True, text=True).stdout)
    dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
    start = next(i for i, l in enumerate(dis) if l.startswith(".text." + kname + ":"))
    cur = ("?", 0); c = Counter()
    for l in dis[start + 1:]:
        if l.startswith("//-----") or l.startswith("\t.section"): break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?(FFMA|FMUL|FADD|MUFU)\S*\s+(.*);", l)
        if m:
            ops = m.group(3).split(",")
            sig = m.group(2) + "".join("-" if o.strip().startswith("-") else "+" for o in ops[1:])
            c[(cur, sig)] += 1
    return c
A, B = fp_ops(ka), fp_ops(kb)
bad = 0
for k in sorted(set(A) | set(B)):
    if A[k] != B[k]:
        bad += 1
        print("%s:%d %s  %d vs %d" % (k[0][0], k[0][1], k[1], A[k], B[k]))
print("differing (line, opcode) pairs:", bad)
