#!/usr/bin/env python
"""Do two kernels of the library round alike?  Compares, per source line, the multiset of floating-point opcodes (with
their negation pattern) of two kernels, e.g. the device-buffer and host-buffer instantiations of a step kernel: a sum of
two products can be contracted into an FMA either way round, and the compiler is free to decide differently in two
instantiations of the same source (it did, once: the Euler angles of round 2).
usage: sass_fp_diff.py <lib.so> <kernel symbol A> <kernel symbol B>"""
import os
import re
import subprocess
import sys
import tempfile
from collections import Counter


def extract(lib):
    """{kernel symbol: cubin path} of every kernel in the library's embedded cubins."""
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, check=True, capture_output=True)
    out = {}
    for f in sorted(os.listdir(d)):
        if f.endswith(".cubin"):
            elf = subprocess.run(["cuobjdump", "-elf", os.path.join(d, f)], capture_output=True, text=True).stdout
            for m in re.finditer(r"\.text\.(_Z\w+)", elf):
                out[m.group(1)] = os.path.join(d, f)
    return out


def fp_ops(cubin, kname):
    dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
    start = next(i for i, l in enumerate(dis) if l.startswith(".text." + kname + ":"))
    cur, c = ("?", 0), Counter()
    for l in dis[start + 1:]:
        if l.startswith("//-----") or l.startswith("\t.section"):
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?(FFMA|FMUL|FADD|MUFU)\S*\s+(.*);", l)
        if m:
            ops = m.group(3).split(",")
            sig = m.group(2)
            if sig != "FADD":  # (a - b as FADD a, -b or FADD -b, a is the same rounding)
                sig += "".join("-" if o.strip().startswith("-") else "+" for o in ops[1:])
            c[(cur, sig)] += 1
    return c


def differences(cubins, ka, kb):
    A, B = fp_ops(cubins[ka], ka), fp_ops(cubins[kb], kb)
    return [(k[0][0], k[0][1], k[1], A[k], B[k]) for k in sorted(set(A) | set(B)) if A[k] != B[k]]


if __name__ == "__main__":
    lib, ka, kb = sys.argv[1:4]
    bad = differences(extract(lib), ka, kb)
    for f, ln, sig, a, b in bad:
        print("%s:%d %s  %d vs %d" % (f, ln, sig, a, b))
    print("differing (line, opcode) pairs:", len(bad))
