#!/usr/bin/env python
"""Join an ncu SASS source-page CSV with nvdisasm -g line info and aggregate per CUDA source line.

usage: ncu_by_line.py <src_sass.csv (ncu --page source --csv)> <cubin> <kernel mangled name> [top N]
"""
import csv
import re
import subprocess
import sys
from collections import defaultdict


def main():
    csv_path, cubin, kname = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
    # locate function
    start = next(i for i, l in enumerate(dis) if l.startswith(".text." + kname + ":"))
    lines = []  # (file, line) per instruction in order
    cur = ("?", 0)
    inl = None
    for l in dis[start + 1:]:
        if l.startswith("//-----") or l.startswith("\t.section"):
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
            lines.append(cur)
    rows = list(csv.reader(open(csv_path)))
    hdr = rows[1]
    ci = {h: i for i, h in enumerate(hdr)}
    body = rows[2:]
    n = min(len(body), len(lines))
    agg = defaultdict(lambda: [0, 0, 0, 0])
    tot_inst = tot_samp = 0
    for k in range(n):
        r = body[k]
        inst = int(r[ci["Instructions Executed"]] or 0)
        samp = int(r[ci["# Samples"]] or 0)
        thr = int(r[ci["Thread Instructions Executed"]] or 0)
        exc = int(r[ci["L1 Wavefronts Shared Excessive"]] or 0)
        a = agg[lines[k]]
        a[0] += inst; a[1] += samp; a[2] += thr; a[3] += exc
        tot_inst += inst; tot_samp += samp
    print("instructions(csv)=%d sass(nvdisasm)=%d total inst=%d samples=%d" % (len(body), len(lines), tot_inst, tot_samp))
    print("%-22s %12s %7s %8s %7s %6s %10s" % ("file:line", "warp-inst", "inst%", "samples", "samp%", "lanes", "bank-exc"))
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print("%-22s %12d %6.2f%% %8d %6.2f%% %6.1f %10d" % ("%s:%d" % key, a[0], 100.0 * a[0] / max(tot_inst, 1), a[1],
                                                   100.0 * a[1] / max(tot_samp, 1), a[2] / max(a[0], 1), a[3]))


if __name__ == "__main__":
    main()
