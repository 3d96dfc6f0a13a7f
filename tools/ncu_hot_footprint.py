#!/usr/bin/env python
"""Static code footprint of the hot path: SASS instructions executed by (almost) every warp every substep.
usage: ncu_hot_footprint.py <ncu source csv> <cubin> <kernel> <hot threshold> <warm threshold>"""
import bisect
import csv
import os
import re
import subprocess
import sys
from collections import Counter

csv_path, cubin, kname = sys.argv[1:4]
hot_thr, warm_thr = int(sys.argv[4]), int(sys.argv[5])
rows = list(csv.reader(open(csv_path)))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
body = rows[2:]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.startswith(".text." + kname + ":"))
lines = []
cur = ("?", 0)
for l in dis[start + 1:]:
    if l.startswith("//-----") or l.startswith("\t.section"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur)
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "mocca_envs_b200", "csrc")
src = open(os.path.join(root, "mb_core.cuh")).read().splitlines()
funcs = [(i + 1, l.strip()[:50]) for i, l in enumerate(src)
         if "MB_HD static" in l or "MB_NOINLINE static" in l or l.startswith("MB_HD") or l.startswith("template <class M> MB_HD")]
starts = [f[0] for f in funcs]
hot, warm = Counter(), Counter()
nh = nw = 0
for k, r in enumerate(body):
    ex = int(r[ci["Instructions Executed"]] or 0)
    f, ln = lines[k]
    key = f
    if f == "mb_core.cuh":
        key = "core:" + funcs[bisect.bisect_right(starts, ln) - 1][1]
    if ex >= hot_thr:
        hot[key] += 1
        nh += 1
    elif ex >= warm_thr:
        warm[key] += 1
        nw += 1
print("hot static instrs (>= %d warp executions): %d = %.1f KB; warm (>= %d): %d = %.1f KB"
      % (hot_thr, nh, nh * 16 / 1024, warm_thr, nw, nw * 16 / 1024))
for k, v in hot.most_common(25):
    print("  %5d %s" % (v, k))
print("warm:")
for k, v in warm.most_common(8):
    print("  %5d %s" % (v, k))
