#!/usr/bin/env python
"""Rarely executed code that sits inline inside the hot loop ("cold islands"): every taken branch over such a block
breaks the sequential instruction prefetch, and the hot loop of the step kernel is larger than the 32 KB L1.5
instruction cache.  usage: ncu_cold_islands.py <ncu source-page csv> <libmocca_b200.so | cubin> [kernel symbol]"""
import csv, os, re, subprocess, sys, tempfile

csv_path, cubin = sys.argv[1:3]
kname = sys.argv[3] if len(sys.argv) > 3 else "_Z22k_step_walker3d_custom8StepArgs"
if cubin.endswith(".so"):
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(cubin)], cwd=d, check=True, capture_output=True)
    # one cubin per env kind since round 2: take the one that defines the kernel
    _kn = (sys.argv[4] if len(sys.argv) > 4 else "_Z22k_step_walker3d_custom8StepArgs") if "by_phase" in __file__ else (sys.argv[3] if len(sys.argv) > 3 else "_Z22k_step_walker3d_custom8StepArgs")
    cubin = next(os.path.join(d, f) for f in sorted(os.listdir(d)) if f.endswith(".cubin") and
                 (".text." + _kn) in subprocess.run(["cuobjdump", "-elf", os.path.join(d, f)], capture_output=True, text=True).stdout)
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.startswith(".text." + kname + ":"))
lines, cur = [], ("?", 0)
for l in dis[start + 1:]:
    if l.startswith("//-----") or l.startswith("\t.section"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur)
rows = list(csv.reader(open(csv_path)))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
ex = [int(r[ci["Instructions Executed"]] or 0) for r in rows[2:]]
n = min(len(ex), len(lines))
HOT, COLD = 30000, 1000  # warp executions per launch of 16384 envs x 4 substeps
hot_idx = [k for k in range(n) if ex[k] >= HOT]
lo, hi = hot_idx[0], hot_idx[-1]
isl, k = [], lo
while k <= hi:
    if ex[k] <= COLD:
        j = k
        while j <= hi and ex[j] <= COLD:
            j += 1
        isl.append((j - k, k, j))
        k = j
    else:
        k += 1
tot = sum(a for a, _, _ in isl)
print("hot region: SASS %d..%d (%d instrs = %.1f KB); cold islands inside it: %d instrs = %.1f KB in %d blocks"
      % (lo, hi, hi - lo + 1, (hi - lo + 1) * 16 / 1024, tot, tot * 16 / 1024, len(isl)))
for size, a, b in sorted(isl, reverse=True)[:30]:
    from collections import Counter
    c = Counter(lines[x] for x in range(a, b))
    (f, ln), _ = c.most_common(1)[0]
    print("  %5d instrs @%6d  max exec %6d  mostly %s:%d" % (size, a, max(ex[a:b]), f, ln))
