#!/usr/bin/env python
"""Distil ncu output under gpurun_out/ into the tracked summaries under profiles/.
usage: make_profiles.py <tag> <prof.ncu-rep> [launches.csv]
  profiles/<tag>_step_kernel_raw.csv   subset of `ncu --page raw` (time, DRAM, issue, stalls, caches, occupancy)
  profiles/<tag>_by_phase.txt          SASS page joined with nvdisasm line info, per source function
  profiles/<tag>_launches.csv          (if given) the gpu__time_duration launch list, as captured"""
import csv, os, shutil, subprocess, sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
tag, rep = sys.argv[1], sys.argv[2]
out = os.path.join(ROOT, "profiles")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2:]
KEEP = ("Kernel Name", "gpu__time_duration", "dram__bytes", "dram__throughput", "gpu__dram_throughput", "smsp__issue_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed", "issue_stalled", "sm__warps_active",
        "launch__", "l1tex__t_sector_hit_rate", "lts__t_sector_hit_rate", "sm__icc_request_hit_rate",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__pipe_fma_cycles_active", "sm__pipe_alu_cycles_active", "sm__inst_executed_pipe_lsu",
        "l1tex__t_sectors_pipe_lsu_mem_local", "gcc__cache_requests_type_instruction.sum", "sm__throughput",
        "smsp__cycles_active.avg")
with open(os.path.join(out, tag + "_step_kernel_raw.csv"), "w") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit"] + ["launch %d" % (i + 1) for i in range(len(vals))])
    for k, name in enumerate(hdr):
        if any(s in name for s in KEEP) and "pcsamp" not in name:
            w.writerow([name, units[k]] + [v[k] for v in vals])
src_csv = os.path.join(ROOT, "gpurun_out", "src_" + tag + ".csv")
with open(src_csv, "w") as f:  # the first captured launch only (the page repeats per launch)
    page = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    lines = page.splitlines(True)
    starts = [i for i, l in enumerate(lines) if l.startswith('"Kernel Name"')]
    f.write("".join(lines[starts[0]:starts[1]] if len(starts) > 1 else lines))
lib = os.path.join(ROOT, "mocca_envs_b200", "libmocca_b200.so")
with open(os.path.join(out, tag + "_by_phase.txt"), "w") as f:
    f.write(subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_by_phase.py"), src_csv, lib],
                           capture_output=True, text=True, cwd=ROOT).stdout)
    f.write("\n" + subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_cold_islands.py"), src_csv, lib],
                                  capture_output=True, text=True, cwd=ROOT).stdout.splitlines()[0] + "\n")
if len(sys.argv) > 3:
    shutil.copy(sys.argv[3], os.path.join(out, tag + "_launches.csv"))
print("wrote profiles/%s_*" % tag)
