import csv,re,subprocess,bisect
from collections import defaultdict
import sys, os, tempfile
# usage: ncu_by_phase.py <source-page csv> <libmocca_b200.so | cubin> [mb_core.cuh of that build] [kernel symbol]
csv_path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/src_r1s.csv"
cubin = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/sass/mb200.sm_100a.cubin"
core_src = sys.argv[3] if len(sys.argv) > 3 else "mocca_envs_b200/csrc/mb_core.cuh"
kname = sys.argv[4] if len(sys.argv) > 4 else "_Z22k_step_walker3d_custom8StepArgs"
if cubin.endswith(".so") or cubin.endswith(".o"):
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(cubin)], cwd=d, check=True, capture_output=True)
    # one cubin per env kind since round 2: take the one that defines the kernel
    _kn = (sys.argv[4] if len(sys.argv) > 4 else "_Z22k_step_walker3d_custom8StepArgs") if "by_phase" in __file__ else (sys.argv[3] if len(sys.argv) > 3 else "_Z22k_step_walker3d_custom8StepArgs")
    cubin = next(os.path.join(d, f) for f in sorted(os.listdir(d)) if f.endswith(".cubin") and
                 (".text." + _kn) in subprocess.run(["cuobjdump", "-elf", os.path.join(d, f)], capture_output=True, text=True).stdout)
dis = subprocess.run(["nvdisasm","-g","-c",cubin],capture_output=True,text=True).stdout.splitlines()
start = next(i for i,l in enumerate(dis) if l.startswith(".text."+kname+":"))
lines=[]; cur=("?",0)
for l in dis[start+1:]:
    if l.startswith("//-----") or l.startswith("\t.section"): break
    m=re.search(r'//## File "([^"]+)", line (\d+)',l)
    if m: cur=(m.group(1).split("/")[-1],int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/",l): lines.append(cur)
rows=list(csv.reader(open(csv_path))); hdr=rows[1]; ci={h:i for i,h in enumerate(hdr)}; body=rows[2:]
src=open(core_src).read().splitlines()
funcs=[(i+1,l.strip()[:56]) for i,l in enumerate(src) if "MB_HD static" in l or "MB_NOINLINE static" in l or l.startswith("MB_HD") or l.startswith("template <class M> MB_HD") or l.strip().startswith("template <bool BOXES>")]
starts=[f[0] for f in funcs]
agg=defaultdict(lambda:[0,0,0,0]); ti=ts=0
for k in range(min(len(body),len(lines))):
    f,ln=lines[k]; r=body[k]
    inst=int(r[ci["Instructions Executed"]] or 0); samp=int(r[ci["# Samples"]] or 0); thr=int(r[ci["Thread Instructions Executed"]] or 0)
    key=f
    if f=="mb_core.cuh": key="core:"+funcs[bisect.bisect_right(starts,ln)-1][1]
    agg[key][0]+=inst; agg[key][1]+=samp; agg[key][2]+=thr; agg[key][3]+=1; ti+=inst; ts+=samp
print("%s, 16384 envs, ncu --set full (" % kname + os.path.basename(csv_path) + "): %d SASS instructions, %d warp-instructions executed = %.0f per env-substep"%(len(body),ti,ti/65536))
print("%-60s %8s %14s %9s %7s %7s"%("source function","static","warp-inst/substep","inst%","stall%","lanes"))
for k,a in sorted(agg.items(), key=lambda kv:-kv[1][0])[:24]:
    print("%-60s %8d %14.0f %8.2f%% %6.2f%% %7.1f" % (k,a[3],a[0]/65536,100*a[0]/ti,100*a[1]/ts,a[2]/max(a[0],1)))
