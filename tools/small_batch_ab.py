import os, sys, subprocess
sys.path.insert(0, '/root/repo')
if len(sys.argv) > 1:
    import torch
    from tools.sweep import run_single
    dev = torch.device("cuda", 0)
    for kind in ("custom", "stepper"):
        for n in (1024, 2048, 3072, 4096):
            v, ms = run_single(kind, n, 300, dev)
            print("MB200_WARPS=%s %-8s N=%5d %8.3f M/s %7.3f ms" % (os.environ.get("MB200_WARPS", "auto"), kind, n, v / 1e6, ms), flush=True)
else:
    for w in ("14", "7", "4", None):
        env = dict(os.environ)
        if w: env["MB200_WARPS"] = w
        else: env.pop("MB200_WARPS", None)
        subprocess.check_call([sys.executable, __file__, "child"], env=env)
