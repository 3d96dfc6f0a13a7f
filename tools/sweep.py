#!/usr/bin/env python
"""BASELINE.json configs[4]: throughput sweep 1K-65K envs/GPU for each env alone, plus a mixed run (N/4 envs of
each of the four env types resident on one GPU, one step kernel per type on its own CUDA stream).

  python tools/sweep.py [--steps K] [--out gpurun_out/sweep.json]

Device-timed with CUDA events (per configuration: 20 warm-up steps, K timed steps, random actions from a device pool).
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from mocca_envs_b200.vec_env import (CassieVecEnv, Child3DCustomVecEnv, Crab2DCustomVecEnv,  # noqa: E402
                                     MikeStepperVecEnv, Monkey3DCustomVecEnv, Walker2DCustomVecEnv,
                                     Walker3DCustomVecEnv, Walker3DStepperVecEnv)

KINDS = {"custom": (Walker3DCustomVecEnv, 1.0), "stepper": (Walker3DStepperVecEnv, 1.0),
         "monkey": (Monkey3DCustomVecEnv, 1.0), "cassie": (CassieVecEnv, 0.1)}
# SURVEY 8 f3 envs: swept alone, not part of the mixed run (BASELINE config 5 names the four envs above)
EXTRA = {"child": (Child3DCustomVecEnv, 1.0), "mike": (MikeStepperVecEnv, 1.0),
         "walker2d": (Walker2DCustomVecEnv, 1.0), "crab2d": (Crab2DCustomVecEnv, 1.0)}


def make(kind, n, dev):
    cls, scale = {**KINDS, **EXTRA}[kind]
    env = cls(n, device=dev, seed=1234)
    if kind in ("stepper", "mike"):
        import numpy as np

        env.set_env_params({"curriculum": np.array([0, 5, 9] * (n // 3 + 1))[:n]})
    env.reset()
    g = torch.Generator(device=dev).manual_seed(1)
    pool = (torch.rand(16, n, env.act_dim, device=dev, generator=g) * 2 - 1) * scale
    return env, pool


def run_single(kind, n, steps, dev):
    env, pool = make(kind, n, dev)
    for i in range(20):
        env.step(pool[i % 16])
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        env.step(pool[i % 16])
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1)
    env.close()
    return n * steps / (ms * 1e-3), ms / steps


def run_mixed(n_total, steps, dev):
    n = n_total // 4
    envs = {k: make(k, n, dev) for k in KINDS}
    streams = {k: torch.cuda.Stream(device=dev) for k in KINDS}
    torch.cuda.synchronize(dev)

    def step_all(i):
        for k, (env, pool) in envs.items():
            with torch.cuda.stream(streams[k]):
                env.step(pool[i % 16])

    for i in range(10):
        step_all(i)
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step_all(i)
    for s in streams.values():
        torch.cuda.current_stream(dev).wait_stream(s)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1)
    for env, _ in envs.values():
        env.close()
    return 4 * n * steps / (ms * 1e-3), ms / steps, time.perf_counter() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep.json"))
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    out = {"steps": args.steps, "single": {}, "mixed": {}}
    for kind in list(KINDS) + list(EXTRA):
        out["single"][kind] = {}
        for n in (1024, 2048, 4096, 8192, 16384, 32768, 65536):
            k = args.steps if kind != "cassie" else max(20, args.steps // 10)
            v, ms = run_single(kind, n, k, dev)
            out["single"][kind][n] = {"env_steps_per_s": v, "ms_per_step": ms}
            print("%-8s N=%6d  %10.3f M env-steps/s  %8.3f ms/step" % (kind, n, v / 1e6, ms), flush=True)
    for n in (4096, 16384, 65536):
        v, ms, wall = run_mixed(n, max(20, args.steps // 10), dev)
        out["mixed"][n] = {"env_steps_per_s": v, "ms_per_step": ms}
        print("mixed    N=%6d  %10.3f M env-steps/s  %8.3f ms/step (4 env types x N/4, 4 streams)" % (n, v / 1e6, ms), flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
