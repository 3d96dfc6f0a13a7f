#!/usr/bin/env python
"""Where the host-buffer step (mb200_step_host) spends its time: H2D of the actions, step kernel (+ scheduler sort),
D2H of obs / reward / done / trunc, and the host-side remainder (call overhead, stream-synchronise wake-up)."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
from mocca_envs_b200.vec_env import Walker3DCustomVecEnv

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
dev = torch.device("cuda:0")
env = Walker3DCustomVecEnv(N, device=dev, seed=1)
env.reset()
A, OB = env.act_dim, env.obs_dim
h_act = (torch.rand(N, A) * 2 - 1).pin_memory()
h_obs = torch.empty(N, OB).pin_memory(); h_rew = torch.empty(N).pin_memory()
h_done = torch.empty(N, dtype=torch.uint8).pin_memory(); h_trunc = torch.empty(N, dtype=torch.uint8).pin_memory()
d_act = torch.empty(N, A, device=dev)
outs = (h_obs.numpy(), h_rew.numpy(), h_done.numpy(), h_trunc.numpy())
for _ in range(50):
    env.step_host(h_act.numpy(), outs)
K = 200
ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(K)]
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(K):
    ev[i][0].record()
    d_act.copy_(h_act, non_blocking=True)
    ev[i][1].record()
    obs, rew, done, info = env.step(d_act)
    ev[i][2].record()
    h_obs.copy_(obs, non_blocking=True); h_rew.copy_(rew, non_blocking=True)
    h_done.copy_(done, non_blocking=True); h_trunc.copy_(info["TimeLimit.truncated"], non_blocking=True)
    ev[i][3].record()
    ev[i][3].synchronize()
wall_py = (time.perf_counter() - t0) / K
h2d = np.mean([e[0].elapsed_time(e[1]) for e in ev]); ker = np.mean([e[1].elapsed_time(e[2]) for e in ev])
d2h = np.mean([e[2].elapsed_time(e[3]) for e in ev])
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(K):
    env.step_host(h_act.numpy(), outs)
wall_c = (time.perf_counter() - t0) / K
print("N=%d  pieces (python-driven): H2D %.1f us, step kernel + sort %.1f us, D2H %.1f us, sum %.1f us, wall %.1f us"
      % (N, h2d * 1e3, ker * 1e3, d2h * 1e3, (h2d + ker + d2h) * 1e3, wall_py * 1e6))
print("mb200_step_host wall: %.1f us per step = %.2f M env-steps/s; bytes H2D %d, D2H %d -> %.1f / %.1f GB/s"
      % (wall_c * 1e6, N / wall_c / 1e6, N * A * 4, N * (OB * 4 + 6), N * A * 4 / (h2d * 1e-3) / 1e9,
         N * (OB * 4 + 6) / (d2h * 1e-3) / 1e9))
