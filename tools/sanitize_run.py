"""Workload for compute-sanitizer (tools/sanitize.sh): every step kernel and its _host instantiation, the reset, physics
and debug kernels of every env kind, 64 envs (pad envs in the last CTA included), a few steps through auto-resets."""
import sys

import numpy as np
import torch

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from mocca_envs_b200 import vec_env as V  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 6
KINDS = [(V.Walker3DCustomVecEnv, {}), (V.Walker3DStepperVecEnv, {}), (V.Walker3DStepperVecEnv, {"plank_class": "Pillar"}),
         (V.Monkey3DCustomVecEnv, {}), (V.CassieVecEnv, {}), (V.Child3DCustomVecEnv, {}), (V.Walker2DCustomVecEnv, {}),
         (V.Crab2DCustomVecEnv, {}), (V.MikeStepperVecEnv, {}), (V.MikeStepperVecEnv, {"plank_class": "Pillar"})]
only = sys.argv[3] if len(sys.argv) > 3 else None
for cls, kw in KINDS:
    if only and only not in cls.__name__:
        continue
    env = cls(N, device="cuda:0", seed=1, return_final_obs=True, **kw)
    env.reset()
    g = torch.Generator(device="cuda:0").manual_seed(0)
    rng = np.random.RandomState(0)
    for k in range(STEPS):
        env.step(torch.rand(N, env.act_dim, device="cuda:0", generator=g) * 2 - 1)
        env.step_host(rng.uniform(-1, 1, (N, env.act_dim)).astype(np.float32))
    h = torch.empty(N, env.act_dim).pin_memory()
    outs = (torch.empty(N, env.obs_dim).pin_memory().numpy(), torch.empty(N).pin_memory().numpy(),
            torch.empty(N, dtype=torch.uint8).pin_memory().numpy(), torch.empty(N, dtype=torch.uint8).pin_memory().numpy())
    for k in range(STEPS):
        h.uniform_(-1, 1)
        env.step_host(h.numpy(), outs)
    env.step_physics(torch.zeros(N, env.nu - 6, device="cuda:0"))
    env.mass_matrix()
    env.reset_host(mask=(np.arange(N) % 2).astype(np.uint8))
    env.step_info()
    torch.cuda.synchronize()
    st = env.stats()
    print(cls.__name__, kw, "ok", st["episodes"], st["nonfinite"], st["overflow"], flush=True)
    env.close()
