#!/usr/bin/env python
"""State-trajectory export (the debuggability half of the reference's render path): run a batch on the
GPU and write what a viewer needs to replay it -- per frame and env the base pose and joint angles, plus actions,
rewards, done flags and, for the terrain envs, the stepping-stone / bar tables.

  python tools/export_trajectory.py --env Walker3DStepperEnv-v0 --envs 4 --steps 300 --policy zero --out traj.npz
  python tools/replay_pybullet.py traj.npz            # on a host with pybullet + the reference installed

The file is a plain .npz: states [T + 1, N, 13 + 2A] = pos3 quat4(xyzw) omega3 vel3 q[A] qd[A] (the layout of
mb200_get_state), actions [T, N, A], rewards [T, N], dones [T, N], env_id, dt, and terrain [N, 20, 6] / bars
[N, 32, 4] when the env has them.  After an auto-reset the next frame is the first state of the new episode."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--env", default="Walker3DCustomEnv-v0")
    ap.add_argument("--envs", type=int, default=4)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--policy", default="random", choices=["random", "zero"])
    ap.add_argument("--out", default="trajectory.npz")
    a = ap.parse_args()
    from mocca_envs_b200 import make

    env = make(a.env, num_envs=a.envs, seed=a.seed)
    env.reset()
    A = env.act_dim
    g = torch.Generator(device=env.device).manual_seed(a.seed)
    states, actions, rewards, dones, terrains = [env.get_state().cpu().numpy()], [], [], [], []

    def terrain():
        if hasattr(env, "terrain_info"):
            return env.terrain_info().cpu().numpy()
        return None

    terrains.append(terrain())
    for _ in range(a.steps):
        act = torch.zeros(a.envs, A, device=env.device) if a.policy == "zero" else \
            torch.rand(a.envs, A, device=env.device, generator=g) * 2 - 1
        _, rew, done, _ = env.step(act)
        states.append(env.get_state().cpu().numpy())
        actions.append(act.cpu().numpy())
        rewards.append(rew.cpu().numpy())
        dones.append(done.cpu().numpy())
        terrains.append(terrain())
    out = dict(env_id=np.array(a.env), dt=np.array(env.physics.dt * env.physics.substeps), states=np.stack(states),
               actions=np.stack(actions), rewards=np.stack(rewards), dones=np.stack(dones),
               joint_names=np.array(env.table.get("joint_names", [])))
    if terrains[0] is not None:
        out["terrain"] = np.stack(terrains)  # [T + 1, N, ...]: the table changes when an env resets
    np.savez_compressed(a.out, **out)
    print("wrote %s: %d frames x %d envs, state dim %d" % (a.out, len(states), a.envs, states[0].shape[1]))
    env.close()


if __name__ == "__main__":
    main()
