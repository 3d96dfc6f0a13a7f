#!/usr/bin/env python
"""Replay a trajectory written by tools/export_trajectory.py in the reference's own PyBullet scene (GUI or DIRECT).
Runs ONLY on a host that has `pybullet`, `gym<=0.21` and the reference importable -- none of which exist in the build
container, so this script is untested here (like tools/gen_pybullet_golden.py).  It performs no physics: every frame
it writes the exported base pose and joint angles into the reference env's robot with resetBasePositionAndOrientation
/ resetJointState (the calls the reference itself uses, robots.py:212-227, bullet_utils.py:157-175,266-285) and, for
the stepping-stone env, places the planks through the env's own set_step_state.

  python tools/replay_pybullet.py traj.npz [--env-index 0] [--gui] [--fps 60]

With --check it also steps the reference env with the exported actions from the exported first state and prints the
per-frame state difference -- a rollout-level parity probe."""
import argparse
import sys
import time

import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("file")
    ap.add_argument("--env-index", type=int, default=0)
    ap.add_argument("--gui", action="store_true")
    ap.add_argument("--fps", type=float, default=60.0)
    ap.add_argument("--check", action="store_true")
    a = ap.parse_args()
    try:
        import gym
        import pybullet  # noqa: F401
    except ImportError as e:  # pragma: no cover
        sys.exit("replay_pybullet: %s -- run this on a host with pybullet and gym<=0.21 installed" % e)
    d = np.load(a.file)
    env_id = str(d["env_id"])
    env = gym.make("mocca_envs:" + env_id, render=a.gui)
    u = env.unwrapped
    env.reset()
    p, robot = u._p, u.robot
    states = d["states"][:, a.env_index]
    A = (states.shape[1] - 13) // 2

    def put(s):
        p.resetBasePositionAndOrientation(robot.id if hasattr(robot, "id") else robot.object_id[0], s[0:3], s[3:7])
        robot.reset_joint_states(s[13:13 + A], s[13 + A:13 + 2 * A])
        robot.robot_body.reset_velocity(s[10:13], s[7:10])

    def get():
        pos, quat = robot.robot_body.pose().xyz(), robot.robot_body.pose().orientation()
        q = np.array([j.get_position() for j in robot.ordered_joints])
        return np.concatenate([pos, quat, q])

    for t, s in enumerate(states):
        if "terrain" in d and hasattr(u, "set_step_state"):
            u.terrain_info = d["terrain"][t, a.env_index].astype(np.float64)
            for k in range(min(u.rendered_step_count, len(u.steps))):
                u.set_step_state(k, k)
        put(s)
        if a.check and t + 1 < len(states) and not d["dones"][t, a.env_index]:
            env.step(d["actions"][t, a.env_index].astype(np.float64))
            ref = get()
            mine = np.concatenate([states[t + 1][0:7], states[t + 1][13:13 + A]])
            print("frame %4d  max |pybullet - exported| = %.3e" % (t, np.abs(ref - mine).max()))
        if a.gui:
            time.sleep(1.0 / a.fps)
    env.close()


if __name__ == "__main__":
    main()
