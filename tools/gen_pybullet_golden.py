#!/usr/bin/env python
"""Golden-vector generator.  Runs ONLY on a host that has `pybullet` and `gym<=0.21` with the
reference importable (`pip install pybullet gym==0.21 && pip install -e /path/to/mocca_envs`); neither exists in
the build container, so no golden file is committed yet and every PyBullet-parity statement in this repository
reads "vs restatement".  Output: tests/golden/walker3d_pybullet_<version>.npz, consumed by
tests/test_pybullet_golden.py when present.

Dumps: joint/link tables (OQ1-OQ4), M(q) and inverse dynamics at 16 random states (G1), contact-free and in-contact
single steps (G2, G3), a 1000-step random-action trace for seed 0 (G4) and reset states for seeds 0..15 (G6).
"""
import os
import sys

import numpy as np


def main(out_dir=os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")):
    try:
        import gym
        import pybullet as pb
    except ImportError as e:  # pragma: no cover
        sys.exit("gen_pybullet_golden: %s -- run this on a host with pybullet and gym<=0.21 installed" % e)
    env = gym.make("mocca_envs:Walker3DCustomEnv-v0")
    u = env.unwrapped
    p, rid = u._p, u.robot.id
    nj = p.getNumJoints(rid)
    out = {"pybullet_api": np.array(p.getAPIVersion())}
    out["joint_info"] = np.array([str(p.getJointInfo(rid, j)) for j in range(nj)])
    out["dynamics_info"] = np.array([str(p.getDynamicsInfo(rid, l)) for l in range(-1, nj)])
    rng = np.random.RandomState(0)
    ids = u.robot.ordered_joint_ids
    lo = np.array([j.lowerLimit for j in u.robot.ordered_joints])
    hi = np.array([j.upperLimit for j in u.robot.ordered_joints])
    Ms, states, steps = [], [], []
    for k in range(16):
        q = lo + (hi - lo) * rng.uniform(0.35, 0.65, len(ids))
        qd = rng.uniform(-1, 1, len(ids))
        quat = rng.randn(4)
        quat /= np.linalg.norm(quat)
        pos = [0.0, 0.0, 3.0]
        w, v = 0.5 * rng.randn(3), rng.randn(3)
        p.resetBasePositionAndOrientation(rid, pos, quat.tolist())
        p.resetBaseVelocity(rid, v.tolist(), w.tolist())
        for j, a, b in zip(ids, q, qd):
            p.resetJointState(rid, j, a, b)
        full_q = [0.0] * nj
        for j, a in zip(ids, q):
            full_q[j] = a
        Ms.append(np.array(p.calculateMassMatrix(rid, [full_q[j] for j in ids])))
        tau = rng.uniform(-1, 1, len(ids)) * u.robot.ordered_joint_base_gains
        p.setJointMotorControlArray(rid, ids, p.TORQUE_CONTROL, forces=tau.tolist())
        before = np.concatenate([pos, quat, w, v, q, qd])
        p.stepSimulation()
        bp, bq = p.getBasePositionAndOrientation(rid)
        bv, bw = p.getBaseVelocity(rid)
        js = p.getJointStates(rid, ids)
        after = np.concatenate([bp, bq, bw, bv, [s[0] for s in js], [s[1] for s in js]])
        states.append(before)
        steps.append(np.concatenate([tau, after]))
    out["mass_matrix"] = np.array(Ms)
    out["free_states"] = np.array(states)
    out["free_steps"] = np.array(steps)
    env.seed(0)
    trace = [env.reset()]
    arng = np.random.RandomState(1)
    rews, dones = [], []
    for t in range(1000):
        o, r, d, _ = env.step(arng.uniform(-1, 1, 21))
        trace.append(o)
        rews.append(r)
        dones.append(d)
        if d:
            trace.append(env.reset())
    out["trace_obs"] = np.array(trace)
    out["trace_rew"] = np.array(rews)
    out["trace_done"] = np.array(dones)
    os.makedirs(out_dir, exist_ok=True)
    path = os.path.join(out_dir, "walker3d_pybullet_%s.npz" % p.getAPIVersion())
    np.savez_compressed(path, **out)
    print("wrote", path)


if __name__ == "__main__":
    main(*sys.argv[1:])
