#!/usr/bin/env python
"""BASELINE.json configs[0] / SURVEY 8d "Config 1": Walker3DCustomEnv-v0, ONE env, random actions, 1000 control steps.

The reference would record this trace with PyBullet DIRECT on the CPU; PyBullet cannot be installed here, so the trace
is recorded from the float64 ORACLE (oracle/mocca_oracle.c) and labelled as such: `source = "restatement"`.  Same file
layout as the reference-recorded fixtures (tests/golden/ref_*.npz, tools/gen_reference_golden.py), so the same
teacher-forced harness (tests/teacher.py) drives the kernel source and the device along it.  When a host with pybullet
is available, tools/gen_pybullet_golden.py records the same 1000 actions through the reference itself (G4).

usage: python tools/gen_config1_trace.py [out.npz]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def record(steps=1000, seed=0, action_seed=0):
    from mocca_envs_b200.model_compiler import load_table
    from oracle import oracle as O

    O.build()
    t = load_table(os.path.join(ROOT, "mocca_envs_b200", "models", "walker3d.json"))
    o = O.Walker3DCustomOracle(t, seed=seed)
    o.seed(seed)  # gym.make(...) then env.seed(seed): quirk Q1, like the reference-recorded fixtures
    rng = np.random.RandomState(action_seed)
    obs, states, rewards, dones, actions, resets = [o.reset()], [o.state_vector()], [], [], [], [0]
    for k in range(steps):
        a = rng.uniform(-1, 1, 21)
        ob, r, d, _ = o.step(a)
        actions.append(a); obs.append(ob); rewards.append(r); dones.append(d); states.append(o.state_vector())
        if d:
            obs.append(o.reset())
            states.append(o.state_vector())
            resets.append(len(obs) - 1)
    return dict(source="restatement (float64 oracle, oracle/mocca_oracle.c); NOT PyBullet output",
                config="BASELINE.json configs[0]: Walker3DCustomEnv-v0 single env, random actions U(-1,1)^21, 1000 steps",
                seed=seed, construction_seed=seed, action_seed=action_seed, eval_mode=0,
                actions=np.array(actions), obs=np.array(obs), states=np.array(states), rewards=np.array(rewards),
                dones=np.array(dones), resets=np.array(resets))


if __name__ == "__main__":
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "restatement_walker3d_custom_config1.npz")
    d = record()
    np.savez_compressed(out, **d)
    print(out, "steps", len(d["actions"]), "episodes", int(d["dones"].sum()), "bytes", os.path.getsize(out))
