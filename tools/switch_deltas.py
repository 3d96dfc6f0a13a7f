#!/usr/bin/env python
"""What each Bullet-version switch does to the rollout statistics: first-episode length and
return of Walker3DCustomEnv-v0 under the random and the scripted-PD policy, float64 oracle, 1024 envs per setting.

  warmstart          multibody contact warm starting (oracle AND kernel: mb200_physics.warmstart)
  limit_rows_always  joint-limit rows created for every limited joint, not only violated ones (Bullet < 2.88; oracle
                     only: 42 always-on rows do not fit the kernel's 48-row budget next to the contacts)
  persistent_manifold  btPersistentManifold semantics against the ground plane: <= 4 cached points per link, refreshed
                     per substep, one new (deepest) point per geom per substep (oracle only; the kernel and the default
                     oracle take every sphere / capsule end within the breaking threshold)

usage: python tools/switch_deltas.py [n_envs]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(n=1024):
    import ctypes as C

    from mocca_envs_b200.model_compiler import load_table
    from oracle import oracle as O

    O.build()
    t = load_table(os.path.join(ROOT, "mocca_envs_b200", "models", "walker3d.json"))
    m = O.model_from_table(t)
    L = O.lib()
    A = 21
    lo, hi = np.array(t["lower"]), np.array(t["upper"])
    ref = 2 * (np.array(t["base_joint_angles"]) - lo) / (hi - lo) - 1

    def run(p, policy):
        envs = (O.W3DEnv * n)()
        obs = np.zeros((n, 52))
        for i in range(n):
            w = O.gym_seed_words(31000 + i)
            key = (C.c_uint32 * len(w))(*w)
            L.orc_w3d_seed(C.byref(envs[i]), key, len(w), 1)
            L.orc_w3d_reset(C.byref(m), C.byref(p), C.byref(envs[i]), obs[i].ctypes.data_as(C.c_void_p))
        rew, done = np.zeros(n), np.zeros(n, dtype=np.int32)
        lens, rets, alive = np.zeros(n), np.zeros(n), np.ones(n, dtype=bool)
        rng = np.random.RandomState(3)
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        for k in range(1000):
            a = rng.uniform(-1, 1, (n, A)) if policy == "random" else np.clip(
                1.0 * (ref - obs[:, 6:6 + A]) - 0.1 * (obs[:, 6 + A:6 + 2 * A] * 10.0), -1, 1)
            a = np.ascontiguousarray(a)
            L.orc_w3d_step_batch(C.byref(m), C.byref(p), envs, n, vp(a), vp(obs), vp(rew), vp(done), os.cpu_count() or 1)
            lens[alive] += 1
            rets[alive] += rew[alive]
            alive &= done == 0
            if not alive.any():
                break
        return lens, rets

    print("%-28s %-8s %12s %12s %14s %14s" % ("setting", "policy", "mean length", "(s.e.)", "mean return", "(s.e.)"))
    for name, kw in (("reference defaults", {}), ("warmstart = 0.85", {"warmstart": 0.85}), ("warmstart = 0.1", {"warmstart": 0.1}),
                     ("limit_rows_always = 1", {"limit_rows_always": 1}), ("persistent_manifold = 1", {"persistent_manifold": 1})):
        for policy in ("random", "pd"):
            p = O.default_params()
            for k, v in kw.items():
                setattr(p, k, v)
            lens, rets = run(p, policy)
            print("%-28s %-8s %12.3f %12.3f %14.3f %14.3f" % (name, policy, lens.mean(), lens.std() / np.sqrt(n), rets.mean(),
                                                               rets.std() / np.sqrt(n)))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 1024)
