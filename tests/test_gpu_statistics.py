"""north_star: "episode return and length for fixed scripted and random policies statistically indistinguishable
(stated bound)".  The random policy on Walker3DCustom / Stepper / Monkey3D is covered next to each env's other GPU
tests; this file adds the SCRIPTED policy (the bench's PD controller toward the running-start pose)
on Walker3DCustomEnv and Walker3DStepperEnv, and CassieEnv under random residual targets with >= 1000 oracle episodes.

Sampling scheme (the same on both sides, so that neither is biased towards short episodes): the FIRST episode of every
env of a batch.  Stated bound: the means of episode length and return differ by less than 4 pooled standard errors."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _pd_reference(t):
    lo, hi = np.array(t["lower"]), np.array(t["upper"])
    return 2 * (np.array(t["base_joint_angles"]) - lo) / (hi - lo) - 1


def _oracle_first_episodes(O, kind, table, n, policy, seed0, max_steps, threads=0):
    """(lengths, returns) of the first episode of n oracle envs stepped as one OpenMP batch."""
    import os

    struct, pre, ob = {"custom": (O.W3DEnv, "orc_w3d", 52), "stepper": (O.StepperEnv, "orc_stepper", 65),
                       "cassie": (O.CassieEnvS, "orc_cassie", 36)}[kind]
    m = O.model_from_table(table)
    p = O.cassie_params() if kind == "cassie" else O.default_params()
    L = O.lib()
    envs = (struct * n)()
    obs = np.zeros((n, ob))
    for i in range(n):
        if kind != "cassie":
            w = O.gym_seed_words(seed0 + i)
            key = (C.c_uint32 * len(w))(*w)
            if kind == "stepper":
                envs[i].curriculum = (0, 5, 9)[i % 3]
            getattr(L, pre + "_seed")(C.byref(envs[i]), key, len(w), 1)
        getattr(L, pre + "_reset")(C.byref(m), C.byref(p), C.byref(envs[i]), obs[i].ctypes.data_as(C.c_void_p))
    rew, done = np.zeros(n), np.zeros(n, dtype=np.int32)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    lens, rets = np.zeros(n, dtype=np.int64), np.zeros(n)
    alive = np.ones(n, dtype=bool)
    for k in range(max_steps):
        a = np.ascontiguousarray(policy(k, obs), dtype=np.float64)
        getattr(L, pre + "_step_batch")(C.byref(m), C.byref(p), envs, n, vp(a), vp(obs), vp(rew), vp(done),
                                        threads or (os.cpu_count() or 1))
        lens[alive] += 1
        rets[alive] += rew[alive]
        alive &= done == 0
        if not alive.any():
            break
    return lens, rets, alive


def _device_first_episodes(env, policy, max_steps):
    import torch

    n = env.num_envs
    obs = env.reset()
    lens = torch.zeros(n, dtype=torch.int64, device=env.device)
    rets = torch.zeros(n, dtype=torch.float64, device=env.device)
    alive = torch.ones(n, dtype=torch.bool, device=env.device)
    for k in range(max_steps):
        obs, rew, done, _ = env.step(policy(k, obs))
        lens += alive
        rets += torch.where(alive, rew.double(), torch.zeros_like(rets))
        alive &= ~done.bool()
        if not bool(alive.any()):
            break
    return lens.cpu().numpy(), rets.cpu().numpy(), alive.cpu().numpy()


def _compare(dev, orc, what):
    (dl, dr, da), (ol, orr, oa) = dev, orc
    assert not da.any() and not oa.any(), "%s: episodes still running at the step bound" % what
    for name, a_, b_ in (("length", dl.astype(np.float64), ol.astype(np.float64)), ("return", dr, orr)):
        se = np.sqrt(a_.var() / len(a_) + b_.var() / len(b_))
        assert abs(a_.mean() - b_.mean()) < 4 * se + 1e-9, (what, name, a_.mean(), b_.mean(), se)


@pytest.mark.parametrize("kind", ["custom", "stepper"])
def test_scripted_pd_policy_statistics(kind, walker_table, oracle_mod):
    """Scripted PD toward the running-start pose (kp = 1, kd = 0.1 in normalised joint units, the bench's --actions pd):
    4096 device episodes vs 1024 oracle episodes of Walker3DCustomEnv / Walker3DStepperEnv (curriculum 0 / 5 / 9)."""
    import torch

    from mocca_envs_b200.vec_env import Walker3DCustomVecEnv, Walker3DStepperVecEnv

    O, t = oracle_mod, walker_table
    A = 21
    ref = _pd_reference(t)
    orc = _oracle_first_episodes(O, kind, t, 1024, lambda k, obs: np.clip(
        1.0 * (ref - obs[:, 6:6 + A]) - 0.1 * (obs[:, 6 + A:6 + 2 * A] * 10.0), -1, 1), 7000, 1000)
    N = 4096
    env = (Walker3DCustomVecEnv if kind == "custom" else Walker3DStepperVecEnv)(N, device="cuda:0", seed=90000)
    if kind == "stepper":
        env.set_env_params({"curriculum": np.array([0, 5, 9] * (N // 3 + 1))[:N]})
    ref_t = torch.tensor(ref, device="cuda:0", dtype=torch.float32)
    dev = _device_first_episodes(env, lambda k, obs: torch.clamp(
        1.0 * (ref_t - obs[:, 6:6 + A]) - 0.1 * (obs[:, 6 + A:6 + 2 * A] * 10.0), -1, 1), 1000)
    assert env.stats()["nonfinite"] == 0
    env.close()
    _compare(dev, orc, "%s PD" % kind)


def test_cassie_random_policy_statistics(cassie_table, oracle_mod):
    """CassieEnv-v0 under random residual PD targets a ~ U(-1, 1)^10 (the robot topples within ~20 env steps = 1000
    substeps): 4096 device episodes vs 1024 oracle episodes."""
    import torch

    from mocca_envs_b200.vec_env import CassieVecEnv

    O, t = oracle_mod, cassie_table
    rng = np.random.RandomState(4)
    orc = _oracle_first_episodes(O, "cassie", t, 1024, lambda k, obs: rng.uniform(-1, 1, (obs.shape[0], 10)), 0, 1000)
    N = 4096
    env = CassieVecEnv(N, device="cuda:0")
    g = torch.Generator(device="cuda:0").manual_seed(6)
    dev = _device_first_episodes(env, lambda k, obs: torch.rand(N, 10, device="cuda:0", generator=g) * 2 - 1, 1000)
    assert env.stats()["nonfinite"] == 0
    env.close()
    _compare(dev, orc, "cassie random")
