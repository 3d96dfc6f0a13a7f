"""GPU parity tests for Monkey3DCustomEnv-v0 (BASELINE config 5) through the C ABI, against the CPU oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_mod():
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


def _env(n, seed=0, **kw):
    from mocca_envs_b200.vec_env import Monkey3DCustomVecEnv

    return Monkey3DCustomVecEnv(n, device="cuda:0", seed=seed, **kw)


def test_monkey_mass_matrix_and_inverse_dynamics(monkey_table, oracle_mod, torch_mod):
    """north_star: mass matrix and inverse dynamics within 1e-4 relative error (Monkey3D, 29 generalised coords)."""
    from tests.helpers import oracle_state, random_states

    torch, O, t = torch_mod, oracle_mod, monkey_table
    A, N = t["n_dof"], 16
    m = O.model_from_table(t)
    rows = random_states(t, np.random.RandomState(0), N)
    env = _env(N)
    env.set_state(torch.tensor(rows, dtype=torch.float32))
    M = env.mass_matrix().cpu().numpy()
    acc = np.random.RandomState(1).randn(N, 6 + A)
    tau = env.inverse_dynamics(torch.tensor(acc, dtype=torch.float32)).cpu().numpy()
    for i in range(N):
        s = oracle_state(O, A, rows[i].astype(np.float32).astype(np.float64))
        Mref = O.mass_matrix(m, s)
        assert np.abs(M[i] - Mref).max() / np.abs(Mref).max() < 1e-4
        tref = O.rnea(m, s, acc[i].astype(np.float32).astype(np.float64), 9.8)
        assert np.abs(tau[i] - tref).max() / np.abs(tref).max() < 1e-4
    env.close()


def test_monkey_reset_and_bars(monkey_table, oracle_mod, torch_mod):
    """Reset state bit-exact; bar layout from the same seed equal after f32 rounding (rows pinned to the hands
    inherit the f32 forward kinematics at 20 m altitude: tolerance 2e-5)."""
    torch, O, t = torch_mod, oracle_mod, monkey_table
    N = 8
    env = _env(N, seed=100)
    oracles = [O.Monkey3DOracle(t, seed=100 + i) for i in range(N)]
    for _ in range(2):
        obs = env.reset().cpu().numpy()
        st = env.get_state().cpu().numpy()
        ter = env.terrain_info().cpu().numpy()
        rec = env.get_record().cpu().numpy().view(np.int32)
        for i, o in enumerate(oracles):
            oref = o.reset()
            ref = np.array([list(r) for r in o.e.terrain]).astype(np.float32)
            assert np.abs(ter[i] - ref).max() < 2e-5
            assert np.array_equal(st[i, 13:36], np.array(o.e.base.s.q[:23]).astype(np.float32))
            assert np.array_equal(st[i, 0:3], np.array([0.0, 0.0, 20.0], dtype=np.float32))
            assert (rec[i, env.EM_SWING], rec[i, env.EM_PIVOT]) == (o.e.swing_leg, o.e.pivot_leg)
            assert np.abs(obs[i] - oref).max() < 2e-5
    env.close()


def test_monkey_env_step_teacher_forced(oracle_mod, torch_mod):
    """Monkey3DCustomEnv.step on the device from f32-identical states and bookkeeping: hand / palm contacts with the
    bars, scripted finger joints, swing progress, free-fall termination; 12 envs x 50 steps.  Steps outside 1e-3 (obs) / 1e-2 (reward) must be explained by a verified
    discontinuity and bounded (tests/teacher.py), else the test fails; integer bookkeeping read back and compared
    exactly after every structurally identical step."""
    from tests import teacher as T

    js = T.run_vs_oracle(oracle_mod, "monkey", "gpu", range(300, 312), 50, lambda rng, k: 0.5 * rng.uniform(-1, 1, 23))
    assert np.median(np.concatenate([j.errs for j in js])) < 3e-4


def test_monkey_rollout_statistics(monkey_table, oracle_mod, torch_mod):
    """Free-running random-policy episodes: length / return / bars reached of 512 GPU envs vs 48 oracle envs agree
    within 4 standard errors (north_star: statistically indistinguishable episode return and length)."""
    torch, O, t = torch_mod, oracle_mod, monkey_table
    N, A, T = 512, 23, 260
    env = _env(N, seed=900)
    env.reset()
    g = torch.Generator(device="cuda:0").manual_seed(1)
    for _ in range(T):
        env.step(torch.rand(N, A, device="cuda:0", generator=g) * 2 - 1)
    s = env.stats()
    assert s["nonfinite"] == 0
    gl, gr = s["length_sum"] / s["episodes"], s["return_sum"] / s["episodes"]
    lens, rets = [], []
    rng = np.random.RandomState(2)
    for i in range(48):
        o = O.Monkey3DOracle(t, seed=5000 + i)
        o.reset()
        L, R = 0, 0.0
        while True:
            _, r, d, _ = o.step(rng.uniform(-1, 1, A))
            L += 1
            R += r
            if d:
                break
        lens.append(L)
        rets.append(R)
    lens, rets = np.array(lens), np.array(rets)
    assert abs(gl - lens.mean()) < 4 * lens.std() / np.sqrt(len(lens)) + 1.0, (gl, lens.mean(), lens.std())
    assert abs(gr - rets.mean()) < 4 * rets.std() / np.sqrt(len(rets)) + 1.0, (gr, rets.mean(), rets.std())
    env.close()


def test_monkey_gym_facade(torch_mod):
    """N=1 gym protocol: float64 obs of length 69, unbounded action space, in-place finger override (quirk Q11)."""
    import mocca_envs_b200 as mb

    env = mb.make("mocca_envs:Monkey3DCustomEnv-v0", seed=3)
    obs = env.reset()
    assert obs.shape == (69,) and obs.dtype == np.float64
    assert np.isinf(env.action_space.high).all()
    a = np.zeros(23)
    o, r, d, info = env.step(a)
    assert sorted([a[17], a[22]]) == [-1.0, 1.0]
    assert o.shape == (69,) and isinstance(r, float) and isinstance(d, bool)
    env.close()
