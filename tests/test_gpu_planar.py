"""SURVEY 8 f3 on the device: the planar walkers Walker2DCustomEnv-v0 / Crab2DCustomEnv-v0 through the C ABI against
the CPU oracle (CPU twins on the emulated kernel source: test_planar_emulation.py)."""
import numpy as np
import pytest

from tests.helpers import force_oracle_state, oracle_record, oracle_state

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_mod():
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


def _make(name, n, seed=0, **kw):
    from mocca_envs_b200.vec_env import Crab2DCustomVecEnv, Walker2DCustomVecEnv

    return {"walker2d": Walker2DCustomVecEnv, "crab2d": Crab2DCustomVecEnv}[name](n, device="cuda:0", seed=seed, **kw)


def _table(name, walker2d_table, crab2d_table):
    return {"walker2d": walker2d_table, "crab2d": crab2d_table}[name]


@pytest.mark.parametrize("name", ["walker2d", "crab2d"])
def test_mass_matrix_and_inverse_dynamics(name, walker2d_table, crab2d_table, oracle_mod, torch_mod):
    """north_star bound (1e-4 relative) on in-plane states."""
    from tests.test_planar_emulation import _planar_states

    torch, O, t = torch_mod, oracle_mod, _table(name, walker2d_table, crab2d_table)
    A = t["n_dof"]
    m = O.model_from_table(t)
    rng = np.random.RandomState(0)
    N = 8
    st = _planar_states(t, rng, N).astype(np.float32)
    env = _make(name, N)
    assert env.obs_dim == 6 + 2 * A + 2 + 2 and env.act_dim == A
    env.set_state(torch.tensor(st))
    M = env.mass_matrix().cpu().numpy()
    acc = rng.randn(N, 6 + A).astype(np.float32)
    tau = env.inverse_dynamics(torch.tensor(acc)).cpu().numpy()
    for i in range(N):
        s = oracle_state(O, A, st[i].astype(np.float64))
        Mo = O.mass_matrix(m, s)
        assert np.abs(M[i] - Mo).max() / np.abs(Mo).max() < 1e-4
        ref = O.rnea(m, s, acc[i].astype(np.float64), 9.8)
        assert np.abs(tau[i] - ref).max() / np.abs(ref).max() < 1e-4
    env.close()


@pytest.mark.parametrize("name", ["walker2d", "crab2d"])
def test_reset_and_env_step(name, walker2d_table, crab2d_table, oracle_mod, torch_mod):
    """Reset bit-exact (zero pose +- noise, pelvis origin at the world origin, zeros in the target slots), then
    teacher-forced env steps within 1e-3 / 1e-2 (outliers explained-or-fatal, tests/teacher.py), done never set, state exactly planar."""
    torch, O, t = torch_mod, oracle_mod, _table(name, walker2d_table, crab2d_table)
    A = t["n_dof"]
    N = 16
    env = _make(name, N, seed=60, return_final_obs=True)
    oracles = [O.Walker3DCustomOracle(t, seed=60 + i) for i in range(N)]
    obs = env.reset().cpu().numpy()
    st = env.get_state().cpu().numpy()
    for i, o in enumerate(oracles):
        oref = o.reset()
        assert np.array_equal(st[i, 13:13 + A], np.array(o.e.s.q[:A]).astype(np.float32))
        assert np.array_equal(st[i, 0:7], np.array([0, 0, 0, 0, 0, 0, 1], dtype=np.float32))
        assert obs[i, -1] == 0.0 and obs[i, -2] == 0.0
        assert np.abs(obs[i] - oref).max() < 1e-5
    env.close()
    from tests import teacher as T

    seen = {"contacts": 0}

    def planar(tt, backend, o, done, d1):
        out = backend.env.get_state().cpu().numpy()
        assert not done and not d1
        assert np.all(out[:, [1, 3, 5, 7, 9, 11]] == 0.0)
        seen["contacts"] += int(o.e.feet_contact[0] + o.e.feet_contact[1])

    js = T.run_vs_oracle(O, name, "gpu", range(60, 68), 60, lambda rng, k: rng.uniform(-1.2, 1.2, A), on_step=planar)
    assert seen["contacts"] > 0
    assert np.median(np.concatenate([j.errs for j in js])) < 5e-4


@pytest.mark.parametrize("name", ["walker2d", "crab2d"])
def test_full_size_properties(name, torch_mod):
    """16 384 envs, 1 005 steps of random actions: everything finite, every env exactly planar at every check, the
    only episode end is the TimeLimit at step 1000 (all envs at once, truncated), auto-reset afterwards."""
    torch = torch_mod
    N = 16384
    env = _make(name, N, seed=5)
    env.reset()
    g = torch.Generator(device="cuda:0").manual_seed(0)
    ndone = 0
    for k in range(1005):
        a = torch.rand(N, env.act_dim, device="cuda:0", generator=g) * 2 - 1
        obs, rew, done, info = env.step(a)
        if k == 999:
            assert bool(done.all()) and bool(info["TimeLimit.truncated"].all())
        else:
            ndone += int(done.sum().item()) if k % 100 == 0 or k > 995 else 0
        if k % 250 == 0 or k == 1004:
            st = env.get_state()
            assert torch.isfinite(st).all() and torch.isfinite(obs).all() and torch.isfinite(rew).all()
            assert bool((st[:, [1, 3, 5, 7, 9, 11]] == 0).all())
    assert ndone == 0
    stats = env.stats()
    assert stats["episodes"] == N and stats["nonfinite"] == 0
    env.close()


def test_gym_facade_and_mirror_indices(torch_mod):
    from mocca_envs_b200 import make

    env = make("mocca_envs:Walker2DCustomEnv-v0", seed=0)
    obs = env.reset()
    assert obs.shape == (24,) and obs.dtype == np.float64 and obs[-1] == 0.0 and obs[-2] == 0.0
    o, r, d, info = env.step(np.zeros(7))
    assert o.shape == (24,) and not d and np.isfinite(r)
    neg, right, left, neg_a, right_a, left_a = env.get_mirror_indices()
    assert list(right_a) == [1, 2, 3] and list(left_a) == [4, 5, 6] and len(neg_a) == 0  # robots.py:365-369
    assert list(right) == [7, 8, 9, 14, 15, 16, 20] and list(left) == [10, 11, 12, 17, 18, 19, 21]
    assert list(neg) == [2, 4, 22]
    crab = make("Crab2DCustomEnv-v0", num_envs=3, seed=1)
    assert crab.reset().shape == (3, 22)
    crab.close()
