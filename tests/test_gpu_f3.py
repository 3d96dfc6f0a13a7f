"""SURVEY 8 f3 on the device: Child3DCustomEnv-v0 and MikeStepperEnv-v0 through the C ABI against the CPU oracle
(the CPU twins of these tests, on the emulated kernel source, are in test_f3_emulation.py)."""
import numpy as np
import pytest

from tests.helpers import force_oracle_state, oracle_record, oracle_state, random_states

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_mod():
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


def _make(name, n, seed=0, **kw):
    from mocca_envs_b200.vec_env import Child3DCustomVecEnv, MikeStepperVecEnv

    return {"child": Child3DCustomVecEnv, "mike": MikeStepperVecEnv}[name](n, device="cuda:0", seed=seed, **kw)


@pytest.mark.parametrize("name", ["child", "mike"])
def test_mass_matrix_and_inverse_dynamics(name, child_table, mike_table, oracle_mod, torch_mod):
    """north_star: mass matrix and inverse dynamics within 1e-4 relative, for the two extra model tables."""
    torch, O, t = torch_mod, oracle_mod, {"child": child_table, "mike": mike_table}[name]
    A = t["n_dof"]
    m = O.model_from_table(t)
    rng = np.random.RandomState(0)
    N = 8
    st = random_states(t, rng, N).astype(np.float32)
    env = _make(name, N)
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


def test_child_reset_and_env_step(child_table, oracle_mod, torch_mod):
    """Child3DCustomEnv: crawl start pose bit-exact (base pitched by 90 degrees at z = 0.38), then teacher-forced
    env steps.  Tolerance 1e-3 / 1e-2; every step outside it must be an explained, bounded discontinuity
    (the child's tiny inertias make its mass matrix numerically singular more often: verified per step, tests/teacher.py), median < 5e-4."""
    torch, O, t = torch_mod, oracle_mod, child_table
    N = 16
    env = _make("child", N, seed=40, return_final_obs=True)
    oracles = [O.Walker3DCustomOracle(t, seed=40 + i) for i in range(N)]
    obs = env.reset().cpu().numpy()
    st = env.get_state().cpu().numpy()
    for i, o in enumerate(oracles):
        oref = o.reset()
        assert np.array_equal(st[i, 13:34], np.array(o.e.s.q[:21]).astype(np.float32))
        assert np.array_equal(st[i, 0:3], np.array([0, 0, 0.38], dtype=np.float32))
        assert np.allclose(st[i, 3:7], [0, np.sin(np.pi / 4), 0, np.cos(np.pi / 4)], atol=1e-7)
        assert np.abs(obs[i] - oref).max() < 1e-5
    env.close()
    from tests import teacher as T

    js = T.run_vs_oracle(O, "child3d", "gpu", range(40, 48), 30, lambda rng, k: rng.uniform(-1.2, 1.2, 21))
    assert np.median(np.concatenate([j.errs for j in js])) < 5e-4


def test_mike_reset_and_env_step(mike_table, oracle_mod, torch_mod):
    """MikeStepperEnv: start at (0.3, 0, 1.0), terrain bit-exact, then teacher-forced env steps on the planks
    (1e-3 / 1e-2; every step outside must be an explained, bounded discontinuity: tests/teacher.py)."""
    torch, O, t = torch_mod, oracle_mod, mike_table
    N = 12
    curs = [0, 5, 9] * 4
    env = _make("mike", N, seed=300, return_final_obs=True)
    env.set_env_params({"curriculum": curs})
    oracles = [O.Walker3DStepperOracle(t, seed=300 + i, curriculum=curs[i]) for i in range(N)]
    obs = env.reset().cpu().numpy()
    st = env.get_state().cpu().numpy()
    ter = env.terrain_info().cpu().numpy()
    for i, o in enumerate(oracles):
        oref = o.reset()
        assert np.array_equal(ter[i], np.array(o.e.terrain[:]).astype(np.float32))
        assert np.array_equal(st[i, 0:3], np.array([0.3, 0.0, 1.0], dtype=np.float32))
        assert np.abs(obs[i] - oref).max() < 1e-5
    env.close()
    from tests import teacher as T

    errs = []
    for i in range(6):
        js = T.run_vs_oracle(O, "mike", "gpu", [300 + i], 40, lambda rng, k: 0.3 * rng.uniform(-1, 1, 21),
                             curriculum=[0, 5, 9][i % 3])
        errs += js[0].errs
    assert np.median(errs) < 5e-4


@pytest.mark.parametrize("name", ["child", "mike"])
def test_full_size_properties(name, torch_mod):
    """16384 envs: finite outputs, no cap overflows, determinism across two identically seeded batches."""
    torch = torch_mod
    N = 16384
    outs = []
    for rep in range(2):
        env = _make(name, N, seed=11)
        env.reset()
        g = torch.Generator(device="cuda:0").manual_seed(5)
        for _ in range(12):
            obs, rew, done, info = env.step(torch.rand(N, 21, device="cuda:0", generator=g) * 2 - 1)
        assert torch.isfinite(obs).all() and torch.isfinite(rew).all()
        st = env.stats()
        assert st["nonfinite"] == 0
        outs.append((obs.clone(), rew.clone(), done.clone()))
        env.close()
    assert all(torch.equal(a, b) for a, b in zip(outs[0], outs[1]))


def test_gym_facades_and_kwargs(torch_mod):
    """make() ids and the reference's constructor kwargs: Child3D / Mike facades return float64 NumPy like the
    reference; random_reward / plank_class reach the stepper-family envs; an unknown plank class is refused."""
    from mocca_envs_b200 import make

    for eid, obs_dim in (("mocca_envs:Child3DCustomEnv-v0", 52), ("MikeStepperEnv-v0", 65)):
        env = make(eid, seed=3)
        o = env.reset()
        assert o.dtype == np.float64 and o.shape == (obs_dim,)
        o, r, d, info = env.step(np.zeros(21))
        assert o.shape == (obs_dim,) and isinstance(r, float) and isinstance(d, bool)
        env.close()
    env = make("MikeStepperEnv-v0", seed=3, random_reward=True, plank_class="Plank")
    assert env.vec.random_reward and env.vec.plank_class == "Plank"
    env.reset()
    env.step(np.zeros(21))
    env.close()
    env = make("MikeStepperEnv-v0", seed=3, plank_class="Pillar")  # cylinder stones (bullet_objects.py:86-90)
    assert env.vec.plank_class == "Pillar"
    o = env.reset()
    for _ in range(5):
        o, r, d, info = env.step(np.zeros(21))
    assert np.isfinite(o).all() and np.isfinite(r)
    env.close()
    with pytest.raises(ValueError):
        make("Walker3DStepperEnv-v0", num_envs=2, plank_class="Pilar")


def test_env_param_accessors(torch_mod):
    """EnvBase.get_env_param / set_robot_params (env_base.py:108-117): the attribute mirror and quirk Q4."""
    from mocca_envs_b200 import make

    env = make("Walker3DStepperEnv-v0", num_envs=4, seed=0)
    assert env.get_env_param("curriculum", -1) == 0 and env.get_env_param("max_curriculum", -1) == 9
    env.set_env_params({"curriculum": 7})
    assert env.get_env_param("curriculum", -1) == 7
    assert env.get_env_param("no_such_param", "dflt") == "dflt"
    with pytest.raises(AttributeError):
        env.set_robot_params({"power": 0.5})
    env.close()


@pytest.mark.gpu
@pytest.mark.parametrize("env_id,kwargs", [("Walker3DStepperEnv-v0", {}), ("Walker3DStepperEnv-v0", {"plank_class": "Pillar"}),
                                           ("Monkey3DCustomEnv-v0", {}), ("CassieEnv-v0", {}),
                                           ("Child3DCustomEnv-v0", {}), ("MikeStepperEnv-v0", {}),
                                           ("Walker2DCustomEnv-v0", {}), ("Crab2DCustomEnv-v0", {})])
def test_step_host_pinned_matches_device_every_env(env_id, kwargs):
    """The `_host` instantiation of every step kernel (zero-copy result stores into pinned host buffers, actions read
    from the pinned host buffer in place) returns exactly what the device-buffer step returns."""
    import torch
    from mocca_envs_b200 import make

    N = 200
    e1, e2 = make(env_id, num_envs=N, seed=9, **kwargs), make(env_id, num_envs=N, seed=9, **kwargs)
    e1.reset(); e2.reset()
    A, OB = e1.act_dim, e1.obs_dim
    h_act = torch.empty(N, A).pin_memory()
    outs = tuple(t.numpy() for t in (torch.empty(N, OB).pin_memory(), torch.empty(N).pin_memory(),
                                     torch.empty(N, dtype=torch.uint8).pin_memory(),
                                     torch.empty(N, dtype=torch.uint8).pin_memory()))
    rng = np.random.RandomState(3)
    for _ in range(12):
        a = rng.uniform(-1, 1, (N, A)).astype(np.float32)
        h_act.numpy()[:] = a
        o1, r1, d1, _ = e1.step(torch.tensor(a))
        for o in outs:
            o.fill(0)
        e2.step_host(h_act.numpy(), outs)
        assert np.array_equal(o1.cpu().numpy(), outs[0]) and np.array_equal(r1.cpu().numpy(), outs[1])
        assert np.array_equal(d1.cpu().numpy(), outs[2])
    e1.close(); e2.close()
