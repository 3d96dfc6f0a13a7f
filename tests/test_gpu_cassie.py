"""GPU parity tests for CassieEnv-v0 (BASELINE config 4) through the C ABI, against the CPU oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_mod():
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


def _env(n, **kw):
    from mocca_envs_b200.vec_env import CassieVecEnv

    return CassieVecEnv(n, device="cuda:0", **kw)


def test_cassie_mass_matrix_and_inverse_dynamics(cassie_table, oracle_mod, torch_mod):
    """north_star: mass matrix and inverse dynamics within 1e-4 relative error (Cassie, 24 generalised coords)."""
    from tests.helpers import oracle_state, random_states
    from tests.test_cassie_emulation import _rand_table

    torch, O, t = torch_mod, oracle_mod, cassie_table
    A, N = t["n_dof"], 16
    m = O.model_from_table(t)
    rows = random_states(_rand_table(t), np.random.RandomState(0), N)
    env = _env(N)
    assert (env.obs_dim, env.act_dim, env.state_dim, env.nu) == (36, 10, 49, 24)
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


def test_cassie_airborne_step_with_loop_closures(cassie_table, oracle_mod, torch_mod):
    """One 0.6 ms step away from the ground: forward dynamics + the loop-closure rows; state within 3e-4 (the ERP term
    divides the f32 pivot gap by dt = 0.6 ms: 1e-7 m of rounding is 1.7e-4 m/s)."""
    from tests.helpers import oracle_state, state_error

    torch, O, t = torch_mod, oracle_mod, cassie_table
    A, N = t["n_dof"], 16
    m = O.model_from_table(t)
    p = O.cassie_params()
    rng = np.random.RandomState(1)
    base = np.array(t["base_joint_angles"])
    rows = np.zeros((N, 13 + 2 * A))
    for k in range(N):
        rows[k] = np.concatenate([[0, 0, 3.0], [0, 0, 0, 1], 0.3 * rng.randn(3), 0.3 * rng.randn(3),
                                  base + 0.01 * rng.randn(A), 0.2 * rng.randn(A)])
    rows = rows.astype(np.float32).astype(np.float64)
    taus = 20 * rng.uniform(-1, 1, (N, A))
    env = _env(N)
    env.set_state(torch.tensor(rows, dtype=torch.float32))
    held = taus - np.array(t["damping"]) * rows[:, 13 + A:]
    nrows, nc = env.step_physics(torch.tensor(held, dtype=torch.float32))
    out = env.get_state().cpu().numpy()
    for i in range(N):
        s = oracle_state(O, A, rows[i])
        _, r = O.step_physics(m, p, s, taus[i])
        assert int(nc[i]) == 0 and abs(int(nrows[i]) - r) <= 1
        assert state_error(out[i], O.state_vector(s, A)) < 3e-4
    env.close()


def test_cassie_env_step_teacher_forced(oracle_mod, torch_mod):
    """CassieEnv.step (50 PD substeps, toe contacts, loop closures) on the device from f32-identical states and
    bookkeeping for 8 action streams: the oracle's state and low-pass joint velocities are injected before every step.
    Tolerance 2e-2 (obs; raw joint speeds in rad/s dominate) / 2e-3 (reward), same done; every step outside must be
    explained by a verified discontinuity and bounded (tests/teacher.py), else the test fails."""
    from tests import teacher as T

    T.run_vs_oracle(oracle_mod, "cassie", "gpu", range(8), 25,
                    lambda rng, k: (0.3 if (k // 5) % 2 == 0 else 0.6) * rng.uniform(-1, 1, 10))


def test_cassie_determinism_and_gym_facade(torch_mod):
    import mocca_envs_b200 as mb

    torch = torch_mod
    a = _env(64)
    b = _env(7)
    a.reset()
    b.reset()
    g = torch.Generator(device="cuda:0").manual_seed(0)
    for _ in range(5):
        act = 0.2 * (torch.rand(64, 10, device="cuda:0", generator=g) * 2 - 1)
        oa, ra, da, _ = a.step(act)
        ob, rb, db, _ = b.step(act[:7])
        assert torch.equal(oa[:7], ob) and torch.equal(ra[:7], rb)  # batch-size independence, bit-exact
    a.close()
    b.close()
    env = mb.make("mocca_envs:CassieEnv-v0")
    obs = env.reset()
    assert obs.shape == (36,) and obs.dtype == np.float64
    o, r, d, info = env.step(np.zeros(10))
    assert set(info) >= {"AliveRew", "ProgressRew"} and abs(info["AliveRew"] + info["ProgressRew"] - r) < 1e-6
    env.close()


def test_cassie_hull_self_collision(cassie_table, oracle_mod, torch_mod):
    """Mesh-hull self-collision on the device (warp-cooperative GJK, one hull vertex per lane; env_cassie.py:81-85)
    against the oracle's float64 GJK: contact geometry through mb200_step_physics_points (distance 2e-6, point 2e-6,
    normal 2e-3), contact / row counts, the state after one stepSimulation (5e-3, median 5e-4); with
    physics={"self_collision": 0} the same states disagree (the feature is live on the device)."""
    from tests.helpers import cassie_hull_contact_states, oracle_state, state_error

    torch, O, t = torch_mod, oracle_mod, cassie_table
    A = t["n_dof"]
    m = O.model_from_table(t)
    p = O.cassie_params()
    rng = np.random.RandomState(3)
    N = 32
    st = cassie_hull_contact_states(O, t, rng, N)
    outs = {}
    for flag in (1, 0):
        env = _env(N, physics={"self_collision": flag})
        env.set_state(torch.tensor(st))
        rows, nc, pts = env.step_physics_points(torch.zeros(N, A))
        outs[flag] = (env.get_state().cpu().numpy(), rows.cpu().numpy(), nc.cpu().numpy(), pts.cpu().numpy())
        env.close()
    out, rows, nc, pts = outs[1]
    errs = []
    for i in range(N):
        s = oracle_state(O, A, st[i].astype(np.float64))
        c0 = O.collide(m, p, s)
        c, rows_ref = O.step_physics(m, p, s, np.zeros(A))
        assert nc[i] == c.n and rows[i] == rows_ref
        for k in range(c0.n):
            if c0.partner[k] >= 1000:
                assert int(pts[i, k, 9]) >= 1000
                assert abs(pts[i, k, 6] - c0.dist[k]) < 2e-6
                assert np.abs(pts[i, k, 0:3] - np.array(c0.pos_a[k][:])).max() < 1e-3
                assert np.abs(pts[i, k, 3:6] - np.array(c0.normal[k][:])).max() < 2e-3
        errs.append(state_error(out[i], O.state_vector(s, A)))
    assert max(errs) < 5e-3 and np.median(errs) < 5e-4, sorted(errs)[-5:]
    assert np.median([state_error(outs[0][0][i], out[i]) for i in range(N)]) > 0.1
