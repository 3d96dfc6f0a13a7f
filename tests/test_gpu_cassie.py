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


def test_cassie_env_step_teacher_forced(cassie_table, oracle_mod, torch_mod):
    """CassieEnv.step (50 PD substeps, toe contacts, loop closures) from identical states and bookkeeping for 8
    action streams: the oracle's state (rounded to f32) and low-pass joint velocities are injected before every
    step.  >= 95 % of env-steps within 1e-2 (obs; raw joint speeds in rad/s dominate) / 2e-3 (reward), same done."""
    torch, O, t = torch_mod, oracle_mod, cassie_table
    N, A = 8, t["n_dof"]
    env = _env(N, return_final_obs=True)
    obs0 = env.reset().cpu().numpy()
    oracles = [O.CassieOracle(t) for _ in range(N)]
    for i, o in enumerate(oracles):
        assert np.abs(o.reset() - obs0[i]).max() < 1e-5
    rng = np.random.RandomState(0)
    bad, total, errs = 0, 0, []
    for step in range(16):
        st = np.zeros((N, 13 + 2 * A), dtype=np.float32)
        rec = env.get_record().cpu().numpy()
        ri = rec.view(np.int32)
        for i, o in enumerate(oracles):
            sv = o.state_vector().astype(np.float32)
            st[i] = sv
            b = o.e.base
            for k in range(3):
                b.s.pos[k] = float(sv[k]); b.s.omega[k] = float(sv[7 + k]); b.s.vel[k] = float(sv[10 + k])
            for k in range(4):
                b.s.quat[k] = float(sv[3 + k])
            for k in range(A):
                b.s.q[k] = float(sv[13 + k]); b.s.qd[k] = float(sv[13 + A + k])
            ri[i, 8] = b.elapsed
            rec[i, env.EC_POTENTIAL] = o.e.potential
            rec[i, 23], rec[i, 24] = sv[0], sv[1]  # EC_PREVX / EC_PREVY: position at the last calc_potential
            rec[i, env.EC_JVEL:env.EC_JVEL + 14] = np.array(o.e.jvel[:14], dtype=np.float32)
        env.set_state(torch.tensor(st))
        env.set_record(torch.tensor(rec))
        acts = 0.1 * rng.uniform(-1, 1, (N, 10))
        obs, rew, done, info = env.step(torch.tensor(acts, dtype=torch.float32))
        obs, rew, done = obs.cpu().numpy(), rew.cpu().numpy(), done.cpu().numpy()
        fin = info["terminal_observation"].cpu().numpy()
        for i, o in enumerate(oracles):
            # the oracle's rad_angles / speeds caches must correspond to the injected f32 state
            o1, r1, d1, _ = o.step(acts[i])
            err = float(np.abs(o1 - (fin[i] if done[i] else obs[i])).max())
            ok = bool(d1) == bool(done[i]) and err < 1e-2 and abs(r1 - rew[i]) < 2e-3
            total += 1
            bad += 0 if ok else 1
            errs.append(err)
            if d1:
                o.reset()
    assert bad <= 0.05 * total, (bad, total, sorted(errs)[-5:])
    assert np.median(errs) < 3e-3, np.median(errs)
    env.close()


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
