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


def test_monkey_env_step_teacher_forced(monkey_table, oracle_mod, torch_mod):
    """Monkey3DCustomEnv.step from identical states and bookkeeping: hand / palm contacts with the bars, scripted
    finger joints, swing progress, free-fall termination.  >= 95 % of env-steps within 5e-3 (obs) / 5e-2 (reward)."""
    torch, O, t = torch_mod, oracle_mod, monkey_table
    N, A = 12, 23
    env = _env(N, seed=300, return_final_obs=True)
    oracles = [O.Monkey3DOracle(t, seed=300 + i) for i in range(N)]
    env.reset()
    for o in oracles:
        o.reset()
    arng = np.random.RandomState(5)
    bad, total, errs, contacts = 0, 0, [], 0
    for step in range(50):
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
            rec[i, 0:3] = np.array(b.walk_target[:], dtype=np.float32)
            rec[i, 9], rec[i, 10] = b.feet_contact[0], b.feet_contact[1]
            ri[i, 8] = b.elapsed
            ri[i, env.EM_NEXT], ri[i, env.EM_FREEFALL], ri[i, env.EM_TIMESTEP] = (
                o.e.next_step_index, o.e.free_fall_count, o.e.timestep)
            ri[i, env.EM_SWING], ri[i, env.EM_PIVOT] = o.e.swing_leg, o.e.pivot_leg
            rec[i, 27] = o.e.swing_potential
            rec[i, env.EM_TERRAIN:env.EM_TERRAIN + 128] = np.array([list(r) for r in o.e.terrain], dtype=np.float32).ravel()
            for k in range(4):
                bar = o.e.bars[k]
                rec[i, env.EM_BAR + 8 * k:env.EM_BAR + 8 * k + 8] = np.array(
                    list(bar.center) + list(bar.axis) + [bar.halflen, bar.radius], dtype=np.float32)
        env.set_state(torch.tensor(st))
        env.set_record(torch.tensor(rec))
        acts = 0.5 * arng.uniform(-1, 1, (N, A))
        obs, rew, done, info = env.step(torch.tensor(acts, dtype=torch.float32))
        obs, rew, done = obs.cpu().numpy(), rew.cpu().numpy(), done.cpu().numpy()
        fin = info["terminal_observation"].cpu().numpy()
        for i, o in enumerate(oracles):
            o1, r1, d1, _ = o.step(acts[i])
            contacts += o.e.base.last_contacts.n
            ocmp = fin[i] if done[i] else obs[i]
            err = float(np.abs(o1 - ocmp).max())
            ok = bool(d1) == bool(done[i]) and err < 5e-3 and abs(r1 - rew[i]) < 5e-2 + 1e-3 * abs(r1)
            total += 1
            bad += 0 if ok else 1
            errs.append(err)
            if d1:
                o.reset()
    assert contacts > 200
    assert bad <= 0.05 * total, (bad, total)
    assert np.median(errs) < 3e-4, np.median(errs)
    env.close()


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
