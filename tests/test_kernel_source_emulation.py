"""CPU checks of the CUDA kernel SOURCE: mb_core.cuh / mb_env.cuh compiled by g++ as a 32-lane loop
(tests/emu) and diffed against the float64 oracle.  The GPU tests (-m gpu) repeat these through the C ABI on the
real device; this file exists so that kernel-logic regressions are caught on a box without a GPU."""
import numpy as np
import pytest

from tests.emu import emu as E
from tests.helpers import contact_states, oracle_state, random_states, state_error


def test_mass_matrix_and_bias(walker_table, oracle_mod):
    O, t = oracle_mod, walker_table
    A = t["n_dof"]
    m = O.model_from_table(t)
    p = O.default_params()
    ep = E.default_phys()
    rng = np.random.RandomState(0)
    for row in random_states(t, rng, 6):
        s = oracle_state(O, A, row)
        M = O.mass_matrix(m, s)
        Me, be = E.mass_matrix(ep, row.astype(np.float32), 6 + A)
        assert np.abs(Me - M).max() / np.abs(M).max() < 1e-6
        acc = O.forward_dynamics(m, p, s, np.zeros(A), with_damping=True)
        bias = -M @ acc
        assert np.abs(be - bias).max() / np.abs(bias).max() < 1e-5


def test_contact_free_step(walker_table, oracle_mod):
    """north_star: contact-free single-step state within 1e-4."""
    O, t = oracle_mod, walker_table
    A = t["n_dof"]
    m = O.model_from_table(t)
    p = O.default_params()
    ep = E.default_phys()
    rng = np.random.RandomState(1)
    gain = np.array(t["gain"])
    for row in random_states(t, rng, 8, spin=0.5, margin=0.35):
        tau = gain * rng.uniform(-1, 1, A)
        s = oracle_state(O, A, row)
        _, rows = O.step_physics(m, p, s, tau)
        out, erows, _ = E.step_physics(ep, row.astype(np.float32), tau)
        assert rows == erows == 0
        assert state_error(out, O.state_vector(s, A)) < 1e-4


def test_contact_step(walker_table, oracle_mod):
    """north_star: per-step contact rollouts within a stated tolerance over 1 frame: 2e-3 (state-scaled)."""
    O, t = oracle_mod, walker_table
    A = t["n_dof"]
    m = O.model_from_table(t)
    p = O.default_params()
    ep = E.default_phys()
    rng = np.random.RandomState(2)
    gain = np.array(t["gain"])
    worst = 0.0
    for row in contact_states(O, t, rng, 12):
        row = row.astype(np.float32).astype(np.float64)
        tau = 0.3 * gain * rng.uniform(-1, 1, A)
        s = oracle_state(O, A, row)
        c, rows = O.step_physics(m, p, s, tau)
        out, erows, enc = E.step_physics(ep, row.astype(np.float32), tau)
        assert enc == c.n
        assert abs(rows - erows) <= 2  # a limit row can flip at an exact boundary
        worst = max(worst, state_error(out, O.state_vector(s, A)))
    assert worst < 2e-3, worst


def _mt_row(O, seed):
    st = np.random.RandomState(O.gym_seed_words(seed)).get_state()
    return np.concatenate([st[1], [st[2]]]).astype(np.uint32)


def test_reset_bit_exact(walker_table, oracle_mod):
    """north_star: bit-exact reset-state generation from the same seed (after rounding to the f32 state)."""
    O, t = oracle_mod, walker_table
    for seed in range(3):
        env = O.Walker3DCustomOracle(t, seed=seed)
        emu = E.EmuW3D(_mt_row(O, seed))
        for _ in range(3):
            o_ref = env.reset()
            o_emu = emu.reset()
            q_ref = np.array(env.e.s.q[:21]).astype(np.float32)
            assert np.array_equal(emu.state[13:34], q_ref)
            assert np.array_equal(emu.rec[:3], np.array(env.e.walk_target[:], dtype=np.float32))
            assert emu.rec[5] == env.e.stop_frames
            assert np.abs(o_ref - o_emu).max() < 1e-6


def test_env_step_teacher_forced(walker_table, oracle_mod):
    """obs / reward / done of Walker3DCustomEnv.step from identical states and bookkeeping (oracle state and
    record injected every step).  The restated Bullet step is discontinuous (limit rows appear at q<=lo, the
    split-impulse threshold at pen=-0.04 switches the positional term, contacts appear at the breaking threshold),
    so f32 and f64 occasionally land on different sides: >= 97% of env-steps must agree."""
    from tests.helpers import force_oracle_state, oracle_record

    O, t = oracle_mod, walker_table
    N = 6
    oracles = [O.Walker3DCustomOracle(t, seed=5 + i) for i in range(N)]
    emus = [E.EmuW3D(_mt_row(O, 5 + i)) for i in range(N)]
    for o, e in zip(oracles, emus):
        o.reset()
        e.reset()
    arng = np.random.RandomState(7)
    bad, total, errs = 0, 0, []
    for step in range(40):
        for o, e in zip(oracles, emus):
            a = arng.uniform(-1.2, 1.2, 21)
            sv = o.state_vector().astype(np.float32)
            e.state[:55] = sv
            oracle_record(o, e.rec)
            force_oracle_state(o, sv.astype(np.float64))
            o1, r1, d1, _ = o.step(a)
            o2, r2, d2, tr2, fin = e.step(a)
            ocmp = fin if d2 else o2
            err = float(np.abs(o1 - ocmp).max())
            ok = d1 == d2 and err < 5e-3 and abs(r1 - r2) < 5e-2 + 1e-3 * abs(r1)
            total += 1
            bad += 0 if ok else 1
            errs.append(err)
            if d1:
                o.reset()
    assert bad <= 0.03 * total, (bad, total)
    assert np.median(errs) < 2e-4
