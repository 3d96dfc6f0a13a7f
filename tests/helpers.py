"""Shared helpers for parity tests (oracle vs kernel source)."""
import numpy as np


def random_states(table, rng, n, airborne=True, spin=1.0, margin=0.15):
    """[n, 13+2A] float64 rows: pos3 quat4 omega3 vel3 q qd, joints strictly inside their limits."""
    A = table["n_dof"]
    lo, hi = np.array(table["lower"]), np.array(table["upper"])
    out = np.zeros((n, 13 + 2 * A))
    for i in range(n):
        quat = rng.randn(4)
        quat /= np.linalg.norm(quat)
        out[i, 0:3] = rng.uniform(-1, 1, 3) + [0, 0, 3.0 if airborne else 1.3]
        out[i, 3:7] = quat
        out[i, 7:10] = spin * rng.randn(3)
        out[i, 10:13] = rng.randn(3)
        out[i, 13:13 + A] = lo + (hi - lo) * rng.uniform(margin, 1 - margin, A)
        out[i, 13 + A:] = spin * rng.uniform(-3, 3, A)
    return out


def oracle_state(O, A, row):
    return O.make_state(A, row[0:3], row[3:7], row[7:10], row[10:13], row[13:13 + A], row[13 + A:13 + 2 * A])


def contact_states(O, table, rng, n, steps=(6, 30)):
    """States sampled from oracle rollouts that are in ground contact (random bounded torques)."""
    A = table["n_dof"]
    m = O.model_from_table(table)
    p = O.default_params()
    gain = np.array(table["gain"])
    out = []
    while len(out) < n:
        q0 = np.array(table["base_joint_angles"]) + rng.uniform(-0.1, 0.1, A)
        s = O.make_state(A, [0, 0, 1.32], [0, 0, 0, 1], [0] * 3, [0] * 3, q0, np.zeros(A))
        k = rng.randint(*steps)
        for _ in range(k):
            c, _ = O.step_physics(m, p, s, 0.3 * gain * rng.uniform(-1, 1, A))
        if c.n > 0 and s.pos[2] > 0.5:
            out.append(O.state_vector(s, A))
    return np.array(out)


def state_error(out, ref):
    """max abs error scaled per component by max(1, |ref|)."""
    return float(np.max(np.abs(out - ref) / np.maximum(1.0, np.abs(ref))))


def oracle_record(o, rec_row):
    """Teacher-force the per-env bookkeeping record (ER_* in csrc/mb_env.cuh) from an oracle env."""
    e = o.e
    rec_row[0:3] = np.array(e.walk_target[:], dtype=np.float32)
    rec_row[3] = e.dist
    rec_row[4] = e.angle
    rec_row[5] = e.stop_frames
    rec_row.view(np.int32)[6] = e.close_count
    rec_row[7] = e.linear_potential
    rec_row.view(np.int32)[8] = e.elapsed
    rec_row[9] = e.feet_contact[0]
    rec_row[10] = e.feet_contact[1]
    rec_row[17] = e.body_xyz[0]
    return rec_row


def force_oracle_state(o, sv):
    """Overwrite the oracle env's physics state with the (f32-rounded) row the kernel sees."""
    A = o.A
    for k in range(3):
        o.e.s.pos[k] = sv[k]
        o.e.s.omega[k] = sv[7 + k]
        o.e.s.vel[k] = sv[10 + k]
    for k in range(4):
        o.e.s.quat[k] = sv[3 + k]
    for k in range(A):
        o.e.s.q[k] = sv[13 + k]
        o.e.s.qd[k] = sv[13 + A + k]


def self_contact_states(O, table, rng, n, max_tries=4000):
    """Airborne f32-rounded states (random held torques from the base pose) whose next physics step sees at least
    one self-contact (robots.py:259-264) according to the oracle."""
    A = table["n_dof"]
    m = O.model_from_table(table)
    p = O.default_params()
    gain = np.array(table["gain"])
    out = []
    for _ in range(max_tries):
        q0 = np.array(table["base_joint_angles"]) + rng.uniform(-0.1, 0.1, A)
        s = O.make_state(A, [0, 0, 3.0], [0, 0, 0, 1], [0] * 3, [0] * 3, q0, np.zeros(A))
        a = rng.uniform(-1, 1, A)
        for _k in range(rng.randint(3, 25)):
            O.step_physics(m, p, s, gain * a)
        row = O.state_vector(s, A).astype(np.float32).astype(np.float64)
        c = O.collide(m, p, oracle_state(O, A, row))
        if any(c.partner[k] >= 1000 for k in range(c.n)):
            out.append(row)
            if len(out) == n:
                break
    assert len(out) == n
    return np.array(out)


def cassie_hull_contact_states(O, table, rng, n, deepest=-1.0e-3, max_tries=40000):
    """Airborne Cassie poses (zero velocity) inside the joint limits whose left-leg / right-leg mesh hulls are in
    shallow contact according to the oracle (every hull contact no deeper than `deepest`: the GJK path; deeper overlaps
    take the centre-line fallback)."""
    A = table["n_dof"]
    m = O.model_from_table(table)
    p = O.cassie_params()
    lo, hi = np.array(table["lower"]), np.array(table["upper"])
    base = np.array(table["base_joint_angles"])
    out = []
    for _ in range(max_tries):
        q = np.clip(base + rng.uniform(-1, 1, A) * 0.5 * (hi - lo) * rng.uniform(0.2, 1.0), lo + 1e-3, hi - 1e-3)
        s = O.make_state(A, [0, 0, 3.0], [0, 0, 0, 1], [0] * 3, [0] * 3, q, np.zeros(A))
        c = O.collide(m, p, s)
        d = [c.dist[k] for k in range(c.n) if c.partner[k] >= 1000]
        if d and min(d) > deepest:
            out.append(np.concatenate([[0, 0, 3.0], [0, 0, 0, 1], np.zeros(6), q, np.zeros(A)]))
            if len(out) == n:
                return np.array(out, dtype=np.float32)
    raise AssertionError("not enough hull-contact states")
