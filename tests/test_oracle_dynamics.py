"""Pins for the CPU oracle's dynamics (no PyBullet available -> independent formulations + invariants).

* M(q) from the world-frame RNEA == M(q) assembled from per-link geometric Jacobians (NumPy, independent)
* ABA (Bullet link-frame formulation) and RNEA round trip: ID(q, qd, FD(q, qd, tau)) == tau
* M^-1 via the ABA unit-impulse response == inverse of M
* free flight without damping conserves linear/angular momentum and (to O(dt)) energy
"""
import numpy as np
import pytest


def _rand_state(O, n, rng, spin=1.0):
    quat = rng.randn(4)
    quat /= np.linalg.norm(quat)
    return O.make_state(n, rng.uniform(-1, 1, 3) + [0, 0, 2.0], quat, spin * rng.randn(3), rng.randn(3),
                        rng.uniform(-0.6, 0.6, n), spin * rng.uniform(-3, 3, n))


def _quat_to_mat(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def jacobian_mass_matrix(t, O, m, s):
    """M = sum_links Jv^T m Jv + Jw^T (R I R^T) Jw with geometric Jacobians built from FK only."""
    n = t["n_dof"]
    pos, rot = O.fk(m, s)  # [1+L,3], local->world
    base_p = pos[0]
    nu = 6 + n
    M = np.zeros((nu, nu))
    masses = [t["base"]["mass"]] + t["mass"]
    inertias = [t["base"]["inertia"]] + t["inertia"]
    for li in range(t["n_links"] + 1):
        Jv = np.zeros((3, nu))
        Jw = np.zeros((3, nu))
        c = pos[li]
        # base angular (world) and linear
        for k in range(3):
            e = np.zeros(3)
            e[k] = 1
            Jw[:, k] = e
            Jv[:, k] = np.cross(e, c - base_p)
            Jv[:, 3 + k] = e
        l = li - 1
        while l >= 0:
            if t["joint_type"][l] == 1:
                R = rot[l + 1]
                a = R @ np.array(t["axis"][l])
                piv = pos[l + 1] - R @ np.array(t["d_vec"][l])
                dcol = 6 + t["dof_of_link"][l]
                Jw[:, dcol] = a
                Jv[:, dcol] = np.cross(a, c - piv)
            l = t["parent"][l]
        Iw = rot[li] @ np.diag(inertias[li]) @ rot[li].T
        M += masses[li] * Jv.T @ Jv + Jw.T @ Iw @ Jw
    return M


def test_fk_tpose(walker_table, oracle_mod):
    O, t = oracle_mod, walker_table
    n = t["n_dof"]
    m = O.model_from_table(t)
    s = O.make_state(n, [0, 0, 1.32], [0, 0, 0, 1], [0] * 3, [0] * 3, np.zeros(n), np.zeros(n))
    pos, _ = O.fk(m, s)
    # foot body origin at z = 0.027 in the T-pose
    for f in t["foot_links"]:
        assert abs(pos[1 + f][2] - 0.027) < 1e-4
    assert abs(t["total_mass"] - 60.0) < 0.1


def test_mass_matrix_vs_jacobians(walker_table, oracle_mod):
    O, t = oracle_mod, walker_table
    m = O.model_from_table(t)
    rng = np.random.RandomState(3)
    for _ in range(4):
        s = _rand_state(O, t["n_dof"], rng)
        M = O.mass_matrix(m, s)
        Mj = jacobian_mass_matrix(t, O, m, s)
        assert np.allclose(M, M.T, atol=1e-12)
        assert np.linalg.eigvalsh(M).min() > 0
        assert np.abs(M - Mj).max() < 1e-10 * np.abs(M).max() + 1e-12


def test_fd_id_roundtrip_and_minv(walker_table, oracle_mod):
    O, t = oracle_mod, walker_table
    m = O.model_from_table(t)
    p = O.default_params()
    n = t["n_dof"]
    rng = np.random.RandomState(4)
    for _ in range(4):
        s = _rand_state(O, n, rng)
        tau = rng.randn(n) * 20
        acc = O.forward_dynamics(m, p, s, tau, with_damping=False)
        back = O.rnea(m, s, acc, p.gravity)
        assert np.abs(back[:6]).max() < 1e-9
        assert np.abs(back[6:] - tau).max() < 1e-9
        M = O.mass_matrix(m, s)
        f = rng.randn(6 + n)
        assert np.abs(M @ O.minv_mult(m, p, s, f) - f).max() < 1e-10


def test_free_flight_conservation(walker_table, oracle_mod):
    """No gravity, no damping, no contacts: momentum/energy drift is pure O(dt) integrator error
    (semi-implicit Euler), i.e. it halves when dt halves and is small in absolute terms."""
    O, t = oracle_mod, walker_table
    n = t["n_dof"]
    errs = []
    for sub in (20, 40):
        m = O.model_from_table(t)
        p = O.default_params()
        p.gravity = 0.0
        p.lin_damping = p.ang_damping = 0.0
        p.has_ground = 0
        p.self_collision = 0  # random joint angles beyond the limits interpenetrate the limbs
        p.dt = 1.0 / 240.0 / sub
        p.substeps = sub
        for d in range(n):  # take joint limits out of play
            m.lower[d], m.upper[d] = 1.0, -1.0
        rng = np.random.RandomState(5)
        s = _rand_state(O, n, rng, spin=0.5)
        e0 = O.energy_momentum(m, s, 0.0)
        for _ in range(10):
            O.step_physics(m, p, s, np.zeros(n))
        e1 = O.energy_momentum(m, s, 0.0)
        errs.append((np.abs(e1["P"] - e0["P"]).max() / np.abs(e0["P"]).max(),
                     np.abs(e1["L"] - e0["L"]).max() / np.abs(e0["L"]).max(),
                     abs(e1["KE"] - e0["KE"]) / e0["KE"]))
    for a, b in zip(*errs):
        assert a < 1e-4
        assert 0.45 < b / a < 0.55  # first-order convergence


def test_free_fall_analytic(walker_table, oracle_mod):
    """Contact-free drop: base COM z follows semi-implicit Euler free fall with Bullet's velocity damping off."""
    O, t = oracle_mod, walker_table
    m = O.model_from_table(t)
    n = t["n_dof"]
    p = O.default_params()
    p.lin_damping = p.ang_damping = 0.0
    s = O.make_state(n, [0, 0, 3.0], [0, 0, 0, 1], [0] * 3, [0] * 3, np.zeros(n), np.zeros(n))
    O.step_physics(m, p, s, np.zeros(n))
    # total COM accelerates at -g; with zero joint torque and symmetric pose the base follows closely
    e = O.energy_momentum(m, s, p.gravity)
    vz = e["P"][2] / t["total_mass"]
    assert abs(vz - (-p.gravity * 4 * p.dt)) < 1e-9


def test_self_contact_is_an_internal_force(walker_table, oracle_mod):
    """Self-collision rows apply equal and opposite impulses to two links of the same multibody:
    without gravity, damping and ground the total linear momentum must not change, and the angular momentum only by
    the friction couple over the contact gap (the penetration depth)."""
    from tests.helpers import oracle_state, self_contact_states

    O, t = oracle_mod, walker_table
    n = t["n_dof"]
    m = O.model_from_table(t)
    rng = np.random.RandomState(11)
    changed = 0
    for row in self_contact_states(O, t, rng, 6):
        p = O.default_params()
        p.gravity = 0.0
        p.lin_damping = p.ang_damping = 0.0
        p.has_ground = 0
        p.substeps = 1
        p.max_coord_vel = 1e9  # the +-100 clamp of btMultiBody is not momentum-conserving
        for d in range(n):
            m.lower[d], m.upper[d] = 1.0, -1.0
        s = oracle_state(O, n, row)
        e0 = O.energy_momentum(m, s, 0.0)
        c, rows = O.step_physics(m, p, s, np.zeros(n))
        assert rows >= 3 and any(c.partner[k] >= 1000 for k in range(c.n))
        p.self_collision = 0
        s2 = oracle_state(O, n, row)
        O.step_physics(m, p, s2, np.zeros(n))
        # the contact changes the motion ...
        v1, v2 = O.state_vector(s, n), O.state_vector(s2, n)
        changed += np.abs(v1 - v2).max() > 1e-4  # (a separating speculative contact applies no impulse)
        # ... but not the momentum.  Evaluated at the start-of-step configuration with the two end-of-step
        # velocities (the position update that follows has its own O(dt |u|^2) drift, different for the two runs).
        mom = []
        for v in (v1, v2):
            mixed = row.copy()
            mixed[7:13] = v[7:13]
            mixed[13 + n:] = v[13 + n:]
            mom.append(O.energy_momentum(m, oracle_state(O, n, mixed), 0.0))
        scale = np.abs(e0["P"]).max() + 1.0
        assert np.abs(mom[0]["P"] - mom[1]["P"]).max() / scale < 1e-9
        # the friction impulses act at the two surface points, |penetration| apart: a small couple remains
        assert np.abs(mom[0]["L"] - mom[1]["L"]).max() / (np.abs(e0["L"]).max() + 1.0) < 5e-2
    assert changed >= 3


def test_resting_contact_pins(walker_table, oracle_mod):
    """Analytic pins of the contact model on the oracle (device twin: tests/test_gpu_analytic.py): a collapsed Walker3D
    at rest on the ground plane receives M g dt of normal impulse per substep from the ground (self-contacts are internal)
    and its loaded contacts rest at distance -slop (ERP 0.9, no split impulse)."""
    O, t = oracle_mod, walker_table
    A, M = 21, t["total_mass"]
    m = O.model_from_table(t)
    p = O.default_params()
    p.substeps = 1
    s = O.make_state(A, [0, 0, 1.32], [0, 0, 0, 1], [0] * 3, [0] * 3, np.array(t["base_joint_angles"]), np.zeros(A))
    for _ in range(1300):
        c, _ = O.step_physics(m, p, s, np.zeros(A))
    assert abs(s.vel[2]) < 1e-2
    ground = [i for i in range(c.n) if c.partner[i] == 0]
    ratio = sum(c.impulse[i] for i in ground) / (M * 9.8 * p.dt)
    assert abs(ratio - 1.0) < 5e-3, ratio
    loaded = [c.dist[i] for i in ground if c.impulse[i] > 0.02 * M * 9.8 * p.dt]
    assert loaded and min(loaded) > -1e-4 and max(loaded) < 2e-5, loaded


def test_persistent_manifold_switch(walker_table, oracle_mod):
    """Oracle-only hypothesis switch orc_params.persistent_manifold (btPersistentManifold semantics against the ground
    plane): a walker dropped onto the plane keeps at most four cached points per link, every one of
    them a candidate the default mode would also report (within the link's breaking threshold), each geom adds at most
    its deeper end per substep -- and with the switch off the persistent buffer's manifold region stays untouched."""
    import ctypes as C

    O, t = oracle_mod, walker_table
    A = t["n_dof"]
    m = O.model_from_table(t)
    rng = np.random.RandomState(0)

    def drop(pm, steps=160):
        p = O.default_params()
        p.persistent_manifold = pm
        p.self_collision = 0
        q0 = np.array(t["base_joint_angles"]) + rng.uniform(-0.2, 0.2, A)
        s = O.make_state(A, [0, 0, 1.0], [0.3, 0.1, 0, 0.95], [0] * 3, [0] * 3, q0, np.zeros(A))
        warm = (C.c_double * O.WARMSZ)()
        seen_fewer = False
        for _ in range(steps):
            c, _ = O.step_physics(m, p, s, np.zeros(A), warm=warm)
            ids = [int(c.point_id[i]) for i in range(c.n)]
            links = [int(c.link[i]) for i in range(c.n)]
            if pm:
                assert len(set(ids)) == len(ids)
                assert max([links.count(x) for x in set(links)] or [0]) <= 4
                for i in range(c.n):  # a cached point is still within its link's breaking threshold
                    assert c.dist[i] <= m.link_thresh[links[i] + 1] + 1e-12
            man = np.array(warm[O.MAXW:])
            if not pm:
                assert not man.any()
            else:
                seen_fewer = seen_fewer or (0 < c.n)
        return np.array(warm[O.MAXW:]), O.state_vector(s, A), seen_fewer

    rng = np.random.RandomState(0)
    man_off, s_off, _ = drop(0)
    rng = np.random.RandomState(0)
    man_on, s_on, touched = drop(1)
    assert touched and man_on.any() and not man_off.any()
    assert s_on[2] < 0.3 and s_off[2] < 0.3  # both heaps ended on the ground
