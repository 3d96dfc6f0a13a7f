#!/usr/bin/env python
"""PyBullet golden-vector generator (OQ1-OQ12, G1-G6) for the four north-star envs.

Runs on any host that has `pybullet` and `gym<=0.21` with the reference importable
(`pip install pybullet gym==0.21 && pip install -e /path/to/mocca_envs`):

    python tools/gen_pybullet_golden.py [--envs Walker3DCustomEnv,Walker3DStepperEnv,Monkey3DCustomEnv,CassieEnv] [--out DIR]

writes tests/golden/pybullet_<env>_<apiversion>.npz, consumed by tests/test_pybullet_golden.py (oracle legs on the CPU,
device legs under -m gpu).  PyBullet is not installable in the build container (no wheel, no network), so no such file
is committed and every statement about Bullet's arithmetic in this repository reads "vs restatement".

`--standin` runs the SAME script against the oracle-backed stand-in Bullet client of tools/gen_reference_golden.py and
writes pybullet_<env>_standin.npz to a scratch directory: that exercises every code path of this generator and of the
consuming tests here (tests/test_pybullet_golden.py::test_generator_and_consumers_run_on_the_standin); a stand-in file
pins nothing and is never committed.

Per env the file holds (every call wrapped: an API a PyBullet version lacks leaves an "error:<msg>" string instead):
  OQ1  getJointInfo of every joint            OQ2/OQ3  getDynamicsInfo of every link (base = -1)
  OQ4  getCollisionShapeData of every link    OQ12     getAPIVersion
  G1   calculateMassMatrix / calculateInverseDynamics at 16 random states
  G2   contact-free single stepSimulation from 16 random airborne states under random torques; OQ6: the same states
       stepped as 4 x numSubSteps=1 with the torque re-applied
  OQ5  free fall + spin decay of the robot (zero torque, 60 steps): base velocity trace
  G3   in-contact traces: from the reset pose, zero action, 30 env steps; per step the state, len(getContactPoints)
       and each point's (linkA, bodyB, linkB, posA, normal, distance, normalForce)  [OQ8, OQ9, OQ10, OQ11]
  G4   1000 random-action env steps after env.seed(0), actions RandomState(0).uniform(-1, 1, A) (the stream of
       tools/gen_config1_trace.py): observation, reward, done and the physics state before every step
  G5   episode length / return of 200 episodes each under the zero, random and scripted-PD policies
  G6   reset states and first observations for seeds 0..15
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ENVS = ["Walker3DCustomEnv", "Walker3DStepperEnv", "Monkey3DCustomEnv", "CassieEnv"]


def _try(f, *a, **kw):
    try:
        return f(*a, **kw)
    except Exception as e:  # noqa: BLE001 -- the dump must survive an API that a PyBullet version lacks
        return "error:%s: %s" % (type(e).__name__, e)


def make_env(name, standin):
    if standin:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import gen_reference_golden as R

        if "pybullet" not in sys.modules:
            R.install_gym()
            R.install_pybullet()
            import gym.utils.seeding as seeding

            real = seeding.np_random  # EnvBase.__init__ seeds from os.urandom: fixed, as in gen_reference_golden.main
            seeding.np_random = lambda seed=None: real(R.CONSTRUCTION_SEED if seed is None else seed)
        if R.REF not in sys.path:
            sys.path.insert(0, R.REF)
        if name == "CassieEnv":
            return R.make_cassie_env()
        import mocca_envs.env_locomotion as EL

        return getattr(EL, name)()
    import gym

    return gym.make("mocca_envs:%s-v0" % name).unwrapped


def movable_joints(u):
    """Joint indices of the robot's movable joints in Bullet's own order (the order of calculateMassMatrix's
    objPositions and of the model tables / oracle / kernel state vectors) -- not the env's ordered_joints order."""
    p, rid = u._p, u.robot.id if hasattr(u.robot, "id") else u.robot.object_id[0]
    return [j for j in range(p.getNumJoints(rid)) if p.getJointInfo(rid, j)[2] != 4]  # 4 = JOINT_FIXED


def robot_state(u):
    """[pos3 quat4 omega3 vel3 q[A] qd[A]] of the robot, joints in movable_joints() order."""
    p, rid = u._p, u.robot.id if hasattr(u.robot, "id") else u.robot.object_id[0]
    ids = movable_joints(u)
    bp, bq = p.getBasePositionAndOrientation(rid)
    bv, bw = p.getBaseVelocity(rid)
    js = [p.getJointState(rid, j) for j in ids]
    return np.concatenate([bp, bq, bw, bv, [s[0] for s in js], [s[1] for s in js]])


def set_robot_state(u, st):
    p, rid = u._p, u.robot.id if hasattr(u.robot, "id") else u.robot.object_id[0]
    ids = movable_joints(u)
    A = len(ids)
    p.resetBasePositionAndOrientation(rid, list(map(float, st[0:3])), list(map(float, st[3:7])))
    p.resetBaseVelocity(rid, list(map(float, st[10:13])), list(map(float, st[7:10])))
    for k, j in enumerate(ids):
        p.resetJointState(rid, j, float(st[13 + k]), float(st[13 + A + k]))


def contact_rows(u):
    p, rid = u._p, u.robot.id if hasattr(u.robot, "id") else u.robot.object_id[0]
    rows = []
    for c in p.getContactPoints(bodyA=rid):
        if len(c) >= 10:
            rows.append([c[3], c[2], c[4], *c[5], *c[7], c[8], c[9]])
        else:  # a client that reports ids only
            rows.append([c[3], c[2], c[4]] + [np.nan] * 8)
    return np.array(rows, dtype=np.float64).reshape(-1, 11)


def dump_env(name, standin, quick=False):
    """quick: 120 G4 steps and 5 G5 episodes per policy (the smoke test of the generator itself)."""
    u = make_env(name, standin)
    p = u._p
    rid = u.robot.id if hasattr(u.robot, "id") else u.robot.object_id[0]
    out = {"env": name, "api_version": str(_try(p.getAPIVersion)), "standin": int(bool(standin))}
    nj = p.getNumJoints(rid)
    out["oq1_joint_info"] = np.array([repr(_try(p.getJointInfo, rid, j)) for j in range(nj)])
    out["oq2_dynamics_info"] = np.array([repr(_try(p.getDynamicsInfo, rid, l)) for l in range(-1, nj)])
    out["oq4_collision_shapes"] = np.array([repr(_try(p.getCollisionShapeData, rid, l)) for l in range(-1, nj)])
    ids = movable_joints(u)
    A = len(ids)
    out["movable_joint_ids"] = np.array(ids)
    out["ordered_joint_ids"] = np.array([j.jointIndex for j in u.robot.ordered_joints])
    info = [p.getJointInfo(rid, j) for j in ids]
    lo = np.array([i[8] for i in info], dtype=np.float64)
    hi = np.array([i[9] for i in info], dtype=np.float64)
    lo, hi = np.where(lo <= hi, lo, -1.0), np.where(lo <= hi, hi, 1.0)  # unlimited joints: sample in [-1, 1]
    og = np.asarray(getattr(u.robot, "ordered_joint_base_gains", getattr(u.robot, "torque_limits", [])), dtype=np.float64)
    gains = np.ones(A)
    if og.shape == (len(u.robot.ordered_joints),):
        for gq, j in zip(og, u.robot.ordered_joints):
            gains[ids.index(j.jointIndex)] = gq
    u.seed(0)
    u.reset()
    rng = np.random.RandomState(0)
    # ---- G1 / G2 / OQ6
    states, Ms, ids_, taus, after, after4 = [], [], [], [], [], []
    for k in range(16):
        q = lo + (hi - lo) * rng.uniform(0.35, 0.65, A)
        qd = rng.uniform(-1, 1, A)
        quat = rng.randn(4)
        quat /= np.linalg.norm(quat)
        st = np.concatenate([[0.0, 0.0, 3.0], quat, 0.5 * rng.randn(3), rng.randn(3), q, qd])
        acc = rng.randn(6 + A)
        tau = rng.uniform(-1, 1, A) * gains
        set_robot_state(u, st)
        states.append(st)
        Ms.append(_try(p.calculateMassMatrix, rid, [float(x) for x in q]))
        idr = _try(p.calculateInverseDynamics, rid, [float(x) for x in q], [float(x) for x in qd], [float(x) for x in acc[6:]])
        ids_.append(np.full(A, np.nan) if isinstance(idr, str) else np.array(idr, dtype=np.float64))
        taus.append(np.concatenate([tau, acc]))
        p.setJointMotorControlArray(rid, ids, p.TORQUE_CONTROL, forces=[float(x) for x in tau])
        p.stepSimulation()
        after.append(robot_state(u))
        # OQ6: the same state as numSubSteps x (one substep, torque re-applied)
        set_robot_state(u, st)
        r4 = _try(_four_single_substeps, u, rid, ids, tau)
        after4.append(r4 if not isinstance(r4, str) else np.full(13 + 2 * A, np.nan))
    out["g1_states"] = np.array(states)
    if any(isinstance(m_, str) for m_ in Ms):
        out["g1_mass_matrix"] = np.array([repr(m_) for m_ in Ms])
    else:
        out["g1_mass_matrix"] = np.array(Ms, dtype=np.float64)
    out["g1_inverse_dynamics"] = np.array(ids_)
    out["g2_tau_acc"] = np.array(taus)
    out["g2_after_step"] = np.array(after)
    out["oq6_after_4x1_substeps"] = np.array(after4)
    # ---- OQ5: free fall and spin decay
    st = np.concatenate([[0.0, 0.0, 30.0], [0, 0, 0, 1], [0.0, 0.0, 3.0], [0.0, 0.0, 0.0], 0.5 * (lo + hi), np.zeros(A)])
    set_robot_state(u, st)
    fall = []
    for k in range(60):
        p.setJointMotorControlArray(rid, ids, p.TORQUE_CONTROL, forces=[0.0] * A)
        p.stepSimulation()
        fall.append(robot_state(u)[:13])
    out["oq5_free_fall"] = np.array(fall)
    # ---- G3: in-contact traces with contact lists
    u.seed(0)
    u.reset()
    g3s, g3n, g3c = [], [], []
    for k in range(30):
        g3s.append(robot_state(u))
        o, r, d, _ = u.step(np.zeros(u.action_space.shape[0]))
        rows = contact_rows(u)
        g3n.append(len(rows))
        g3c.append(np.concatenate([np.full((len(rows), 1), k), rows], axis=1))
        if d:
            u.reset()
    out["g3_states"] = np.array(g3s)
    out["g3_contact_counts"] = np.array(g3n)
    out["g3_contacts"] = np.concatenate(g3c) if g3c else np.zeros((0, 12))
    # ---- G4: the config-1 stream through the env (TimeLimit(1000) applied by hand: `unwrapped` has none)
    u.seed(0)
    obs, states, rews, dones, acts = [u.reset()], [robot_state(u)], [], [], []
    arng = np.random.RandomState(0)
    AD = u.action_space.shape[0]
    elapsed = 0
    for k in range(120 if quick else 1000):
        a = arng.uniform(-1, 1, AD)
        o, r, d, _ = u.step(a.copy())
        elapsed += 1
        d = bool(d) or elapsed >= 1000
        acts.append(a); obs.append(o); rews.append(r); dones.append(d); states.append(robot_state(u))
        if d:
            obs.append(u.reset())
            states.append(robot_state(u))
            elapsed = 0
    out.update(seed=0, construction_seed=12345 if standin else -1, action_seed=0, eval_mode=0, actions=np.array(acts), obs=np.array(obs, dtype=np.float64),
               states=np.array(states), rewards=np.array(rews), dones=np.array(dones))
    # ---- G5: episode statistics
    olo = np.array([j.lowerLimit for j in u.robot.ordered_joints], dtype=np.float64)
    ohi = np.array([j.upperLimit for j in u.robot.ordered_joints], dtype=np.float64)
    ref = 2 * (np.array(getattr(u.robot, "base_joint_angles", 0.5 * (olo + ohi)), dtype=np.float64) - olo) / np.where(
        ohi > olo, ohi - olo, 1.0) - 1  # the bench's scripted PD: running-start pose in normalised joint units
    for pol in ("zero", "random", "pd"):
        prng = np.random.RandomState(7)
        lens, rets = [], []
        for ep in range(5 if quick else 200):
            o = u.reset()
            L, Rr = 0, 0.0
            while True:
                if pol == "zero":
                    a = np.zeros(AD)
                elif pol == "random":
                    a = prng.uniform(-1, 1, AD)
                else:
                    a = np.clip(1.0 * (ref[:AD] - o[6:6 + AD]) - 0.1 * (o[6 + A:6 + A + AD] * 10.0), -1, 1) if AD == A else np.zeros(AD)
                o, r, d, _ = u.step(a)
                L += 1
                Rr += r
                if d or L >= 1000:
                    break
            lens.append(L)
            rets.append(Rr)
        out["g5_%s_lengths" % pol] = np.array(lens)
        out["g5_%s_returns" % pol] = np.array(rets)
    # ---- G6: reset states
    rs, ro = [], []
    for seed in range(16):
        u.seed(seed)
        ro.append(u.reset())
        rs.append(robot_state(u))
    out["g6_reset_states"] = np.array(rs)
    out["g6_reset_obs"] = np.array(ro, dtype=np.float64)
    return out


def _four_single_substeps(u, rid, ids, tau):
    p = u._p
    dt = u.control_step / u.llc_frame_skip
    n = u.sim_frame_skip
    p.setPhysicsEngineParameter(fixedTimeStep=dt / n, numSolverIterations=5, numSubSteps=1)
    try:
        for _ in range(n):
            p.setJointMotorControlArray(rid, ids, p.TORQUE_CONTROL, forces=[float(x) for x in tau])
            p.stepSimulation()
        return robot_state(u)
    finally:
        p.setPhysicsEngineParameter(fixedTimeStep=dt, numSolverIterations=5, numSubSteps=n)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", default=",".join(ENVS))
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden"))
    ap.add_argument("--standin", action="store_true")
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    if not args.standin:
        try:
            import gym  # noqa: F401
            import pybullet  # noqa: F401
        except ImportError as e:
            sys.exit("gen_pybullet_golden: %s -- run this on a host with pybullet and gym<=0.21, or pass --standin" % e)
    os.makedirs(args.out, exist_ok=True)
    for name in args.envs.split(","):
        d = dump_env(name, args.standin, args.quick)
        tag = "standin" if args.standin else str(d["api_version"])
        path = os.path.join(args.out, "pybullet_%s_%s.npz" % (name, tag))
        np.savez_compressed(path, **d)
        print("wrote", path, "G4 steps", len(d["actions"]), "episodes", int(d["dones"].sum()))


if __name__ == "__main__":
    main()
