"""ctypes binding of the CPU oracle (libmocca_oracle.so).  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED against PyBullet (see mocca_oracle.h).  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module; the product package
``mocca_envs_b200`` never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmocca_oracle.so")

MAXL, MAXD, MAXG, MAXP = 40, 40, 192, 96
MAXW = 2 * MAXG
WARMSZ = MAXW + 16 * (MAXL + 1)  # + the ground-plane manifolds of the persistent_manifold switch
MAXU = 6 + MAXD
MAXSP = 256
MAXHULL, HULLV, MAXHPAIR = 24, 32, 64

d = C.c_double
i32 = C.c_int


class Model(C.Structure):
    _fields_ = [
        ("n_links", i32), ("n_dof", i32), ("n_geoms", i32), ("n_feet", i32),
        ("parent", i32 * MAXL), ("joint_type", i32 * MAXL), ("dof_of_link", i32 * MAXL), ("link_of_dof", i32 * MAXD),
        ("axis", (d * 3) * MAXL), ("rot_p2t", (d * 4) * MAXL), ("e_vec", (d * 3) * MAXL), ("d_vec", (d * 3) * MAXL),
        ("mass", d * MAXL), ("inertia", (d * 3) * MAXL),
        ("base_mass", d), ("base_inertia", d * 3),
        ("lower", d * MAXD), ("upper", d * MAXD), ("damping", d * MAXD), ("armature", d * MAXD), ("gain", d * MAXD),
        ("link_thresh", d * (MAXL + 1)), ("link_group", i32 * (MAXL + 1)), ("link_mask", i32 * (MAXL + 1)),
        ("geom_link", i32 * MAXG), ("geom_type", i32 * MAXG),
        ("geom_p0", (d * 3) * MAXG), ("geom_p1", (d * 3) * MAXG), ("geom_quat", (d * 4) * MAXG),
        ("geom_size", (d * 3) * MAXG), ("geom_friction", d * MAXG),
        ("foot_link", i32 * 4),
        ("base_joint_angles", d * MAXD), ("base_position", d * 3),
        ("base_orientation", d * 4), ("termination_height", d), ("planar_env", d), ("stepper_init_position", d * 3),
        ("n_right", i32), ("right_idx", i32 * MAXD), ("left_idx", i32 * MAXD),
        ("n_neg", i32), ("neg_idx", i32 * 8),
        ("palm_link", i32 * 2),
        ("n_p2p", i32), ("p2p_link_a", i32 * 2), ("p2p_link_b", i32 * 2),
        ("p2p_pivot_a", (d * 3) * 2), ("p2p_pivot_b", (d * 3) * 2), ("p2p_max_impulse", d * 2),
        ("n_self", i32), ("self_a", i32 * MAXSP), ("self_b", i32 * MAXSP),
        ("n_hulls", i32), ("hull_link", i32 * MAXHULL), ("hull_verts", ((d * 3) * HULLV) * MAXHULL),
        ("hull_center", (d * 3) * MAXHULL), ("hull_radius", d * MAXHULL), ("hull_friction", d * MAXHULL),
        ("n_hpairs", i32), ("hpair_a", i32 * MAXHPAIR), ("hpair_b", i32 * MAXHPAIR), ("hull_margin", d),
        ("n_ordered", i32), ("ordered_dof", i32 * MAXD),
        ("n_pd", i32), ("pd_ordered_index", i32 * 16), ("pd_kp", d * 16), ("pd_kd", d * 16),
    ]


class Params(C.Structure):
    _fields_ = [
        ("gravity", d), ("dt", d), ("substeps", i32), ("iterations", i32),
        ("erp_contact", d), ("erp_joint", d), ("linear_slop", d), ("lin_damping", d), ("ang_damping", d),
        ("max_coord_vel", d), ("warmstart", d), ("limit_max_impulse", d), ("split_threshold", d),
        ("residual_threshold", d), ("limit_rows_always", i32), ("gyro", i32), ("has_ground", i32),
        ("ground_friction", d), ("self_collision", i32), ("persistent_manifold", i32),
    ]


class State(C.Structure):
    _fields_ = [("pos", d * 3), ("quat", d * 4), ("omega", d * 3), ("vel", d * 3), ("q", d * MAXD), ("qd", d * MAXD)]


class Contacts(C.Structure):
    _fields_ = [
        ("n", i32), ("point_id", i32 * MAXP), ("link", i32 * MAXP), ("partner", i32 * MAXP),
        ("link_b", i32 * MAXP), ("pos_a", (d * 3) * MAXP), ("pos_b", (d * 3) * MAXP), ("normal", (d * 3) * MAXP), ("dist", d * MAXP), ("friction", d * MAXP),
        ("erp", d * MAXP), ("cfm", d * MAXP), ("impulse", d * MAXP),
    ]


class Box(C.Structure):
    _fields_ = [("center", d * 3), ("R", (d * 3) * 3), ("half", d * 3), ("friction", d), ("stiffness", d),
                ("damping", d), ("id", i32), ("cylinder", i32)]


class Rng(C.Structure):
    _fields_ = [("mt", C.c_uint32 * 624), ("pos", i32)]


class W3DEnv(C.Structure):
    _fields_ = [
        ("s", State), ("warm", d * WARMSZ),
        ("feet_contact", d * 4), ("feet_xyz", (d * 3) * 4), ("body_xyz", d * 3), ("body_rpy", d * 3),
        ("body_vel", d * 3), ("joint_speeds", d * MAXD), ("joints_at_limit", i32), ("mirrored", i32),
        ("robot_state", d * (6 + 2 * MAXD + 4)),
        ("dist", d), ("angle", d), ("stop_frames", d), ("walk_target", d * 3), ("close_count", i32),
        ("linear_potential", d), ("angular_potential", d), ("distance_to_target", d), ("angle_to_target", d),
        ("done", i32), ("eval_mode", i32), ("elapsed", i32),
        ("progress", d), ("posture_penalty", d), ("energy_penalty", d), ("joints_penalty", d), ("tall_bonus", d),
        ("target_bonus", d),
        ("env_rng", Rng), ("robot_rng", Rng), ("rng_aliased", i32), ("rows_sum", d), ("last_contacts", Contacts),
    ]


class StepperEnv(C.Structure):
    _fields_ = [
        ("base", W3DEnv), ("curriculum", i32), ("gain_curriculum", i32), ("terrain", (d * 6) * 20),
        ("plank_index", i32 * 3),
        ("boxes", Box * 6),
        ("next_step_index", i32), ("target_reached_count", i32), ("stop_on_next_step", i32),
        ("set_stop_on_next_step", i32), ("timestep", i32), ("target_reached", i32),
        ("foot_dist_to_target", d * 2), ("targets", (d * 5) * 3), ("step_bonus", d), ("speed_penalty", d),
        ("steps_reached", i32), ("random_reward", i32), ("plank_class", i32),
    ]


class Bar(C.Structure):
    _fields_ = [("center", d * 3), ("axis", d * 3), ("halflen", d), ("radius", d), ("friction", d), ("id", i32)]


class MonkeyEnv(C.Structure):
    _fields_ = [
        ("base", W3DEnv), ("terrain", (d * 4) * 32), ("bars", Bar * 4), ("bar_index", i32 * 4),
        ("next_step_index", i32), ("target_reached_count", i32), ("free_fall_count", i32), ("timestep", i32),
        ("swing_leg", i32), ("pivot_leg", i32), ("target_reached", i32),
        ("foot_dist_to_target", d), ("swing_potential", d), ("targets", (d * 3) * 2),
        ("palm_xyz", (d * 3) * 2), ("palm_quat", (d * 4) * 2), ("step_bonus", d),
    ]


class CassieEnvS(C.Structure):
    _fields_ = [("base", W3DEnv), ("jvel", d * 16), ("rad_angles", d * 16), ("speeds", d * 16), ("potential", d),
                ("initial_z", d), ("alive_rew", d), ("progress_rew", d), ("robot_state", d * 40)]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "mocca_oracle.c")
    hdr = os.path.join(_HERE, "mocca_oracle.h")
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(f) > os.path.getmtime(_LIB_PATH) for f in (src, hdr))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        assert L.orc_sizeof_model() == C.sizeof(Model), (L.orc_sizeof_model(), C.sizeof(Model))
        assert L.orc_sizeof_w3d_env() == C.sizeof(W3DEnv), (L.orc_sizeof_w3d_env(), C.sizeof(W3DEnv))
        assert L.orc_sizeof_stepper_env() == C.sizeof(StepperEnv), (L.orc_sizeof_stepper_env(), C.sizeof(StepperEnv))
        assert L.orc_sizeof_monkey_env() == C.sizeof(MonkeyEnv), (L.orc_sizeof_monkey_env(), C.sizeof(MonkeyEnv))
        assert L.orc_sizeof_cassie_env() == C.sizeof(CassieEnvS), (L.orc_sizeof_cassie_env(), C.sizeof(CassieEnvS))
        L.orc_rng_double.restype = d
        L.orc_rng_uniform.restype = d
        L.orc_rng_uniform.argtypes = [C.c_void_p, d, d]
        L.orc_rng_u32.restype = C.c_uint32
        L.orc_diag_contacts_take.restype = C.c_long
        L.orc_rnea.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, d, C.c_void_p]
        L.orc_energy_momentum.argtypes = [C.c_void_p, C.c_void_p, d, C.c_void_p]
        _lib = L
    return _lib


def _fill(arr, values):
    values = np.asarray(values)
    if values.ndim == 1:
        for k, v in enumerate(values):
            arr[k] = v.item() if hasattr(v, "item") else v
    else:
        for k, row in enumerate(values):
            for j, v in enumerate(row):
                arr[k][j] = float(v)


def model_from_table(t: dict) -> Model:
    m = Model()
    m.n_links, m.n_dof, m.n_geoms, m.n_feet = t["n_links"], t["n_dof"], len(t["geoms"]), len(t["foot_links"])
    _fill(m.parent, np.array(t["parent"], dtype=np.int64))
    _fill(m.joint_type, np.array(t["joint_type"], dtype=np.int64))
    _fill(m.dof_of_link, np.array(t["dof_of_link"], dtype=np.int64))
    _fill(m.link_of_dof, np.array(t["link_of_dof"], dtype=np.int64))
    for name in ("axis", "rot_parent_to_this", "e_vec", "d_vec", "inertia"):
        _fill(getattr(m, "rot_p2t" if name == "rot_parent_to_this" else name), np.array(t[name], dtype=np.float64))
    _fill(m.mass, np.array(t["mass"], dtype=np.float64))
    m.base_mass = float(t["base"]["mass"])
    _fill(m.base_inertia, np.array(t["base"]["inertia"], dtype=np.float64))
    for name in ("lower", "upper", "damping", "armature", "gain"):
        _fill(getattr(m, name), np.array(t[name], dtype=np.float64))
    _fill(m.link_thresh, np.array([t["base"]["contact_threshold"]] + t["contact_threshold"], dtype=np.float64))
    _fill(m.link_group, np.array([t["base"]["group"]] + t["group"], dtype=np.int64))
    _fill(m.link_mask, np.array([t["base"]["mask"]] + t["mask"], dtype=np.int64))
    for g, geom in enumerate(t["geoms"]):
        m.geom_link[g] = geom["link"]
        m.geom_type[g] = geom["type"]
        for k in range(3):
            m.geom_p0[g][k] = geom["p0"][k]
            m.geom_p1[g][k] = geom["p1"][k]
            m.geom_size[g][k] = geom["size"][k]
        for k in range(4):
            m.geom_quat[g][k] = geom["quat"][k]
        m.geom_friction[g] = geom["friction"]
    _fill(m.foot_link, np.array(t["foot_links"], dtype=np.int64))
    _fill(m.base_joint_angles, np.array(t["base_joint_angles"], dtype=np.float64))
    _fill(m.base_position, np.array(t["base_position"], dtype=np.float64))
    _fill(m.base_orientation, np.array(t.get("base_orientation", [0, 0, 0, 1]), dtype=np.float64))
    m.termination_height = float(t.get("termination_height", 0.7))
    m.planar_env = 1.0 if t.get("planar") else 0.0
    _fill(m.stepper_init_position, np.array(t.get("stepper_init_position", [0.3, 0.0, 1.32]), dtype=np.float64))
    m.n_right = len(t["right_joint_indices"])
    _fill(m.right_idx, np.array(t["right_joint_indices"], dtype=np.int64))
    _fill(m.left_idx, np.array(t["left_joint_indices"], dtype=np.int64))
    m.n_neg = len(t["negation_joint_indices"])
    _fill(m.neg_idx, np.array(t["negation_joint_indices"], dtype=np.int64))
    palms = t.get("palm_links", [-1, -1])
    m.palm_link[0], m.palm_link[1] = int(palms[0]), int(palms[1])
    m.n_p2p = len(t.get("p2p", []))
    for k, c in enumerate(t.get("p2p", [])):
        m.p2p_link_a[k], m.p2p_link_b[k] = c["link_a"], c["link_b"]
        for j in range(3):
            m.p2p_pivot_a[k][j] = c["pivot_a"][j]
            m.p2p_pivot_b[k][j] = c["pivot_b"][j]
        m.p2p_max_impulse[k] = c["max_impulse"]
    pairs = t.get("self_pairs", [])
    assert len(pairs) <= MAXSP
    m.n_self = len(pairs)
    for k, (a, b) in enumerate(pairs):
        m.self_a[k], m.self_b[k] = a, b
    hulls = t.get("hulls", [])
    assert len(hulls) <= MAXHULL and len(t.get("hull_pairs", [])) <= MAXHPAIR
    m.n_hulls = len(hulls)
    m.hull_margin = float(t.get("hull_margin", 0.0))
    for k, h in enumerate(hulls):
        v = np.array(h["verts"], dtype=np.float64)
        assert v.shape == (HULLV, 3)
        m.hull_link[k] = h["link"]
        cen = v.mean(0)
        for i in range(HULLV):
            for j in range(3):
                m.hull_verts[k][i][j] = v[i, j]
        for j in range(3):
            m.hull_center[k][j] = cen[j]
        m.hull_radius[k] = float(np.linalg.norm(v - cen, axis=1).max())
        m.hull_friction[k] = float(t["link_friction"][h["link"] + 1])
    m.n_hpairs = len(t.get("hull_pairs", []))
    for k, (a, b) in enumerate(t.get("hull_pairs", [])):
        m.hpair_a[k], m.hpair_b[k] = a, b
    if "ordered_dofs" in t:
        m.n_ordered = len(t["ordered_dofs"])
        _fill(m.ordered_dof, np.array(t["ordered_dofs"], dtype=np.int64))
        pd = t["powered_joint_inds"] + t["spring_joint_inds"]
        m.n_pd = len(pd)
        _fill(m.pd_ordered_index, np.array(pd, dtype=np.int64))
        _fill(m.pd_kp, np.array(t["pd_kp"], dtype=np.float64))
        _fill(m.pd_kd, np.array(t["pd_kd"], dtype=np.float64))
    return m


def default_params() -> Params:
    p = Params()
    lib().orc_default_params(C.byref(p))
    return p


def make_state(n_dof, pos, quat, omega, vel, q, qd) -> State:
    s = State()
    _fill(s.pos, np.asarray(pos, dtype=np.float64))
    _fill(s.quat, np.asarray(quat, dtype=np.float64))
    _fill(s.omega, np.asarray(omega, dtype=np.float64))
    _fill(s.vel, np.asarray(vel, dtype=np.float64))
    _fill(s.q, np.asarray(q, dtype=np.float64))
    _fill(s.qd, np.asarray(qd, dtype=np.float64))
    return s


def state_arrays(s: State, n_dof: int):
    return (np.array(s.pos[:]), np.array(s.quat[:]), np.array(s.omega[:]), np.array(s.vel[:]),
            np.array(s.q[:n_dof]), np.array(s.qd[:n_dof]))


def state_vector(s: State, n_dof: int) -> np.ndarray:
    """[pos3, quat4, omega3, vel3, q, qd] -- same layout as the CUDA library's mb200_get_state."""
    return np.concatenate(state_arrays(s, n_dof))


# ------------------------------------------------------------------ thin functional wrappers
def fk(m: Model, s: State):
    n = m.n_links + 1
    pos = ((d * 3) * n)()
    rot = ((d * 9) * n)()
    lib().orc_fk(C.byref(m), C.byref(s), pos, rot)
    return np.array(pos).reshape(n, 3), np.array(rot).reshape(n, 3, 3)


def forward_dynamics(m, p, s, tau, with_damping=True):
    nu = 6 + m.n_dof
    acc = (d * MAXU)()
    t = (d * MAXD)(*[float(x) for x in tau])
    lib().orc_forward_dynamics(C.byref(m), C.byref(p), C.byref(s), t, int(with_damping), acc)
    return np.array(acc[:nu])


def rnea(m, s, acc, gravity):
    nu = 6 + m.n_dof
    a = (d * MAXU)(*[float(x) for x in acc])
    tau = (d * MAXU)()
    lib().orc_rnea(C.byref(m), C.byref(s), a, float(gravity), tau)
    return np.array(tau[:nu])


def mass_matrix(m, s):
    nu = 6 + m.n_dof
    M = (d * (nu * nu))()
    lib().orc_mass_matrix(C.byref(m), C.byref(s), M)
    return np.array(M).reshape(nu, nu)


def minv_mult(m, p, s, f):
    nu = 6 + m.n_dof
    ff = (d * MAXU)(*[float(x) for x in f])
    out = (d * MAXU)()
    lib().orc_minv_mult(C.byref(m), C.byref(p), C.byref(s), ff, out)
    return np.array(out[:nu])


def collide(m, p, s, boxes=None):
    c = Contacts()
    nb = 0 if boxes is None else len(boxes)
    lib().orc_collide(C.byref(m), C.byref(p), C.byref(s), boxes, nb, C.byref(c))
    return c


def step_physics(m, p, s, tau, boxes=None, warm=None):
    """One stepSimulation (p.substeps substeps). Mutates s. Returns (last contacts, total rows)."""
    t = (d * MAXD)(*[float(x) for x in tau])
    c = Contacts()
    rows = i32(0)
    nb = 0 if boxes is None else len(boxes)
    lib().orc_step_physics(C.byref(m), C.byref(p), C.byref(s), t, boxes, nb, warm, C.byref(c), C.byref(rows))
    return c, rows.value


def step_physics_bars(m, p, s, tau, bars):
    """One stepSimulation with static bars (Monkey3D). Mutates s. Returns (last contacts, total rows)."""
    t = (d * MAXD)(*[float(x) for x in tau])
    c = Contacts()
    rows = i32(0)
    lib().orc_step_physics_bars(C.byref(m), C.byref(p), C.byref(s), t, bars, len(bars), C.byref(c), C.byref(rows))
    return c, rows.value


def energy_momentum(m, s, gravity):
    out = (d * 8)()
    lib().orc_energy_momentum(C.byref(m), C.byref(s), float(gravity), out)
    o = np.array(out)
    return dict(KE=o[0], PE=o[1], P=o[2:5], L=o[5:8])


# ------------------------------------------------------------------ gym-0.21 seeding shim
def contacts_take() -> int:
    """Contact points summed over every oracle substep since the last call (test diagnostics)."""
    return int(lib().orc_diag_contacts_take())


def substep_q_take(n_dof: int) -> np.ndarray:
    """Joint angles at the start of the (last <= 64) oracle substeps since the last call, [k, n_dof]."""
    buf = (d * (64 * MAXD))()
    k = int(lib().orc_diag_q_take(buf))
    return np.array(buf[:k * MAXD]).reshape(k, MAXD)[:, :n_dof]


def gym_seed_words(seed: int):
    """gym.utils.seeding.np_random(seed) -> the uint32 key list handed to RandomState.seed()."""
    import hashlib
    import struct

    seed = int(seed) % 2 ** 64
    h = hashlib.sha512(str(seed).encode("utf8")).digest()[:8]
    h += b"\0" * 4
    unpacked = struct.unpack("3I", h)
    big = sum(2 ** (32 * i) * v for i, v in enumerate(unpacked))
    if big == 0:
        return [0]
    ints = []
    while big > 0:
        big, mod = divmod(big, 2 ** 32)
        ints.append(mod)
    return ints


class Walker3DCustomOracle:
    """Single-env restatement of Walker3DCustomEnv (reference env_locomotion.py:37-222)."""

    def __init__(self, table: dict, seed: int = 0, params: Params | None = None):
        self.table = table
        self.m = model_from_table(table)
        self.p = params or default_params()
        self.e = W3DEnv()
        self.A = table["n_dof"]
        self.obs_dim = 6 + 2 * self.A + len(table["foot_links"]) + 2
        self._seed(seed, True)

    def _seed(self, seed, at_construction):
        words = gym_seed_words(seed)
        key = (C.c_uint32 * len(words))(*words)
        lib().orc_w3d_seed(C.byref(self.e), key, len(words), int(at_construction))

    def seed(self, seed):
        self._seed(seed, False)
        return [seed]

    def reset(self):
        obs = (d * self.obs_dim)()
        lib().orc_w3d_reset(C.byref(self.m), C.byref(self.p), C.byref(self.e), obs)
        return np.array(obs)

    def step(self, action):
        a = (d * MAXD)(*[float(x) for x in action])
        obs = (d * self.obs_dim)()
        r = d(0)
        done = i32(0)
        trunc = i32(0)
        lib().orc_w3d_step(C.byref(self.m), C.byref(self.p), C.byref(self.e), a, obs, C.byref(r), C.byref(done),
                           C.byref(trunc))
        info = {"TimeLimit.truncated": True} if trunc.value else {}
        return np.array(obs), r.value, bool(done.value), info

    def state_vector(self):
        return state_vector(self.e.s, self.A)


class Walker3DStepperOracle:
    """Single-env restatement of Walker3DStepperEnv (reference env_locomotion.py:330-840)."""

    def __init__(self, table: dict, seed: int = 0, curriculum: int = 0, params: Params | None = None,
                 random_reward: bool = False, plank_class: str | None = None):
        self.table = table
        self.m = model_from_table(table)
        self.p = params or default_params()
        self.e = StepperEnv()
        self.e.curriculum = curriculum
        self.e.random_reward = int(random_reward)
        self.e.plank_class = {None: 0, "LargePlank": 0, "Plank": 1, "Pillar": 2}[plank_class]
        self.A = table["n_dof"]
        self.obs_dim = 6 + 2 * self.A + len(table["foot_links"]) + 15
        self._seed(seed, True)

    def _seed(self, seed, at_construction):
        words = gym_seed_words(seed)
        key = (C.c_uint32 * len(words))(*words)
        lib().orc_stepper_seed(C.byref(self.e), key, len(words), int(at_construction))

    def seed(self, seed):
        self._seed(seed, False)
        return [seed]

    def set_env_params(self, params):
        if "curriculum" in params:
            self.e.curriculum = int(params["curriculum"])

    def reset(self):
        obs = (d * self.obs_dim)()
        lib().orc_stepper_reset(C.byref(self.m), C.byref(self.p), C.byref(self.e), obs)
        return np.array(obs)

    def step(self, action):
        a = (d * MAXD)(*[float(x) for x in action])
        obs = (d * self.obs_dim)()
        r = d(0)
        done = i32(0)
        trunc = i32(0)
        lib().orc_stepper_step(C.byref(self.m), C.byref(self.p), C.byref(self.e), a, obs, C.byref(r), C.byref(done),
                               C.byref(trunc))
        info = {}
        if self.e.steps_reached >= 0:
            info["steps_reached"] = self.e.steps_reached
        if trunc.value:
            info["TimeLimit.truncated"] = True
        return np.array(obs), r.value, bool(done.value), info

    def state_vector(self):
        return state_vector(self.e.base.s, self.A)


class Monkey3DOracle:
    """Single-env restatement of Monkey3DCustomEnv (reference env_locomotion.py:1136-1516)."""

    def __init__(self, table: dict, seed: int = 0, params: Params | None = None):
        self.table = table
        self.m = model_from_table(table)
        self.p = params or default_params()
        self.e = MonkeyEnv()
        self.A = table["n_dof"]
        self.obs_dim = 6 + 2 * self.A + 17
        self._seed(seed, True)

    def _seed(self, seed, at_construction):
        words = gym_seed_words(seed)
        key = (C.c_uint32 * len(words))(*words)
        lib().orc_monkey_seed(C.byref(self.e), key, len(words), int(at_construction))

    def seed(self, seed):
        self._seed(seed, False)
        return [seed]

    def reset(self):
        obs = (d * self.obs_dim)()
        lib().orc_monkey_reset(C.byref(self.m), C.byref(self.p), C.byref(self.e), obs)
        return np.array(obs)

    def step(self, action):
        a = (d * MAXD)(*[float(x) for x in action])
        obs = (d * self.obs_dim)()
        r = d(0)
        done = i32(0)
        trunc = i32(0)
        lib().orc_monkey_step(C.byref(self.m), C.byref(self.p), C.byref(self.e), a, obs, C.byref(r), C.byref(done),
                              C.byref(trunc))
        info = {"TimeLimit.truncated": True} if trunc.value else {}
        return np.array(obs), r.value, bool(done.value), info

    def state_vector(self):
        return state_vector(self.e.base.s, self.A)


def cassie_params() -> Params:
    p = Params()
    lib().orc_cassie_params(C.byref(p))
    return p


class CassieOracle:
    """Single-env restatement of CassieEnv-v0 (reference env_cassie.py:285-479, defects fixed by intent: SURVEY App. D
    Q7-Q9)."""

    def __init__(self, table: dict, seed: int = 0, params: Params | None = None):
        self.table = table
        self.m = model_from_table(table)
        self.p = params or cassie_params()
        self.e = CassieEnvS()
        self.A = 10
        self.n_dof = table["n_dof"]
        self.obs_dim = 36

    def seed(self, seed):
        return [seed]

    def reset(self):
        obs = (d * self.obs_dim)()
        lib().orc_cassie_reset(C.byref(self.m), C.byref(self.p), C.byref(self.e), obs)
        return np.array(obs)

    def step(self, action):
        a = (d * 16)(*[float(x) for x in action])
        obs = (d * self.obs_dim)()
        r = d(0)
        done = i32(0)
        trunc = i32(0)
        lib().orc_cassie_step(C.byref(self.m), C.byref(self.p), C.byref(self.e), a, obs, C.byref(r), C.byref(done),
                              C.byref(trunc))
        info = {"AliveRew": self.e.alive_rew, "ProgressRew": self.e.progress_rew}
        if trunc.value:
            info["TimeLimit.truncated"] = True
        return np.array(obs), r.value, bool(done.value), info

    def state_vector(self):
        return state_vector(self.e.base.s, self.n_dof)
