"""ctypes binding of the lane-loop emulation of the kernel source (tests only; see emu.cpp)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libmb_emu.so")
_SRC = [os.path.join(_HERE, "emu.cpp")] + [
    os.path.join(_HERE, "..", "..", "mocca_envs_b200", "csrc", f)
    for f in ("mb_core.cuh", "mb_env.cuh", "mb_tables.h", "generated/walker3d_model.h", "generated/monkey3d_model.h",
              "generated/cassie_model.h", "generated/child3d_model.h", "generated/mike_model.h",
              "generated/walker2d_model.h", "generated/crab2d_model.h")]


class Phys(C.Structure):
    _fields_ = [("dt", C.c_float), ("substeps", C.c_int), ("iterations", C.c_int), ("gravity", C.c_float),
                ("erp_contact", C.c_float), ("erp_joint", C.c_float), ("linear_slop", C.c_float),
                ("lin_damping", C.c_float), ("ang_damping", C.c_float), ("max_coord_vel", C.c_float),
                ("limit_max_impulse", C.c_float), ("split_threshold", C.c_float), ("residual_threshold", C.c_float),
                ("ground_friction", C.c_float), ("has_ground", C.c_int), ("box_friction", C.c_float),
                ("box_erp", C.c_float), ("box_cfm", C.c_float), ("bar_friction", C.c_float), ("self_collision", C.c_int),
                ("warmstart", C.c_float)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        stale = (not os.path.exists(_LIB)) or any(os.path.getmtime(s) > os.path.getmtime(_LIB) for s in _SRC)
        if stale:
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off",
                                   "-Wno-unused-variable", "-o", _LIB, _SRC[0]])
        _lib = C.CDLL(_LIB)
        assert _lib.emu_sizeof_phys() == C.sizeof(Phys)
    return _lib


def default_phys():
    p = Phys()
    lib().emu_default_phys(C.byref(p))
    return p


def _fp(a):
    return a.ctypes.data_as(C.c_void_p)


def step_physics(p, state, tau):
    state = np.ascontiguousarray(state, dtype=np.float32).copy()
    buf = np.zeros(64, dtype=np.float32)
    buf[: len(state)] = state
    tau = np.ascontiguousarray(tau, dtype=np.float32)
    rows, nc = C.c_int(0), C.c_int(0)
    lib().emu_step_physics(C.byref(p), _fp(buf), _fp(tau), C.byref(rows), C.byref(nc))
    return buf[: len(state)].copy(), rows.value, nc.value


def step_physics_warm(p, state, tau, warm):
    """stepSimulation with MbPhysics.warmstart: `warm` (float32[384], updated in place) carries the contact impulses."""
    state = np.ascontiguousarray(state, dtype=np.float32).copy()
    buf = np.zeros(64, dtype=np.float32)
    buf[: len(state)] = state
    tau = np.ascontiguousarray(tau, dtype=np.float32)
    rows, nc = C.c_int(0), C.c_int(0)
    lib().emu_step_physics_warm(C.byref(p), _fp(buf), _fp(tau), _fp(warm), C.byref(rows), C.byref(nc))
    return buf[: len(state)].copy(), rows.value, nc.value


def mass_matrix(p, state, nu):
    buf = np.zeros(64, dtype=np.float32)
    buf[: len(state)] = state
    M = np.zeros((nu, nu), dtype=np.float32)
    b = np.zeros(nu, dtype=np.float32)
    lib().emu_mass_matrix(C.byref(p), _fp(buf), _fp(M), _fp(b))
    return M, b


class EmuW3D:
    prefix = "w3d"  # emu_<prefix>_reset / _step in emu.cpp

    def __init__(self, mt_state, obs_dim=52, act_dim=21, phys=None):
        self.p = phys or default_phys()
        self.state = np.zeros(64, dtype=np.float32)
        self.rec = np.zeros(32, dtype=np.float32)
        self.mt = np.zeros((2, 640), dtype=np.uint32)
        self.mt[0, :625] = mt_state
        self.mt[1, 624] = 624
        self.rec.view(np.int32)[11] = 1  # ER_ALIASED
        self.obs_dim, self.act_dim = obs_dim, act_dim
        self.stats = np.zeros(4)

    def reset(self):
        obs = np.zeros(self.obs_dim, dtype=np.float32)
        getattr(lib(), "emu_%s_reset" % self.prefix)(C.byref(self.p), _fp(self.state), _fp(self.rec), _fp(self.mt[0]), _fp(self.mt[1]), _fp(obs))
        return obs

    def step(self, act):
        act = np.ascontiguousarray(act, dtype=np.float32)
        obs = np.zeros(self.obs_dim, dtype=np.float32)
        fin = np.zeros(self.obs_dim, dtype=np.float32)
        rew = np.zeros(1, dtype=np.float32)
        done = np.zeros(1, dtype=np.uint8)
        trunc = np.zeros(1, dtype=np.uint8)
        st = np.zeros(4)
        getattr(lib(), "emu_%s_step" % self.prefix)(C.byref(self.p), _fp(self.state), _fp(self.rec), _fp(self.mt[0]),
                                                    _fp(self.mt[1]), _fp(act), _fp(obs), _fp(rew), _fp(done), _fp(trunc),
                                                    _fp(fin), _fp(st))
        self.stats += st
        return obs, float(rew[0]), bool(done[0]), bool(trunc[0]), fin


class EmuChild(EmuW3D):
    """Child3DCustomEnv-v0: the Walker3DCustomEnv template on the child3d table."""
    prefix = "child"


class EmuWalker2D(EmuW3D):
    """Walker2DCustomEnv-v0: the Walker3DCustomEnv template on the planar walker2d table."""
    prefix = "walker2d"

    def __init__(self, mt_state, phys=None):
        super().__init__(mt_state, obs_dim=24, act_dim=7, phys=phys)


class EmuCrab2D(EmuW3D):
    """Crab2DCustomEnv-v0."""
    prefix = "crab2d"

    def __init__(self, mt_state, phys=None):
        super().__init__(mt_state, obs_dim=22, act_dim=6, phys=phys)


def model_step_physics(prefix, p, state, tau, rec=None):
    """stepSimulation through emu_<prefix>_step_physics (child: ground plane; mike: the planks of a Stepper record)."""
    buf = np.zeros(64, dtype=np.float32)
    buf[: len(state)] = state
    tau = np.ascontiguousarray(tau, dtype=np.float32)
    rec = np.zeros(192, dtype=np.float32) if rec is None else rec
    rows, nc = C.c_int(0), C.c_int(0)
    getattr(lib(), "emu_%s_step_physics" % prefix)(C.byref(p), _fp(buf), _fp(rec), _fp(tau), C.byref(rows), C.byref(nc))
    return buf[: len(state)].copy(), rows.value, nc.value


def model_mass_matrix(prefix, p, state, nu=27):
    buf = np.zeros(64, dtype=np.float32)
    buf[: len(state)] = state
    M = np.zeros((nu, nu), dtype=np.float32)
    b = np.zeros(nu, dtype=np.float32)
    getattr(lib(), "emu_%s_mass_matrix" % prefix)(C.byref(p), _fp(buf), _fp(M), _fp(b))
    return M, b


def stepper_phys():
    p = Phys()
    lib().emu_stepper_phys(C.byref(p))
    return p


class EmuStepper:
    """Walker3DStepperEnv through the emulated kernel source (record layout: ER_* / ES_* in mb_env.cuh)."""

    prefix = "stepper"

    ES_NEXT, ES_COUNT, ES_STOP, ES_SETSTOP, ES_TIMESTEP, ES_CURRIC, ES_PLANKIDX, ES_STEPS = 22, 23, 24, 25, 26, 27, 28, 31
    ES_BOX, ES_TERRAIN = 32, 68

    def __init__(self, mt_state, curriculum=0):
        self.p = stepper_phys()
        self.state = np.zeros(64, dtype=np.float32)
        self.stride = lib().emu_stepper_rec_stride()
        self.rec = np.zeros(self.stride, dtype=np.float32)
        self.mt = np.zeros((2, 640), dtype=np.uint32)
        self.mt[0, :625] = mt_state
        self.mt[1, 624] = 624
        self.rec.view(np.int32)[11] = 1  # ER_ALIASED
        self.rec.view(np.int32)[self.ES_CURRIC] = curriculum
        self.obs_dim = 65

    def reset(self):
        obs = np.zeros(self.obs_dim, dtype=np.float32)
        getattr(lib(), "emu_%s_reset" % self.prefix)(C.byref(self.p), _fp(self.state), _fp(self.rec), _fp(self.mt[0]), _fp(self.mt[1]), _fp(obs))
        return obs

    def step(self, act):
        act = np.ascontiguousarray(act, dtype=np.float32)
        obs = np.zeros(self.obs_dim, dtype=np.float32)
        fin = np.zeros(self.obs_dim, dtype=np.float32)
        rew = np.zeros(1, dtype=np.float32)
        done = np.zeros(1, dtype=np.uint8)
        trunc = np.zeros(1, dtype=np.uint8)
        st = np.zeros(4)
        getattr(lib(), "emu_%s_step" % self.prefix)(C.byref(self.p), _fp(self.state), _fp(self.rec), _fp(self.mt[0]),
                                                    _fp(self.mt[1]), _fp(act), _fp(obs), _fp(rew), _fp(done), _fp(trunc),
                                                    _fp(fin), _fp(st))
        return obs, float(rew[0]), bool(done[0]), bool(trunc[0]), fin

    def terrain(self):
        return self.rec[self.ES_TERRAIN:self.ES_TERRAIN + 120].reshape(20, 6)

    def step_physics(self, tau):
        tau = np.ascontiguousarray(tau, dtype=np.float32)
        rows, nc = C.c_int(0), C.c_int(0)
        getattr(lib(), "emu_%s_step_physics" % self.prefix)(C.byref(self.p), _fp(self.state), _fp(self.rec), _fp(tau), C.byref(rows), C.byref(nc))
        return rows.value, nc.value


class EmuMike(EmuStepper):
    """MikeStepperEnv-v0: the Walker3DStepperEnv template on the mike table."""
    prefix = "mike"


class EmuMonkey:
    """Monkey3DCustomEnv through the emulated kernel source (record layout: ER_* / EM_* in mb_env.cuh)."""

    EM_NEXT, EM_FREEFALL, EM_TIMESTEP, EM_SWING, EM_PIVOT, EM_SWINGPOT, EM_BARIDX, EM_BAR, EM_TERRAIN = \
        22, 23, 24, 25, 26, 27, 28, 32, 64

    def __init__(self, mt_state):
        self.p = default_phys()
        self.state = np.zeros(64, dtype=np.float32)
        self.stride = lib().emu_monkey_rec_stride()
        self.rec = np.zeros(self.stride, dtype=np.float32)
        self.mt = np.zeros((2, 640), dtype=np.uint32)
        self.mt[0, :625] = mt_state
        self.mt[1, 624] = 624
        self.rec.view(np.int32)[11] = 1  # ER_ALIASED
        self.obs_dim, self.act_dim = 69, 23

    def reset(self):
        obs = np.zeros(self.obs_dim, dtype=np.float32)
        lib().emu_monkey_reset(C.byref(self.p), _fp(self.state), _fp(self.rec), _fp(self.mt[0]), _fp(self.mt[1]), _fp(obs))
        return obs

    def step(self, act):
        act = np.ascontiguousarray(act, dtype=np.float32)
        obs = np.zeros(self.obs_dim, dtype=np.float32)
        fin = np.zeros(self.obs_dim, dtype=np.float32)
        rew = np.zeros(1, dtype=np.float32)
        done = np.zeros(1, dtype=np.uint8)
        trunc = np.zeros(1, dtype=np.uint8)
        st = np.zeros(4)
        lib().emu_monkey_step(C.byref(self.p), _fp(self.state), _fp(self.rec), _fp(self.mt[0]), _fp(self.mt[1]),
                              _fp(act), _fp(obs), _fp(rew), _fp(done), _fp(trunc), _fp(fin), _fp(st))
        return obs, float(rew[0]), bool(done[0]), bool(trunc[0]), fin

    def terrain(self):
        return self.rec[self.EM_TERRAIN:self.EM_TERRAIN + 128].reshape(32, 4)

    def step_physics(self, tau):
        tau = np.ascontiguousarray(tau, dtype=np.float32)
        rows, nc = C.c_int(0), C.c_int(0)
        lib().emu_monkey_step_physics(C.byref(self.p), _fp(self.state), _fp(self.rec), _fp(tau), C.byref(rows), C.byref(nc))
        return rows.value, nc.value


def monkey_mass_matrix(p, state, nu=29):
    buf = np.zeros(64, dtype=np.float32)
    buf[: len(state)] = state
    M = np.zeros((nu, nu), dtype=np.float32)
    b = np.zeros(nu, dtype=np.float32)
    lib().emu_monkey_mass_matrix(C.byref(p), _fp(buf), _fp(M), _fp(b))
    return M, b


def cassie_phys():
    p = Phys()
    lib().emu_cassie_phys(C.byref(p))
    return p


class EmuCassie:
    """CassieEnv-v0 through the emulated kernel source (record layout: ER_* / EC_* in mb_env.cuh)."""

    EC_POTENTIAL, EC_JVEL = 22, 32

    def __init__(self):
        self.p = cassie_phys()
        self.state = np.zeros(64, dtype=np.float32)
        self.stride = lib().emu_cassie_rec_stride()
        self.rec = np.zeros(self.stride, dtype=np.float32)
        self.obs_dim, self.act_dim, self.n_dof = 36, 10, 18

    def reset(self):
        obs = np.zeros(self.obs_dim, dtype=np.float32)
        lib().emu_cassie_reset(C.byref(self.p), _fp(self.state), _fp(self.rec), _fp(obs))
        return obs

    def step(self, act):
        act = np.ascontiguousarray(act, dtype=np.float32)
        obs = np.zeros(self.obs_dim, dtype=np.float32)
        fin = np.zeros(self.obs_dim, dtype=np.float32)
        rew = np.zeros(1, dtype=np.float32)
        done = np.zeros(1, dtype=np.uint8)
        trunc = np.zeros(1, dtype=np.uint8)
        st = np.zeros(4)
        lib().emu_cassie_step(C.byref(self.p), _fp(self.state), _fp(self.rec), _fp(act), _fp(obs), _fp(rew), _fp(done),
                              _fp(trunc), _fp(fin), _fp(st))
        return obs, float(rew[0]), bool(done[0]), bool(trunc[0]), fin

    def step_physics(self, tau):
        tau = np.ascontiguousarray(tau, dtype=np.float32)
        rows, nc = C.c_int(0), C.c_int(0)
        lib().emu_cassie_step_physics(C.byref(self.p), _fp(self.state), _fp(tau), C.byref(rows), C.byref(nc))
        return rows.value, nc.value


def cassie_step_physics_points(p, state, tau):
    """(state after, rows, contacts, points[16, 10]) of one Cassie stepSimulation through the kernel source."""
    buf = np.zeros(64, dtype=np.float32)
    buf[: len(state)] = state
    tau = np.ascontiguousarray(tau, dtype=np.float32)
    rows, nc = C.c_int(0), C.c_int(0)
    pts = np.zeros((16, 10), dtype=np.float32)
    lib().emu_cassie_step_physics_points(C.byref(p), _fp(buf), _fp(tau), C.byref(rows), C.byref(nc), _fp(pts))
    return buf[: len(state)].copy(), rows.value, nc.value, pts


def cassie_mass_matrix(p, state, nu=24):
    buf = np.zeros(64, dtype=np.float32)
    buf[: len(state)] = state
    M = np.zeros((nu, nu), dtype=np.float32)
    b = np.zeros(nu, dtype=np.float32)
    lib().emu_cassie_mass_matrix(C.byref(p), _fp(buf), _fp(M), _fp(b))
    return M, b
