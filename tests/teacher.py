"""Teacher-forced comparison of a kernel backend (the g++ lane-loop emulation on the CPU, or the CUDA library through
its C ABI on the GPU) with the oracle / the reference-recorded traces -- shared by the emulation and the GPU tests.

Every env step outside the tolerance must be EXPLAINED by a discontinuity of the step map that is visible in discrete
or float64 data, or the test fails (no "x % of the steps may be wrong" allowances):

  rows   the oracle and the backend built a different number of constraint rows or found a different number of
         contact points in that env step (a joint reaching its limit, a contact crossing its breaking threshold: the
         f32 and the f64 evaluation landed on different sides);
  cond   the mass matrix is numerically singular along the step (cond(M(q)) >= 1e6: the MJCF importer's massless
         intermediate links of the 2- and 3-hinge shoulder / hip / abdomen joints in gimbal lock, Bullet ignores the
         armature).  There float32 cannot follow float64 (eps_f32 * cond >= 6 %); the oracle itself moves by O(1)
         under 1e-9 input noise.  The error is then bounded by COND_GAIN * cond (relative error <= eps_f32 * cond).

  sens   the step map itself is unstable at that state: the float64 oracle, re-run from the same state perturbed by
         1e-6 (relative; what float32 arithmetic inside one substep amounts to), moves its own observation by `sens`;
         light end links under an active joint limit plus a contact amplify by ~8x per substep within the 5
         under-converged PGS iterations.  The error is then bounded by SENS_GAIN * sens.

Steps with none of these explanations must agree within `tol_obs` / `tol_rew` with identical done flags -- and after EVERY
step whose structure agrees (rows) the backend's integer bookkeeping is read back and compared exactly with the
oracle's."""
import numpy as np

COND_LIMIT = 1e6
COND_GAIN = 6e-8  # allowed observation error per unit of cond(M) on an ill-conditioned step: the float32 unit roundoff
SENS_EPS = 1e-6   # relative size of the input perturbation of the oracle's sensitivity probe
SENS_GAIN = 30.0  # allowed observation error per unit of the oracle's own response to that perturbation


def gym_mt_row(O, seed):
    """[625] uint32 MT19937 state of numpy.random.RandomState(gym_seed_words(seed)) + position."""
    st = np.random.RandomState(O.gym_seed_words(seed)).get_state()
    return np.concatenate([st[1], [st[2]]]).astype(np.uint32)


def path_cond(O, m, A, qs):
    """max cond(M(q)) over the joint angles `qs` [k, A] the oracle had at the start of each substep of the step
    (conditioning does not depend on the base pose: rotating the base is an orthogonal change of coordinates)."""
    worst = 0.0
    for q in qs:
        s = O.make_state(A, [0, 0, 0], [0, 0, 0, 1], [0] * 3, [0] * 3, q, np.zeros(A))
        worst = max(worst, float(np.linalg.cond(O.mass_matrix(m, s))))
    return worst


class Judge:
    """Collects the per-step verdicts of one teacher-forced run."""

    def __init__(self, name, tol_obs, tol_rew, max_explained=0.10, rel_rew=1e-3, bound_rows=5.0):
        self.name, self.tol_obs, self.tol_rew, self.rel_rew = name, tol_obs, tol_rew, rel_rew
        self.max_explained, self.bound_rows = max_explained, bound_rows
        self.n = 0
        self.errs = []
        self.explained = []  # (step, reasons, e_obs)
        self.book_checked = 0

    def step(self, t, e_obs, rew, ref_rew, done, ref_done, rows, ref_rows, nc, ref_nc, cond, sens=None):
        """Verdict for one env step; returns True when the step agrees in structure (bookkeeping must then match)."""
        self.n += 1
        e_rew = abs(rew - ref_rew)
        within = done == ref_done and e_obs < self.tol_obs and e_rew < self.tol_rew + self.rel_rew * abs(ref_rew)
        structure = int(rows) == int(ref_rows) and int(nc) == int(ref_nc)
        if within:
            self.errs.append(e_obs)
            return structure
        reasons = []
        if not structure:
            reasons.append("rows %d/%d contacts %d/%d" % (rows, ref_rows, nc, ref_nc))
        if cond >= COND_LIMIT:
            reasons.append("cond %.2g" % cond)
        sv = None
        if not reasons and sens is not None:
            sv = sens()
            if SENS_GAIN * sv >= min(e_obs, 1.0) or (done != ref_done and sv > 0.1 * self.tol_obs):
                reasons.append("sens %.2g" % sv)
        msg = "%s step %d: obs err %.3g, reward %.5g vs %.5g, done %s vs %s" % (self.name, t, e_obs, rew, ref_rew, done, ref_done)
        assert reasons, "UNEXPLAINED outlier -- " + msg + " (rows %d, contacts %d, cond %.3g, sens %s)" % (
            rows, nc, cond, sv)
        bound = 0.0
        if not structure:
            bound = self.bound_rows
        if cond >= COND_LIMIT:
            bound = max(bound, self.tol_obs + COND_GAIN * cond)
        if sv is not None:
            bound = max(bound, self.bound_rows if SENS_GAIN * sv >= 1.0 else self.tol_obs + SENS_GAIN * sv)
        assert e_obs <= bound, "outlier beyond its bound %.3g -- %s [%s]" % (bound, msg, "; ".join(reasons))
        self.explained.append((t, "; ".join(reasons), e_obs))
        return structure

    def book(self, t, got: dict, ref: dict):
        """Exact comparison of the integer bookkeeping (called for steps whose structure agrees)."""
        self.book_checked += 1
        for k, v in ref.items():
            g = got[k]
            same = np.array_equal(np.asarray(g), np.asarray(v))
            assert same, "%s step %d: bookkeeping field %s: backend %r, oracle %r" % (self.name, t, k, g, v)

    def finish(self, median_below=None):
        assert self.n > 0
        assert len(self.explained) <= self.max_explained * self.n, (self.name, self.explained, self.n)
        if median_below is not None and self.errs:
            assert np.median(self.errs) < median_below, (self.name, np.median(self.errs))
        return self


# ----------------------------------------------------------------------------------------------- record <-> oracle
def _w3d_base(o):
    return o.e.base if hasattr(o.e, "base") else o.e


def fill_record(kind, o, rec):
    """Teacher-force one env's bookkeeping record (ER_* / ES_* / EM_* / EC_* in csrc/mb_env.cuh) from the oracle."""
    ri = rec.view(np.int32)
    b = _w3d_base(o)
    if kind == "custom":
        rec[0:3] = np.array(b.walk_target[:], dtype=np.float32)
        rec[3], rec[4], rec[5] = b.dist, b.angle, b.stop_frames
        ri[6] = b.close_count
        rec[7] = b.linear_potential
        ri[8] = b.elapsed
        rec[9], rec[10] = b.feet_contact[0], b.feet_contact[1]
        rec[17] = b.body_xyz[0]
    elif kind == "stepper":
        rec[0:3] = np.array(b.walk_target[:], dtype=np.float32)
        rec[7] = b.linear_potential
        rec[9], rec[10] = b.feet_contact[0], b.feet_contact[1]
        ri[8] = b.elapsed
        ri[22:27] = (o.e.next_step_index, o.e.target_reached_count, o.e.stop_on_next_step, o.e.set_stop_on_next_step,
                     o.e.timestep)
        ri[6] = o.e.gain_curriculum
        ri[4] = o.e.plank_class
        ri[28:31] = o.e.plank_index[:]
        for p in range(3):
            bx = o.e.boxes[2 * p]
            rec[32 + 12 * p:32 + 12 * p + 3] = np.array(bx.center[:], dtype=np.float32)
            rec[32 + 12 * p + 3:32 + 12 * p + 12] = np.array([list(r) for r in bx.R], dtype=np.float32).ravel()
        rec[68:188] = np.array(o.e.terrain[:], dtype=np.float32).ravel()
    elif kind == "monkey":
        rec[0:3] = np.array(b.walk_target[:], dtype=np.float32)
        rec[9], rec[10] = b.feet_contact[0], b.feet_contact[1]
        ri[8] = b.elapsed
        ri[22:27] = (o.e.next_step_index, o.e.free_fall_count, o.e.timestep, o.e.swing_leg, o.e.pivot_leg)
        rec[27] = o.e.swing_potential
        ri[28:32] = o.e.bar_index[:]
        rec[64:192] = np.array([list(r) for r in o.e.terrain], dtype=np.float32).ravel()
        for k in range(4):
            bar = o.e.bars[k]
            rec[32 + 8 * k:40 + 8 * k] = np.array(list(bar.center) + list(bar.axis) + [bar.halflen, bar.radius],
                                                  dtype=np.float32)
    elif kind == "cassie":
        sv = o.state_vector().astype(np.float32)
        ri[8] = b.elapsed
        rec[22] = o.e.potential
        rec[23], rec[24] = sv[0], sv[1]  # EC_PREVX / EC_PREVY: position at the last calc_potential
        rec[32:46] = np.array(o.e.jvel[:14], dtype=np.float32)
    else:
        raise KeyError(kind)
    return rec


def book_oracle(kind, o):
    """Integer bookkeeping of the oracle env after a step (what the backend's record must hold exactly)."""
    b = _w3d_base(o)
    d = {"elapsed": int(b.elapsed), "feet_contact": [float(b.feet_contact[0]), float(b.feet_contact[1])]}
    if kind == "custom":
        d["close_count"] = int(b.close_count)
    elif kind == "stepper":
        d.update(next_step_index=int(o.e.next_step_index), target_reached_count=int(o.e.target_reached_count),
                 stop_on_next_step=int(o.e.stop_on_next_step), set_stop_on_next_step=int(o.e.set_stop_on_next_step),
                 timestep=int(o.e.timestep), plank_index=[int(x) for x in o.e.plank_index[:]],
                 steps_reached=int(o.e.steps_reached))
    elif kind == "monkey":
        d.update(next_step_index=int(o.e.next_step_index), free_fall_count=int(o.e.free_fall_count),
                 timestep=int(o.e.timestep), swing_leg=int(o.e.swing_leg), pivot_leg=int(o.e.pivot_leg),
                 bar_index=[int(x) for x in o.e.bar_index[:]])
    elif kind == "cassie":
        d.pop("feet_contact")
    return d


def book_record(kind, rec):
    ri = rec.view(np.int32)
    d = {"elapsed": int(ri[8]), "feet_contact": [float(rec[9]), float(rec[10])]}
    if kind == "custom":
        d["close_count"] = int(ri[6])
    elif kind == "stepper":
        d.update(next_step_index=int(ri[22]), target_reached_count=int(ri[23]), stop_on_next_step=int(ri[24]),
                 set_stop_on_next_step=int(ri[25]), timestep=int(ri[26]), plank_index=[int(x) for x in ri[28:31]],
                 steps_reached=int(ri[31]))
    elif kind == "monkey":
        d.update(next_step_index=int(ri[22]), free_fall_count=int(ri[23]), timestep=int(ri[24]), swing_leg=int(ri[25]),
                 pivot_leg=int(ri[26]), bar_index=[int(x) for x in ri[28:32]])
    elif kind == "cassie":
        d.pop("feet_contact")
    return d


# ----------------------------------------------------------------------------------------------- backends
class EmuBackend:
    """tests/emu object (EmuW3D / EmuStepper / EmuMonkey / EmuCassie): the kernel source compiled by g++."""

    def __init__(self, e, width):
        self.e, self.width = e, width

    def force(self, kind, o):
        self.e.state[:self.width] = o.state_vector().astype(np.float32)
        fill_record(kind, o, self.e.rec)
        return float(self.e.rec[15]), float(self.e.rec[16])

    def step(self, a):
        o2, r2, d2, tr2, fin = self.e.step(a)
        return (fin if d2 else o2).astype(np.float64), float(r2), bool(d2)

    def record(self):
        return self.e.rec

    def close(self):
        pass


class GpuBackend:
    """A 1-env VecEnv of mocca_envs_b200 (the CUDA library through its C ABI)."""

    def __init__(self, env):
        import torch

        self.env, self.torch = env, torch

    def force(self, kind, o):
        t = self.torch
        sv = o.state_vector().astype(np.float32)
        self.env.set_state(t.tensor(sv[None]))
        rec = self.env.get_record().cpu().numpy()
        fill_record(kind, o, rec[0])
        self.env.set_record(t.tensor(rec))
        return float(rec[0, 15]), float(rec[0, 16])

    def step(self, a):
        t = self.torch
        obs, rew, done, info = self.env.step(t.tensor(np.asarray(a, dtype=np.float32)[None]))
        d = bool(done[0].item())
        got = (info["terminal_observation"] if d else obs)[0].double().cpu().numpy()
        return got, float(rew[0].item()), d

    def record(self):
        return self.env.get_record().cpu().numpy()[0]

    def close(self):
        self.env.close()


def _perturbed_response(O, o, pre, a, base_obs, A, rng):
    """Largest change of the oracle's own observation when the step is re-run from `pre` (a byte snapshot of the oracle
    env before the step) with the physics state perturbed by SENS_EPS (relative).  Leaves the oracle env untouched."""
    import ctypes as C

    post = bytes(o.e)
    worst = 0.0
    for _ in range(8):  # (the response is heavy-tailed -- a limit or a contact switches for some directions only)
        C.memmove(C.byref(o.e), pre, len(pre))
        s = _w3d_base(o).s
        for k in range(3):
            s.pos[k] += SENS_EPS * rng.randn() * (1 + abs(s.pos[k]))
            s.omega[k] += SENS_EPS * rng.randn() * (1 + abs(s.omega[k]))
            s.vel[k] += SENS_EPS * rng.randn() * (1 + abs(s.vel[k]))
        qn = np.array([s.quat[k] + SENS_EPS * rng.randn() for k in range(4)])
        qn /= np.linalg.norm(qn)
        for k in range(4):
            s.quat[k] = qn[k]
        for k in range(A):
            s.q[k] += SENS_EPS * rng.randn() * (1 + abs(s.q[k]))
            s.qd[k] += SENS_EPS * rng.randn() * (1 + abs(s.qd[k]))
        o2, _, _, _ = o.step(np.asarray(a, dtype=np.float64))
        worst = max(worst, float(np.abs(o2 - base_obs).max()))
    C.memmove(C.byref(o.e), post, len(post))
    O.contacts_take()
    O.substep_q_take(A)
    return worst


def _set_oracle_state(o, A, sv):
    s = _w3d_base(o).s
    for k in range(3):
        s.pos[k], s.omega[k], s.vel[k] = float(sv[k]), float(sv[7 + k]), float(sv[10 + k])
    for k in range(4):
        s.quat[k] = float(sv[3 + k])
    for k in range(A):
        s.q[k], s.qd[k] = float(sv[13 + k]), float(sv[13 + A + k])


def _round_oracle_state(o, A):
    """Overwrite the oracle's physics state with its own float32 rounding (what the backend is given)."""
    s = _w3d_base(o).s
    for arr, n in ((s.pos, 3), (s.quat, 4), (s.omega, 3), (s.vel, 3), (s.q, A), (s.qd, A)):
        for k in range(n):
            arr[k] = float(np.float32(arr[k]))


def run_teacher_forced(kind, O, table, o, backend, actions, judge, refs=None, teleports=None, obs_err=None,
                       skip_compare=(), book_on_done=False, round_oracle=False, on_step=None, force_states=False):
    """Drive oracle `o` along `actions`; before every step hand the backend the oracle's state and bookkeeping; compare
    the backend's observation / reward / done with `refs` (the reference-recorded obs / rewards / dones of a golden
    trace) or, when refs is None, with the oracle's own.  `skip_compare`: steps whose float comparison is waived (their
    bookkeeping is still compared).  Returns the judge."""
    A = table["n_dof"]
    m = o.m
    k = 1
    obs_err = obs_err or (lambda got, ref: float(np.abs(got - ref).max()))
    for t, a in enumerate(actions):
        if teleports and t in teleports:
            teleports[t](o)
        if refs is not None and "states" in getattr(refs, "files", refs) and force_states:
            # a trace recorded from another physics engine (PyBullet): oracle and backend both restart every step from
            # the RECORDED state, so the comparison is one env step deep and errors do not accumulate
            _set_oracle_state(o, A, refs["states"][k - 1])
        if round_oracle:
            _round_oracle_state(o, A)
        rows0, nc0 = backend.force(kind, o)
        pre = bytes(o.e)
        got, rew, done = backend.step(a)
        O.contacts_take()
        O.substep_q_take(A)
        o1, r1, d1, _ = o.step(np.asarray(a, dtype=np.float64))
        ref_nc = O.contacts_take()
        qs = O.substep_q_take(A)
        ref_rows = int(_w3d_base(o).rows_sum)
        rec = backend.record()
        rows, nc = rec[15] - rows0, rec[16] - nc0
        if refs is not None:
            ref_obs, ref_r, ref_d = refs["obs"][k], float(refs["rewards"][t]), bool(refs["dones"][t])
            assert force_states or d1 == ref_d, "the oracle left the recorded trajectory at step %d" % t
        else:
            ref_obs, ref_r, ref_d = o1, r1, d1
        e_obs = obs_err(got, ref_obs)
        if t in skip_compare:
            structure = int(rows) == ref_rows and int(nc) == ref_nc
        else:
            needs_cond = not (done == ref_d and e_obs < judge.tol_obs)
            cond = path_cond(O, m, A, qs) if needs_cond or abs(rew - ref_r) >= judge.tol_rew else 0.0
            sens = lambda: _perturbed_response(O, o, pre, a, o1, A, np.random.RandomState(t))
            structure = judge.step(t, e_obs, rew, ref_r, done, ref_d, rows, ref_rows, nc, ref_nc, cond, sens)
        if structure and done == ref_d and (not done or book_on_done):
            # (on a done step the backend has already auto-reset: its record then belongs to the next episode)
            judge.book(t, book_record(kind, rec), book_oracle(kind, o))
        if on_step is not None:
            on_step(t, backend, o, done, d1)
        k += 1
        if d1:
            o.reset()
            k += 1
    return judge


# ----------------------------------------------------------------------------------------------- cases
# env name -> (record kind, model table, oracle class, emu class name, VecEnv class name)
SPECS = {
    "walker3d": ("custom", "walker3d", "Walker3DCustomOracle", "EmuW3D", "Walker3DCustomVecEnv"),
    "child3d": ("custom", "child3d", "Walker3DCustomOracle", "EmuChild", "Child3DCustomVecEnv"),
    "walker2d": ("custom", "walker2d", "Walker3DCustomOracle", "EmuWalker2D", "Walker2DCustomVecEnv"),
    "crab2d": ("custom", "crab2d", "Walker3DCustomOracle", "EmuCrab2D", "Crab2DCustomVecEnv"),
    "stepper": ("stepper", "walker3d", "Walker3DStepperOracle", "EmuStepper", "Walker3DStepperVecEnv"),
    "mike": ("stepper", "mike", "Walker3DStepperOracle", "EmuMike", "MikeStepperVecEnv"),
    "monkey": ("monkey", "monkey3d", "Monkey3DOracle", "EmuMonkey", "Monkey3DCustomVecEnv"),
    "cassie": ("cassie", "cassie", "CassieOracle", "EmuCassie", "CassieVecEnv"),
}
_TABLES = {}


def table_of(name):
    import os

    from mocca_envs_b200.model_compiler import load_table

    if name not in _TABLES:
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        _TABLES[name] = load_table(os.path.join(root, "mocca_envs_b200", "models", name + ".json"))
    return _TABLES[name]


def monkey_obs_err(got, ref):
    """Monkey3D observation error with the swing palm's quaternion (last four entries) compared up to its sign."""
    err = float(np.abs(got[:65] - ref[:65]).max())
    return max(err, float(min(np.abs(got[65:] - ref[65:]).max(), np.abs(got[65:] + ref[65:]).max())))


def make_pair(O, env_name, backend, seed, seed2=None, plank_class=None, curriculum=None, eval_mode=False,
              random_reward=False):
    """(oracle env, backend) of `env_name`, seeded alike: constructed with `seed`, then -- if seed2 is given -- re-seeded
    through EnvBase.seed(seed2) (quirk Q1: the robot keeps the construction stream).  backend: "emu" or "gpu"."""
    kind, tname, ocls, ecls, vcls = SPECS[env_name]
    t = table_of(tname)
    okw, vkw = {}, {}
    if plank_class not in (None, "LargePlank"):
        okw["plank_class"] = vkw["plank_class"] = plank_class
    if random_reward:
        okw["random_reward"] = vkw["random_reward"] = True
    o = getattr(O, ocls)(t, **okw) if kind == "cassie" else getattr(O, ocls)(t, seed=seed, **okw)
    if seed2 is not None and kind != "cassie":
        o.seed(seed2)
    if curriculum is not None:
        o.set_env_params({"curriculum": curriculum})
    if eval_mode:
        o.e.eval_mode = 1
    if backend == "emu":
        from tests.emu import emu as E

        cls = getattr(E, ecls)
        if plank_class == "Pillar":
            cls = type("EmuPillar", (cls,), {"prefix": "pillar" if env_name == "stepper" else "mike_pillar"})
        if kind == "cassie":
            e = cls()
        elif kind == "stepper":
            e = cls(gym_mt_row(O, seed), curriculum=curriculum or 0)
        else:
            e = cls(gym_mt_row(O, seed))
        if seed2 is not None and kind != "cassie":
            e.mt[1, :625] = gym_mt_row(O, seed)
            e.mt[0, :625] = gym_mt_row(O, seed2)
            e.rec.view(np.int32)[11] = 0  # ER_ALIASED
        if eval_mode:
            e.rec.view(np.int32)[18] = 1  # ER_EVAL
        if random_reward:
            e.rec.view(np.int32)[3] = 1  # ES_RANDOM_REWARD
        b = EmuBackend(e, 13 + 2 * t["n_dof"])
        b.reset = e.reset
    else:
        from mocca_envs_b200 import vec_env as V

        env = getattr(V, vcls)(1, device="cuda:0", seed=seed, return_final_obs=True, **vkw)
        if seed2 is not None and kind != "cassie":
            env.seed(seed2)
        if curriculum is not None:
            env.set_env_params({"curriculum": curriculum})
        if eval_mode:
            env.evaluation_mode()
        b = GpuBackend(env)
        b.reset = lambda: env.reset()[0].double().cpu().numpy()
    return kind, t, o, b


def env_name_of_fixture(basename):
    for key, name in (("Walker3DCustomEnv", "walker3d"), ("Walker3DStepperEnv", "stepper"), ("Monkey3DCustomEnv", "monkey"),
                      ("CassieEnv", "cassie"), ("child3d", "child3d"), ("walker2d", "walker2d"), ("crab2d", "crab2d"), ("mike", "mike"),
                      ("walker3d_stepper", "stepper"), ("monkey3d", "monkey"), ("cassie", "cassie"),
                      ("walker3d_custom", "walker3d")):
        if key in basename:
            return name
    raise KeyError(basename)


# tolerances per record kind: observation / reward; the Cassie observation carries raw joint speeds in rad/s over 50
# PD substeps per env step
# PD substeps per env step (device vs oracle from f32-identical states, 8 seeds x 25 steps, tools/dbg_teacher_errs.py:
# median 1.1e-3, p99 6e-3, max 1.0e-2 .. 1.4e-2 from build to build at cond(M) = 7e4)
TOL = {"custom": (1e-3, 1e-2), "stepper": (1e-3, 1e-2), "monkey": (1e-3, 1e-2), "cassie": (2e-2, 2e-3)}


def run_golden_trace(O, path, backend, force_states=False):
    """One reference-recorded fixture (tests/golden/ref_*.npz) teacher-forced through `backend` ("emu" / "gpu"): the
    backend's observation / reward / done per step against the values the REFERENCE's own code recorded."""
    import os

    from tests.test_reference_golden import _monkey_grab, _teleport

    g = np.load(path)
    base = os.path.basename(path)
    name = env_name_of_fixture(base)
    kind = SPECS[name][0]
    kw = {}
    if kind == "stepper":
        kw = dict(plank_class=str(g["plank_class"]), curriculum=int(g["curriculum"]),
                  random_reward=bool(int(g["random_reward"])) if "random_reward" in g.files else False)
    if kind == "custom":
        kw = dict(eval_mode=bool(int(g["eval_mode"])))
    if kind == "cassie":
        kind, t, o, b = make_pair(O, name, backend, 0)
    else:
        kind, t, o, b = make_pair(O, name, backend, int(g["construction_seed"]), int(g["seed"]), **kw)
    first = b.reset()
    oref = o.reset()
    # same seeds on both sides: the reset draws (pose noise, target, terrain) agree before any forcing
    err0 = monkey_obs_err(np.asarray(first, dtype=np.float64), oref) if kind == "monkey" else float(np.abs(first - oref).max())
    assert err0 < 1e-4, (base, "first observation", err0)
    tele = {}
    if "teleports" in g.files:
        for r in g["teleports"]:
            if kind == "monkey":
                tele[int(r[0])] = (lambda oo, pos=r[1:4]: _monkey_grab(oo, pos))
            else:
                tele[int(r[0])] = (lambda oo, pos=r[1:4]: _teleport(oo, t, pos))
    tol = TOL[kind]
    # (the "_target" trace drops the walker onto the contact threshold at every step to hold it at the walk target:
    # its steps straddle the contact discontinuity systematically -- every one of them is still verified as such)
    judge = Judge(base + ":" + backend, tol[0], tol[1], rel_rew=0.0 if kind == "cassie" else 1e-3, bound_rows=10.0,
                  max_explained=0.30 if "_target" in base else 0.10)
    run_teacher_forced(kind, O, t, o, b, g["actions"], judge, refs=g, teleports=tele,
                       obs_err=monkey_obs_err if kind == "monkey" else None,
                       skip_compare=set(tele) if kind == "monkey" else (), force_states=force_states)
    b.close()
    return judge.finish(median_below=3e-3 if kind == "cassie" else 5e-4)


def run_vs_oracle(O, env_name, backend, seeds, steps, action_fn, on_step=None, **kw):
    """`env_name` teacher-forced against the oracle ITSELF (no recorded trace) from f32-identical states, one env per
    seed; action_fn(rng, step) -> action.  Returns the judges."""
    judges = []
    for s in seeds:
        kind, t, o, b = make_pair(O, env_name, backend, s, **kw)
        b.reset()
        o.reset()
        rng = np.random.RandomState(1000 + s)
        acts = [action_fn(rng, k) for k in range(steps)]
        tol = TOL[kind]
        judge = Judge("%s:%s:seed%d" % (env_name, backend, s), tol[0], tol[1], rel_rew=0.0 if kind == "cassie" else 1e-3,
                      bound_rows=10.0, max_explained=0.30 if env_name == "child3d" else 0.15)
        # (the child's 30 kg of small links, without Bullet-ignored armature, sit at cond(M) >= 1e6 in a fifth of the
        # steps under full-scale random torques: each such step is verified and bounded, their share is not a defect)
        run_teacher_forced(kind, O, t, o, b, acts, judge, refs=None, round_oracle=True, on_step=on_step,
                           obs_err=monkey_obs_err if kind == "monkey" else None)
        b.close()
        judges.append(judge.finish())
    return judges
