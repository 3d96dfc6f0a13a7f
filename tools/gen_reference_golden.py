#!/usr/bin/env python
"""Golden traces of the reference's OWN Python env layer (tests/golden/ref_*.npz).

The reference (/root/reference/mocca_envs, read-only, only present in the build container) cannot run as shipped:
its arithmetic below the env layer is the third-party `pybullet` C extension, which is not installable here.  This
script imports the reference's unmodified modules (env_base.py, env_locomotion.py, robots.py, bullet_utils.py) with
stand-ins for `gym` (spaces, seeding) and `pybullet`: the stand-in Bullet client answers the calls the env layer makes
(loadMJCF / getJointInfo / getJointStates / getLinkState / getBasePositionAndOrientation / getBaseVelocity /
getContactPoints / resetJointState / resetBase... / setJointMotorControlArray / stepSimulation ...) from the float64
CPU oracle's physics (oracle/mocca_oracle.c).  Everything ABOVE that boundary -- apply_action, calc_state, reset draws,
target logic, rewards, termination, the stepping-stone terrain generator -- is then computed by the reference's real
code, and recorded.

What the traces pin: the oracle's restatement of the env layer against the reference's own
code on identical physics (tests/test_reference_golden.py replays the recorded actions through the oracle env and
compares observation, reward and done).  What they do not pin: Bullet's arithmetic (still "parity unpinned").

usage: python tools/gen_reference_golden.py            (writes tests/golden/ref_*.npz)
"""
import ctypes as C
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from mocca_envs_b200.model_compiler import JOINT_REVOLUTE as MC_REVOLUTE, load_table  # noqa: E402
from oracle import oracle as O  # noqa: E402

MODELS = os.path.join(ROOT, "mocca_envs_b200", "models")
JOINT_REVOLUTE, JOINT_FIXED = 0, 4  # pybullet constants


# ----------------------------------------------------------------------------------------------- gym stand-in
def install_gym():
    gym = types.ModuleType("gym")

    class Env:
        metadata = {}

    class Box:
        def __init__(self, low, high, dtype=np.float32):
            self.low, self.high = np.asarray(low, dtype=dtype), np.asarray(high, dtype=dtype)
            self.shape, self.dtype = self.low.shape, np.dtype(dtype)

    spaces = types.ModuleType("gym.spaces")
    spaces.Box = Box
    utils = types.ModuleType("gym.utils")
    seeding = types.ModuleType("gym.utils.seeding")

    def np_random(seed=None):  # gym <= 0.21: RandomState seeded with _int_list_from_bigint(hash_seed(seed))
        assert seed is not None, "the generator always seeds explicitly"
        rng = np.random.RandomState()
        rng.seed(O.gym_seed_words(seed))
        return rng, seed

    seeding.np_random = np_random
    utils.seeding = seeding
    envs = types.ModuleType("gym.envs")
    registration = types.ModuleType("gym.envs.registration")
    registration.register = lambda id, **kw: None
    registration.registry = types.SimpleNamespace(env_specs={})
    envs.registration = registration
    envs.registry = types.SimpleNamespace(env_specs={})
    gym.Env, gym.spaces, gym.utils, gym.envs = Env, spaces, utils, envs
    for name, mod in (("gym", gym), ("gym.spaces", spaces), ("gym.utils", utils), ("gym.utils.seeding", seeding),
                      ("gym.envs", envs), ("gym.envs.registration", registration)):
        sys.modules[name] = mod


# ----------------------------------------------------------------------------------------------- pybullet stand-in
def mat_to_quat(R):
    from mocca_envs_b200.model_compiler import mat_to_quat as m2q

    return m2q(np.asarray(R))


class FakeWorld:
    """One Bullet world: the robot (a model table + an oracle state) on the ground plane or on stepping stones."""

    def __init__(self):
        self.bodies = {}  # id -> dict(kind=...)
        self.next_id = 0
        self.params = O.default_params()
        self.params.has_ground = 0  # until plane_stadium.sdf is loaded (remove_ground=True envs never load it)
        self.planks = []  # body ids of the stepping stones, in creation order
        self.bars = []    # body ids of the monkey bars, in creation order
        self.shapes = []
        self.robot = None
        self.tau = None
        self.contacts = None
        self.warm = None
        self.saved = {}

    def new_body(self, **kw):
        i = self.next_id
        self.next_id += 1
        self.bodies[i] = kw
        return i


W = None  # the world of the env under construction (one env at a time)
W_plank_boxes = [None]


def install_pybullet():
    pb = types.ModuleType("pybullet")
    pb.error = RuntimeError
    for k, v in dict(DIRECT=2, GUI=1, SHARED_MEMORY=3, POSITION_CONTROL=2, VELOCITY_CONTROL=0, TORQUE_CONTROL=1,
                     JOINT_REVOLUTE=0, JOINT_PRISMATIC=1, JOINT_FIXED=4, MJCF_COLORS_FROM_FILE=512,
                     URDF_USE_SELF_COLLISION=8, URDF_USE_SELF_COLLISION_EXCLUDE_ALL_PARENTS=32,
                     URDF_USE_SELF_COLLISION_EXCLUDE_PARENT=16, URDF_USE_INERTIA_FROM_FILE=2,
                     COV_ENABLE_RENDERING=7, COV_ENABLE_GUI=1, COV_ENABLE_KEYBOARD_SHORTCUTS=9,
                     COV_ENABLE_SEGMENTATION_MARK_PREVIEW=5, COV_ENABLE_DEPTH_BUFFER_PREVIEW=4,
                     COV_ENABLE_RGB_BUFFER_PREVIEW=3, GEOM_SPHERE=2, GEOM_MESH=5).items():
        setattr(pb, k, v)

    def connect(mode, options=""):
        global W
        W = FakeWorld()
        return 0

    pb.connect = connect
    pb.disconnect = lambda **kw: None
    pb.configureDebugVisualizer = lambda *a, **kw: None
    pb.setGravity = lambda x, y, z: setattr(W.params, "gravity", -z)
    pb.setDefaultContactERP = lambda erp: setattr(W.params, "erp_contact", erp)

    def setPhysicsEngineParameter(fixedTimeStep=None, numSolverIterations=None, numSubSteps=None, **kw):
        W.params.dt = fixedTimeStep / numSubSteps
        W.params.substeps = numSubSteps
        W.params.iterations = numSolverIterations

    pb.setPhysicsEngineParameter = setPhysicsEngineParameter
    def saveState():
        r = W.robot
        if r is not None:
            W.saved[0] = (O.state_vector(r["state"], r["A"]).copy(), W.tau.copy(), W.base_world.copy())
        return 0

    def restoreState(stateId=0, **kw):
        r = W.robot
        sv, tau, bw = W.saved[0]
        W.base_world = bw.copy()
        A = r["A"]
        r["state"] = O.make_state(A, sv[0:3], sv[3:7], sv[7:10], sv[10:13], sv[13:13 + A], sv[13 + A:13 + 2 * A])
        W.tau[:] = tau
        W.warm = (C.c_double * O.MAXW)()
        W.contacts = None

    pb.saveState = saveState
    pb.restoreState = restoreState
    pb.JOINT_POINT2POINT = 5

    def createConstraint(parentBody, parentLink, childBody, childLink, jointType=None, jointAxis=None,
                         parentFramePosition=None, childFramePosition=None, **kw):
        """Cassie's loop closures (env_cassie.py:114-137): must be the ones compiled into the table."""
        assert jointType == pb.JOINT_POINT2POINT
        c = [q for q in W.robot["table"]["p2p"] if q["link_a"] == parentLink and q["link_b"] == childLink]
        assert len(c) == 1 and np.allclose(c[0]["pivot_a"], parentFramePosition, atol=0) \
            and np.allclose(c[0]["pivot_b"], childFramePosition, atol=0)
        return 0

    pb.createConstraint = createConstraint
    pb.setCollisionFilterGroupMask = lambda *a, **kw: None

    def loadSDF(filename):
        assert filename.endswith("plane_stadium.sdf")
        W.params.has_ground = 1
        return (W.new_body(kind="ground"),)

    pb.loadSDF = loadSDF

    def load_robot_table(name, flags=0, base_pos=None):
        t = load_table(os.path.join(MODELS, name + ".json"))
        # Planar tables (walker2d / crab2d) fold the root's ignorex / ignorez / ignorey joints into a free base.  Towards
        # the env layer the stand-in presents them the way Bullet does: joints 0..2 on a fixed base, the pelvis as link 2,
        # every table link shifted by 3; the base frame sits 1.35 above the pelvis origin (body pos "0 0 -1.35").
        W.shift = 3 if t.get("planar") else 0
        W.base_world = np.zeros(3)
        W.params.self_collision = 1 if flags & pb.URDF_USE_SELF_COLLISION else 0
        A = t["n_dof"]
        W.robot = dict(table=t, model=O.model_from_table(t), A=A,
                       state=O.make_state(A, ([0, 0, -1.35] if t.get("planar") else t["base"]["init_pos"])
                                          if base_pos is None else base_pos, [0, 0, 0, 1],
                                          [0] * 3, [0] * 3, [0] * A, [0] * A))
        W.tau = np.zeros(A)
        W.warm = (C.c_double * O.MAXW)()
        W.contacts = None
        return W.new_body(kind="robot")

    def loadMJCF(path, flags=0):
        return (load_robot_table(os.path.splitext(os.path.basename(path))[0], flags),)

    pb.loadMJCF = loadMJCF

    def quat_to_mat(q):
        x, y, z, w = [float(v) for v in q]
        return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                         [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                         [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])

    def loadURDF(filename, basePosition=(0, 0, 0), baseOrientation=(0, 0, 0, 1), useFixedBase=False,
                 globalScaling=1.0, **kw):
        """A stepping stone (data/objects/steps/*.urdf): static links with one box (or cylinder) collision shape each,
        joined by fixed joints at the origin.  Geometry is read from the URDF, nothing is taken from the oracle."""
        import xml.etree.ElementTree as ET

        if "cassie" in os.path.basename(filename):
            return load_robot_table("cassie", flags=kw.get("flags", 0), base_pos=basePosition)
        root = ET.parse(filename).getroot()
        links = []
        for l in root.findall("link"):
            f3 = lambda e: np.array([float(v) for v in e.get("xyz", "0 0 0").split()])
            iner = f3(l.find("inertial").find("origin")) * globalScaling
            col = l.find("collision")
            corg = f3(col.find("origin")) * globalScaling
            geo = col.find("geometry")
            if geo.find("box") is not None:
                half = 0.5 * globalScaling * np.array([float(v) for v in geo.find("box").get("size").split()])
                cyl = 0
            else:
                c = geo.find("cylinder")
                r, h = float(c.get("radius")) * globalScaling, float(c.get("length")) * globalScaling
                half, cyl = np.array([r, r, 0.5 * h]), 1
            assert float(l.find("inertial").find("mass").get("value")) == 0  # static
            links.append(dict(inertial=iner, col=corg, half=half, cyl=cyl))
        for j in root.findall("joint"):
            assert j.get("type") == "fixed" and j.find("origin").get("xyz") == "0 0 0"
        R = quat_to_mat(baseOrientation)
        b = W.new_body(kind="plank", links=links, com=np.asarray(basePosition, dtype=np.float64) + R @ links[0]["inertial"],
                       quat=np.asarray(baseOrientation, dtype=np.float64), dyn={})
        W.planks.append(b)
        return b

    pb.loadURDF = loadURDF

    def getQuaternionFromEuler(e):  # Bullet: yaw-pitch-roll (ZYX) composition
        roll, pitch, yaw = [float(v) for v in e]
        cr, sr, cp, sp = np.cos(roll * 0.5), np.sin(roll * 0.5), np.cos(pitch * 0.5), np.sin(pitch * 0.5)
        cy, sy = np.cos(yaw * 0.5), np.sin(yaw * 0.5)
        return (sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy,
                cr * cp * cy + sr * sp * sy)

    pb.getQuaternionFromEuler = getQuaternionFromEuler
    pb.GEOM_CYLINDER = 4

    def createCollisionShape(shapeType, radius=None, height=None, **kw):
        assert shapeType == pb.GEOM_CYLINDER
        W.shapes.append(dict(radius=float(radius), height=float(height)))
        return len(W.shapes) - 1

    pb.createCollisionShape = createCollisionShape
    pb.createVisualShape = lambda *a, **kw: -1

    def createMultiBody(baseMass=0.0, baseCollisionShapeIndex=-1, baseVisualShapeIndex=-1, basePosition=(0, 0, 0), **kw):
        """MonkeyBar (bullet_objects.py:148-187): a static cylinder about its local z axis; default dynamics
        (lateral friction 0.5: the changeDynamics call in the reference is commented out)."""
        assert baseMass == 0.0
        sh = W.shapes[baseCollisionShapeIndex]
        b = W.new_body(kind="bar", radius=sh["radius"], height=sh["height"], friction=0.5,
                       com=np.asarray(basePosition, dtype=np.float64).copy(), quat=np.array([0.0, 0, 0, 1]))
        W.bars.append(b)
        return b

    pb.createMultiBody = createMultiBody
    pb.getNumConstraints = lambda: 0
    pb.getConstraintUniqueId = lambda i: i
    pb.removeConstraint = lambda i: None

    def bar_array():
        out = []
        for bi, b in enumerate(W.bars):
            body = W.bodies[b]
            bar = O.Bar()
            ax = quat_to_mat(body["quat"])[:, 2]
            for i in range(3):
                bar.center[i] = body["com"][i]
                bar.axis[i] = ax[i]
            bar.halflen, bar.radius, bar.friction, bar.id = 0.5 * body["height"], body["radius"], body["friction"], 20 + bi
            out.append(bar)
        return (O.Bar * len(out))(*out)

    W_plank_boxes.append(bar_array)

    def plank_boxes():
        """The stones as the oracle's static obstacles: link box centre = base COM + R (collision origin - base
        inertial origin); soft contact from changeDynamics (bullet_objects.py:64-72)."""
        if not W.planks:
            return None
        out = []
        for pi, b in enumerate(W.planks):
            body = W.bodies[b]
            R = quat_to_mat(body["quat"])
            for k, l in enumerate(body["links"]):
                bx = O.Box()
                c = body["com"] + R @ (l["col"] - body["links"][0]["inertial"])
                for i in range(3):
                    bx.center[i] = c[i]
                    bx.half[i] = l["half"][i]
                    for j in range(3):
                        bx.R[i][j] = R[i, j]
                dyn = body["dyn"][k - 1]
                bx.friction = dyn["lateralFriction"]
                bx.stiffness = dyn["contactStiffness"]
                bx.damping = dyn["contactDamping"] + 0.1  # Bullet adds the other body's contact damping (link default 0.1)
                bx.id = 10 + 2 * pi + k
                bx.cylinder = l["cyl"]
                out.append(bx)
        return (O.Box * len(out))(*out)

    W_plank_boxes[0] = plank_boxes

    def changeDynamics(body, link, **kw):
        b = W.bodies[body]
        if b["kind"] == "ground":  # bullet_utils.py:371: lateralFriction 0.8, restitution 0.5
            assert kw == dict(lateralFriction=0.8, restitution=0.5)
            W.params.ground_friction = kw["lateralFriction"]
        elif b["kind"] == "plank":
            b["dyn"][link] = dict(kw)
        elif b["kind"] == "robot" and set(kw) == {"jointDamping"}:
            # Cassie: changeDynamics(jointDamping=...) per ordered joint (env_cassie.py:197-201) -- compiled into the table
            assert W.robot["table"]["damping"][W.robot["table"]["dof_of_link"][link]] == kw["jointDamping"]
        elif b["kind"] == "robot" and set(kw) == {"mass"}:
            # Mike: changeDynamics(waist, mass=8) (robots.py:506-510) -- the compiled table already carries it
            assert abs(W.robot["table"]["mass"][link] - kw["mass"]) < 1e-12
        else:
            raise NotImplementedError(kw)

    pb.changeDynamics = changeDynamics
    pb.getNumJoints = lambda body: (W.robot["table"]["n_links"] + W.shift if W.bodies[body]["kind"] == "robot"
                                    else len(W.bodies[body]["links"]) - 1)

    def getJointInfo(body, j):
        t = W.robot["table"]
        if j < W.shift:  # the planar root joints: slide x, slide z, hinge y, unlimited (lower > upper)
            nm = ("ignorex", "ignorez", "ignorey")[j]
            ln = ("link_dummy_pelvis_0", "link_dummy_pelvis_1", t["base"]["name"])[j]
            return (j, nm.encode(), 1 if j < 2 else 0, 0, 0, 0, 0.0, 0.0, 1.0, -1.0, 0.0, 0.0, ln.encode(),
                    ((1, 0, 0), (0, 0, 1), (0, 1, 0))[j], (0, 0, 0), (0, 0, 0, 1), j - 1)
        j -= W.shift
        rev = t["joint_type"][j] == MC_REVOLUTE
        d = t["dof_of_link"][j]
        lo, hi = (t["lower"][d], t["upper"][d]) if rev else (0.0, -1.0)
        info = [j, t["joint_names_all"][j].encode(), JOINT_REVOLUTE if rev else JOINT_FIXED, 0, 0, 0, 0.0, 0.0, lo, hi,
                0.0, 0.0, t["link_names"][j].encode(), tuple(t["axis"][j]), (0, 0, 0), (0, 0, 0, 1), t["parent"][j]]
        return tuple(info)

    pb.getJointInfo = getJointInfo

    def _dof(j):
        return W.robot["table"]["dof_of_link"][j - W.shift] if j >= W.shift else -1

    def getJointState(body, j):
        s, d = W.robot["state"], _dof(j)
        return (s.q[d], s.qd[d], (0,) * 6, 0.0) if d >= 0 else (0.0, 0.0, (0,) * 6, 0.0)

    pb.getJointState = getJointState
    pb.getJointStates = lambda body, js: [getJointState(body, j) for j in js]

    def resetJointState(body, j, targetValue=0.0, targetVelocity=0.0):
        s, d = W.robot["state"], _dof(j)
        if d >= 0:
            s.q[d], s.qd[d] = float(targetValue), float(targetVelocity)

    pb.resetJointState = resetJointState

    def setJointMotorControl2(*a, **kw):
        args = dict(zip(("bodyIndex", "jointIndex", "controlMode"), a))
        args.update(kw)
        d = _dof(args["jointIndex"])
        if args["controlMode"] == pb.TORQUE_CONTROL:  # Cassie applies its PD torques joint by joint
            W.tau[d] = float(args["force"])
        else:  # POSITION_CONTROL with zero force: the default motor switched off
            assert args.get("force", 0) == 0
            if d >= 0:
                W.tau[d] = 0.0

    pb.setJointMotorControl2 = setJointMotorControl2

    def setJointMotorControlArray(bodyIndex, jointIndices, controlMode, forces=None, **kw):
        if controlMode == pb.TORQUE_CONTROL:
            for j, f in zip(jointIndices, forces):
                W.tau[_dof(j)] = float(f)
        else:  # POSITION_CONTROL with zero force: motors off
            assert all(f == 0 for f in forces)
            for j in jointIndices:
                W.tau[_dof(j)] = 0.0

    pb.setJointMotorControlArray = setJointMotorControlArray

    def getBasePositionAndOrientation(body):
        if W.bodies[body]["kind"] == "robot" and W.shift:
            return tuple(W.base_world), (0.0, 0.0, 0.0, 1.0)
        if W.bodies[body]["kind"] == "plank":
            return tuple(W.bodies[body]["com"]), tuple(W.bodies[body]["quat"])
        s = W.robot["state"]
        return tuple(s.pos[:]), tuple(s.quat[:])

    pb.getBasePositionAndOrientation = getBasePositionAndOrientation
    pb.getBaseVelocity = lambda body: (tuple(W.robot["state"].vel[:]), tuple(W.robot["state"].omega[:]))

    def resetBasePositionAndOrientation(body, posObj=None, ornObj=None):
        pos, orn = posObj, ornObj
        if W.bodies[body]["kind"] in ("plank", "bar"):  # sets the pose of the base link's INERTIAL frame
            W.bodies[body]["com"] = np.asarray(pos, dtype=np.float64).copy()
            W.bodies[body]["quat"] = np.asarray(orn, dtype=np.float64).copy()
            return
        assert W.bodies[body]["kind"] == "robot"
        s = W.robot["state"]
        if W.shift:  # fixed base of a planar robot: the whole chain moves with it, the root joints keep their values
            assert tuple(orn) == (0, 0, 0, 1)
            delta = np.asarray(pos, dtype=np.float64) - W.base_world
            W.base_world = np.asarray(pos, dtype=np.float64).copy()
            for k in range(3):
                s.pos[k] += delta[k]
            W.warm = (C.c_double * O.MAXW)()
            W.contacts = None
            return
        for k in range(3):
            s.pos[k] = float(pos[k])
        for k in range(4):
            s.quat[k] = float(orn[k])
        W.warm = (C.c_double * O.MAXW)()
        W.contacts = None

    pb.resetBasePositionAndOrientation = resetBasePositionAndOrientation

    def resetBaseVelocity(body, lin, ang):
        if W.shift:  # a fixed base has no velocity to set
            return
        s = W.robot["state"]
        for k in range(3):
            s.vel[k], s.omega[k] = float(lin[k]), float(ang[k])

    pb.resetBaseVelocity = resetBaseVelocity

    def getLinkState(body, link, computeLinkVelocity=0):
        s = W.robot["state"]
        if W.shift and link == W.shift - 1:  # the pelvis of a planar robot = the table's base
            out = (tuple(s.pos[:]), tuple(s.quat[:]), (0,) * 3, (0, 0, 0, 1), tuple(s.pos[:]), tuple(s.quat[:]))
            return out + (tuple(s.vel[:]), tuple(s.omega[:])) if computeLinkVelocity else out
        assert not computeLinkVelocity
        link -= W.shift
        pos, rot = O.fk(W.robot["model"], s)  # index 0 = base, link i -> i + 1: COM (inertial) frames
        q = mat_to_quat(rot[link + 1])
        return tuple(pos[link + 1]), tuple(q), (0,) * 3, (0, 0, 0, 1), tuple(pos[link + 1]), tuple(q)

    pb.getLinkState = getLinkState

    def stepSimulation():
        r = W.robot
        if W.bars:
            W.contacts, _ = O.step_physics_bars(r["model"], W.params, r["state"], W.tau, W_plank_boxes[1]())
        else:
            W.contacts, _ = O.step_physics(r["model"], W.params, r["state"], W.tau, boxes=W_plank_boxes[0](), warm=W.warm)

    pb.stepSimulation = stepSimulation

    def getContactPoints(bodyA=None, bodyB=None, linkIndexA=None, **kw):
        out = []
        c = W.contacts
        if c is None:
            return out
        ground = [i for i, b in W.bodies.items() if b["kind"] == "ground"]
        sh = W.shift
        if linkIndexA is not None:
            linkIndexA -= sh  # env-layer link index -> table link index (the pelvis of a planar robot becomes -1)
        for k in range(c.n):
            if c.partner[k] >= 1000:  # self-contact: both bodies are the robot; Bullet reports it for either link
                if linkIndexA is None or c.link[k] == linkIndexA:
                    out.append((0, bodyA, bodyA, c.link[k] + sh, c.link_b[k] + sh))
                elif c.link_b[k] == linkIndexA:
                    out.append((0, bodyA, bodyA, c.link_b[k] + sh, c.link[k] + sh))
                continue
            if linkIndexA is not None and c.link[k] != linkIndexA:
                continue
            if c.partner[k] == 0:  # the ground plane
                out.append((0, bodyA, ground[0], c.link[k] + sh, -1))
            elif 10 <= c.partner[k] < 20:  # a stepping stone: box id 10 + 2 * plank + (0 base | 1 cover)
                pi, kk = divmod(c.partner[k] - 10, 2)
                out.append((0, bodyA, W.planks[pi], c.link[k], kk - 1))
            elif 20 <= c.partner[k] < 30:  # a monkey bar
                out.append((0, bodyA, W.bars[c.partner[k] - 20], c.link[k], -1))
        return out

    pb.getContactPoints = getContactPoints

    def getEulerFromQuaternion(qin):  # Bullet's b3GetEulerFromQuaternion semantics (restated as in the oracle)
        q = np.asarray(qin, dtype=np.float64)
        q = q / np.sqrt((q * q).sum())
        sqx, sqy, sqz, squ = q[0] * q[0], q[1] * q[1], q[2] * q[2], q[3] * q[3]
        sarg = -2 * (q[0] * q[2] - q[3] * q[1])
        if sarg <= -0.99999:
            return (0.0, -0.5 * np.pi, 2 * np.arctan2(q[0], -q[1]))
        if sarg >= 0.99999:
            return (0.0, 0.5 * np.pi, 2 * np.arctan2(-q[0], q[1]))
        return (np.arctan2(2 * (q[1] * q[2] + q[3] * q[0]), squ - sqx - sqy + sqz), np.arcsin(sarg),
                np.arctan2(2 * (q[0] * q[1] + q[3] * q[2]), squ + sqx - sqy - sqz))

    pb.getEulerFromQuaternion = getEulerFromQuaternion

    # ---- calls only tools/gen_pybullet_golden.py --standin makes (the env layer never does): model dump, dynamics
    # queries and the full 10-field contact tuples, all answered from the oracle
    pb.getAPIVersion = lambda: "standin"

    def getDynamicsInfo(body, link):
        t = W.robot["table"]
        if link < 0:
            b = t["base"]
            return (b["mass"], 1.0, tuple(b.get("inertia_diag", (0, 0, 0))), (0, 0, 0), (0, 0, 0, 1), 0.0, 0.0, 0.0, -1.0, -1.0)
        l = link - W.shift
        return (t["mass"][l], 1.0, tuple(t["inertia_diag"][l]) if "inertia_diag" in t else (0, 0, 0), (0, 0, 0),
                (0, 0, 0, 1), 0.0, 0.0, 0.0, -1.0, -1.0)

    pb.getDynamicsInfo = getDynamicsInfo
    pb.getCollisionShapeData = lambda body, link: tuple(
        (body, link, g["type"], tuple(g["size"])) for g in W.robot["table"]["geoms"] if g["link"] == link - W.shift)

    def calculateMassMatrix(body, q):
        r = W.robot
        A = r["table"]["n_dof"]
        s = O.make_state(A, list(r["state"].pos), list(r["state"].quat), [0] * 3, [0] * 3, list(q), [0.0] * A)
        return O.mass_matrix(r["model"], s).tolist()

    pb.calculateMassMatrix = calculateMassMatrix

    def calculateInverseDynamics(body, q, qd, acc):
        r = W.robot
        A = r["table"]["n_dof"]
        st = r["state"]
        s = O.make_state(A, list(st.pos), list(st.quat), list(st.omega), list(st.vel), list(q), list(qd))
        return O.rnea(r["model"], s, np.concatenate([np.zeros(6), acc]), W.params.gravity)[6:].tolist()

    pb.calculateInverseDynamics = calculateInverseDynamics
    _five = pb.getContactPoints

    def getContactPointsFull(bodyA=None, bodyB=None, linkIndexA=None, **kw):
        """the env layer's 5-field tuples extended by positionOnA, positionOnB, contactNormalOnB, distance, normalForce"""
        c = W.contacts
        short = _five(bodyA=bodyA, bodyB=bodyB, linkIndexA=linkIndexA, **kw)
        if c is None or linkIndexA is not None:
            return short
        out = []
        for k, tup in zip([k for k in range(c.n)], short):
            pa = tuple(c.pos_a[k])
            n = tuple(c.normal[k])
            pbp = tuple(pa[i] - c.dist[k] * n[i] for i in range(3))
            out.append(tuple(tup) + (pa, pbp, n, c.dist[k], c.impulse[k] / (W.params.dt)))
        return out

    pb.getContactPoints = getContactPointsFull
    sys.modules["pybullet"] = pb
    return pb


# ----------------------------------------------------------------------------------------------- traces
def trace_walker3d_custom(seed, steps, action_seed, eval_mode=False, env_name="Walker3DCustomEnv", hold_at_target=False):
    """hold_at_target: before every step the walker is put on its feet AT the walk target (recorded state override
    through the reference's own robot calls), so that close_count runs up to stop_frames and the mid-episode target
    re-randomisation (env_locomotion.py:196-222) is exercised, which no random policy reaches."""
    import mocca_envs.env_locomotion as EL

    env = getattr(EL, env_name)()
    env.seed(seed)  # the quirk-Q1 path: np_random rebound, the robot keeps the stream it got at construction
    if eval_mode:
        env.evaluation_mode()
    rs = np.random.RandomState(action_seed)
    A = env.action_space.shape[0]
    obs = [env.reset()]
    acts, rews, dones, targets, resets, teleports = [], [], [], [], [], []
    for t in range(steps):
        a = rs.uniform(-1.2, 1.2, A)
        if hold_at_target:
            tgt = np.array([env.walk_target[0], env.walk_target[1], 1.33])
            env.robot.reset_joint_states(env.robot.base_joint_angles, env.robot.base_joint_speeds)
            env.robot.robot_body.reset_pose([float(v) for v in tgt], [0.0, 0.0, 0.0, 1.0])
            env.robot.robot_body.reset_velocity([0.0, 0.0, 0.0], [0.0, 0.0, 0.0])
            teleports.append([t, *tgt])
            a = a * 0.05
        o, r, d, info = env.step(a)
        acts.append(a); rews.append(r); dones.append(d); targets.append(np.array(env.walk_target, dtype=np.float64))
        if d:
            resets.append(t)
            obs.append(o)          # terminal observation
            o = env.reset()
        obs.append(o)
    return dict(seed=seed, action_seed=action_seed, eval_mode=int(eval_mode), actions=np.array(acts),
                obs=np.array(obs, dtype=np.float64), rewards=np.array(rews), dones=np.array(dones),
                walk_target=np.array(targets), resets=np.array(resets, dtype=np.int64),
                teleports=np.array(teleports, dtype=np.float64).reshape(-1, 4),
                mirror=np.concatenate([np.asarray(x, dtype=np.int64).ravel() for x in env.get_mirror_indices()]),
                construction_seed=CONSTRUCTION_SEED)


def trace_monkey(seed, steps, action_seed, grab_every=0):
    """grab_every > 0: every so many steps the whole monkey is translated (velocities zeroed, pose kept; through the
    reference's own robot calls) so that its swing palm sits on the target bar -- the trace then walks the palm-contact /
    swing-pivot swap / bar-recycling logic that random torques never reach.  The new base positions are recorded."""
    from mocca_envs.env_locomotion import Monkey3DCustomEnv

    env = Monkey3DCustomEnv()
    env.seed(seed)
    rs = np.random.RandomState(action_seed)
    A = env.action_space.shape[0]
    obs = [env.reset()]
    terrain = [env.terrain_info.copy()]
    acts, rews, dones, nexts, resets, teleports = [], [], [], [], [], []
    for t in range(steps):
        a = rs.uniform(-1.0, 1.0, A)
        if grab_every and t > 0 and t % grab_every == 0:
            palm = env.robot.parts["right_palm" if env.swing_leg == 0 else "left_palm"].pose().xyz()
            new = np.array(env.robot.robot_body.pose().xyz()) + (env.terrain_info[env.next_step_index, 0:3] - palm)
            js = env._p.getJointStates(env.robot.id, env.robot.ordered_joint_ids)
            env.robot.reset_joint_states([x[0] for x in js], [0.0 for _ in js])
            env.robot.robot_body.reset_pose([float(v) for v in new], env.robot.robot_body.pose().orientation())
            env.robot.robot_body.reset_velocity([0.0, 0.0, 0.0], [0.0, 0.0, 0.0])
            teleports.append([t, *new])
            a = a * 0.2
        sent = a.copy()
        o, r, d, info = env.step(a)  # overwrites the two finger entries of `a` in place (quirk Q11)
        acts.append(sent); rews.append(r); dones.append(d); nexts.append(env.next_step_index)
        if d:
            resets.append(t)
            obs.append(o)
            o = env.reset()
            terrain.append(env.terrain_info.copy())
        obs.append(o)
    return dict(seed=seed, action_seed=action_seed, actions=np.array(acts), obs=np.array(obs, dtype=np.float64),
                rewards=np.array(rews, dtype=np.float64), dones=np.array(dones), next_step_index=np.array(nexts),
                terrain=np.array(terrain), resets=np.array(resets, dtype=np.int64), construction_seed=CONSTRUCTION_SEED,
                teleports=np.array(teleports, dtype=np.float64).reshape(-1, 4))


def make_cassie_env():
    """CassieEnv-v0.  env_cassie.py does not import as shipped; the three defects are fixed by intent, nothing else:
    Q7 the missing `.loadstep` module (only the mocap variants use it), Q8 BodyPart / Joint not imported, Q9 the
    EnvBase.__init__ call that passes `render` as robot_kwargs and drops `power`."""
    import mocca_envs  # noqa: F401
    import gym
    from mocca_envs import bullet_utils
    from mocca_envs.env_base import EnvBase

    sys.modules["mocca_envs.loadstep"] = types.SimpleNamespace(CassieTrajectory=None)
    from mocca_envs import env_cassie as EC

    EC.BodyPart, EC.Joint = bullet_utils.BodyPart, bullet_utils.Joint

    class CassieEnv(EC.CassieEnv):
        def __init__(self, render=False, planar=False, power_coef=1.0, residual_control=True, rsi=True):
            self.planar, self.residual_control, self.rsi = planar, residual_control, rsi
            EnvBase.__init__(self, EC.Cassie, robot_kwargs={"power": power_coef}, render=render)
            high = np.inf * np.ones(self.robot.observation_space.shape[0] + 2)
            self.observation_space = gym.spaces.Box(-high, high, dtype=np.float32)
            self.action_space = self.robot.action_space

    return CassieEnv()


def trace_cassie(steps, action_seed):
    """CassieEnv-v0 through make_cassie_env()."""
    env = make_cassie_env()
    rs = np.random.RandomState(action_seed)
    obs = [env.reset()]
    acts, rews, dones, alive, prog, resets = [], [], [], [], [], []
    for t in range(steps):
        a = rs.uniform(-1, 1, 10) * (0.1 if t % 40 < 25 else 0.6)  # mostly the standing residual, bursts that topple it
        o, r, d, info = env.step(a)
        acts.append(a); rews.append(r); dones.append(d); alive.append(info["AliveRew"]); prog.append(info["ProgressRew"])
        if d:
            resets.append(t)
            obs.append(o)
            o = env.reset()
        obs.append(o)
    return dict(action_seed=action_seed, actions=np.array(acts), obs=np.array(obs, dtype=np.float64),
                rewards=np.array(rews, dtype=np.float64), dones=np.array(dones), alive=np.array(alive),
                progress=np.array(prog), resets=np.array(resets, dtype=np.int64))


def trace_stepper(env_name, seed, steps, action_seed, curriculum, teleport_every=0, **kwargs):
    """teleport_every > 0: every so many steps the walker is put back on its feet above the NEXT stepping stone (base
    pose, zero velocity, through the env's own robot / Bullet-client calls), so that the trace walks the target-advance,
    plank-recycling, step-bonus and stop-on-step logic that a random policy never reaches.  The teleports are recorded;
    the replay applies the same state override to the oracle."""
    import mocca_envs.env_locomotion as EL

    env = getattr(EL, env_name)(**kwargs)
    env.seed(seed)
    env.set_env_params({"curriculum": curriculum})
    rs = np.random.RandomState(action_seed)
    A = env.action_space.shape[0]
    obs = [env.reset()]
    terrain = [env.terrain_info.copy()]
    acts, rews, dones, nexts, resets, reached, teleports = [], [], [], [], [], [], []
    for t in range(steps):
        a = rs.uniform(-1.0, 1.0, A) * 0.6   # gentler than the flat-ground traces: walkers stay on the stones longer
        if teleport_every and t > 0 and t % teleport_every == 0:
            tgt = env.terrain_info[env.next_step_index, 0:3] + np.array([0.0, 0.0, 1.36])
            env.robot.reset_joint_states(env.robot.base_joint_angles, env.robot.base_joint_speeds)
            env.robot.robot_body.reset_pose([float(v) for v in tgt], [0.0, 0.0, 0.0, 1.0])
            env.robot.robot_body.reset_velocity([0.0, 0.0, 0.0], [0.0, 0.0, 0.0])
            teleports.append([t, *tgt])
            a = a * 0.1
        o, r, d, info = env.step(a)
        acts.append(a); rews.append(r); dones.append(d); nexts.append(env.next_step_index)
        reached.append(info.get("steps_reached", -1))
        if d:
            resets.append(t)
            obs.append(o)
            o = env.reset()
            terrain.append(env.terrain_info.copy())
        obs.append(o)
    return dict(seed=seed, action_seed=action_seed, curriculum=curriculum, actions=np.array(acts),
                obs=np.array(obs, dtype=np.float64), rewards=np.array(rews, dtype=np.float64), dones=np.array(dones),
                next_step_index=np.array(nexts), steps_reached=np.array(reached), terrain=np.array(terrain),
                resets=np.array(resets, dtype=np.int64), construction_seed=CONSTRUCTION_SEED,
                teleports=np.array(teleports, dtype=np.float64).reshape(-1, 4),
                mirror=np.concatenate([np.asarray(x, dtype=np.int64).ravel() for x in env.get_mirror_indices()]),
                plank_class=str(kwargs.get("plank_class") or "LargePlank"), random_reward=int(kwargs.get("random_reward", False)))


CONSTRUCTION_SEED = 12345  # EnvBase.__init__ calls self.seed() with no argument: fixed here instead of os.urandom


def main(out=None):
    assert os.path.isdir(REF), "the reference tree is only present in the build container"
    install_gym()
    install_pybullet()
    sys.path.insert(0, REF)
    import gym.utils.seeding as seeding

    real = seeding.np_random
    seeding.np_random = lambda seed=None: real(CONSTRUCTION_SEED if seed is None else seed)
    out = out or os.path.join(ROOT, "tests", "golden")
    os.makedirs(out, exist_ok=True)
    for seed, steps, aseed, ev in ((0, 160, 1, False), (7, 160, 2, False), (3, 60, 3, True)):
        g = trace_walker3d_custom(seed, steps, aseed, ev)
        fn = os.path.join(out, "ref_walker3d_custom_seed%d%s.npz" % (seed, "_eval" if ev else ""))
        np.savez_compressed(fn, **g)
        print("wrote %s: %d steps, %d episodes ended, reward sum %.6f" % (fn, steps, len(g["resets"]),
                                                                          g["rewards"].sum()))
    g = trace_walker3d_custom(11, 200, 14, False, hold_at_target=True)
    fn = os.path.join(out, "ref_walker3d_custom_seed11_target.npz")
    np.savez_compressed(fn, **g)
    print("wrote %s: %d episodes ended, %d distinct walk targets, reward sum %.6f"
          % (fn, len(g["resets"]), len(np.unique(g["walk_target"][:, 0])), g["rewards"].sum()))
    g = trace_walker3d_custom(1, 120, 4, False, env_name="Child3DCustomEnv")
    fn = os.path.join(out, "ref_child3d_custom_seed1.npz")
    np.savez_compressed(fn, **g)
    print("wrote %s: %d episodes ended, reward sum %.6f" % (fn, len(g["resets"]), g["rewards"].sum()))
    for env_name, tag, seed, aseed in (("Walker2DCustomEnv", "walker2d", 2, 31), ("Crab2DCustomEnv", "crab2d", 4, 32)):
        g = trace_walker3d_custom(seed, 200, aseed, False, env_name=env_name)
        fn = os.path.join(out, "ref_%s_custom_seed%d.npz" % (tag, seed))
        np.savez_compressed(fn, **g)
        print("wrote %s: %d episodes ended (the reference forces done = False), reward sum %.6f, feet contacts %d"
              % (fn, len(g["resets"]), g["rewards"].sum(), int(g["obs"][:, -4:-2].sum())))
    g = trace_cassie(60, 21)
    fn = os.path.join(out, "ref_cassie_a21.npz")
    np.savez_compressed(fn, **g)
    print("wrote %s: %d env steps (x50 substeps), %d episodes ended, reward sum %.6f"
          % (fn, len(g["actions"]), len(g["resets"]), g["rewards"].sum()))
    for seed, steps, aseed, grab in ((0, 260, 11, 0), (5, 260, 12, 0), (8, 200, 15, 10)):
        g = trace_monkey(seed, steps, aseed, grab)
        fn = os.path.join(out, "ref_monkey3d_custom_seed%d%s.npz" % (seed, "_grab" if grab else ""))
        np.savez_compressed(fn, **g)
        print("wrote %s: %d steps, %d episodes ended, max next_step_index %d, reward sum %.6f"
              % (fn, steps, len(g["resets"]), g["next_step_index"].max(), g["rewards"].sum()))
    for name, tag, seed, steps, aseed, cur, kw in (
            ("Walker3DStepperEnv", "walker3d_stepper_c0", 0, 200, 5, 0, {}),
            ("Walker3DStepperEnv", "walker3d_stepper_c9", 4, 200, 6, 9, {}),
            ("Walker3DStepperEnv", "walker3d_stepper_c5_plank", 2, 150, 7, 5, {"plank_class": "Plank"}),
            ("Walker3DStepperEnv", "walker3d_stepper_c7_pillar", 6, 150, 9, 7, {"plank_class": "Pillar"}),
            ("Walker3DStepperEnv", "walker3d_stepper_c3_rr", 5, 150, 8, 3, {"random_reward": True}),
            ("MikeStepperEnv", "mike_stepper_c4", 8, 150, 10, 4, {}),
            ("Walker3DStepperEnv", "walker3d_stepper_c6_walk", 9, 420, 13, 6, {"teleport_every": 14})):
        g = trace_stepper(name, seed, steps, aseed, cur, **kw)
        fn = os.path.join(out, "ref_%s.npz" % tag)
        np.savez_compressed(fn, **g)
        print("wrote %s: %d steps, %d episodes ended, max next_step_index %d, reward sum %.6f"
              % (fn, steps, len(g["resets"]), g["next_step_index"].max(), g["rewards"].sum()))


if __name__ == "__main__":
    main(*sys.argv[1:2])
