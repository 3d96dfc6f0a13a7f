"""Model compiler: MJCF robot -> flat articulation / collision tables.

Flattens the reference's robot description files (``mocca_envs/data/robots/*.xml``,
loaded by the reference at ``mocca_envs/robots.py:101-105`` via ``loadMJCF``) into the
struct-of-arrays tables that the CUDA kernels and the CPU oracle consume.

The compiler restates the conventions of Bullet's MJCF importer + ``URDF2Bullet`` multibody
conversion (third-party ``pybullet``; un-pinned dependency of the reference, ``setup.py:11``;
conventions listed in SURVEY.md App. B.1):

* top-level body without a free joint  -> floating base
* a body with k hinge joints           -> k links (k-1 massless dummies + the body itself),
  link frames at the joint positions, PyBullet joint index == link index in DFS pre-order
* a body with no joint                 -> link attached by a fixed joint (``jointfix_*``)
* ``inertiafromgeom``                  -> mass = 1000 kg/m^3 * sum(geom volumes),
  inertial frame == body frame (no COM recomputation),
  inertia diagonal = box inertia of the AABB of the link's compound collision shape
* MJCF ``damping`` / ``armature``       -> ignored by the importer (switchable here)
* geom ``friction``'s first entry      -> lateral friction; condim 3 -> no spinning/rolling rows
* ``contype``/``conaffinity``           -> collision filter group/mask of the link

Everything here is host-side, one-time work.  Output is a plain dict of
lists (JSON-serialisable, human-diffable) so that a host with PyBullet can diff it against
``getJointInfo`` / ``getDynamicsInfo`` dumps.
"""
from __future__ import annotations

import json
import math
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

GEOM_SPHERE, GEOM_CAPSULE, GEOM_BOX = 0, 1, 2
JOINT_FIXED, JOINT_REVOLUTE = 0, 1

DENSITY = 1000.0  # Bullet MJCF importer default when no <inertial> is given
CONTACT_BREAKING_THRESHOLD = 0.02  # Bullet gContactBreakingThreshold


# ----------------------------------------------------------------------------- small math
def quat_to_mat(q):
    """xyzw quaternion (normalised on the fly, as btMatrix3x3::setRotation does) -> 3x3."""
    x, y, z, w = [float(v) for v in q]
    d = x * x + y * y + z * z + w * w
    s = 2.0 / d
    xs, ys, zs = x * s, y * s, z * s
    wx, wy, wz = w * xs, w * ys, w * zs
    xx, xy, xz = x * xs, x * ys, x * zs
    yy, yz, zz = y * ys, y * zs, z * zs
    return np.array(
        [
            [1.0 - (yy + zz), xy - wz, xz + wy],
            [xy + wz, 1.0 - (xx + zz), yz - wx],
            [xz - wy, yz + wx, 1.0 - (xx + yy)],
        ]
    )


def mat_to_quat(m):
    """3x3 rotation -> xyzw quaternion (w >= 0)."""
    m = np.asarray(m, dtype=np.float64)
    tr = m[0, 0] + m[1, 1] + m[2, 2]
    if tr > 0:
        s = math.sqrt(tr + 1.0) * 2
        w = 0.25 * s
        x = (m[2, 1] - m[1, 2]) / s
        y = (m[0, 2] - m[2, 0]) / s
        z = (m[1, 0] - m[0, 1]) / s
    elif m[0, 0] > m[1, 1] and m[0, 0] > m[2, 2]:
        s = math.sqrt(1.0 + m[0, 0] - m[1, 1] - m[2, 2]) * 2
        w = (m[2, 1] - m[1, 2]) / s
        x = 0.25 * s
        y = (m[0, 1] + m[1, 0]) / s
        z = (m[0, 2] + m[2, 0]) / s
    elif m[1, 1] > m[2, 2]:
        s = math.sqrt(1.0 + m[1, 1] - m[0, 0] - m[2, 2]) * 2
        w = (m[0, 2] - m[2, 0]) / s
        x = (m[0, 1] + m[1, 0]) / s
        y = 0.25 * s
        z = (m[1, 2] + m[2, 1]) / s
    else:
        s = math.sqrt(1.0 + m[2, 2] - m[0, 0] - m[1, 1]) * 2
        w = (m[1, 0] - m[0, 1]) / s
        x = (m[0, 2] + m[2, 0]) / s
        y = (m[1, 2] + m[2, 1]) / s
        z = 0.25 * s
    q = np.array([x, y, z, w])
    if w < 0:
        q = -q
    return q / np.linalg.norm(q)


def shortest_arc_quat(v0, v1):
    """Bullet's shortestArcQuat(v0, v1): rotation taking unit v0 onto unit v1 (xyzw)."""
    v0 = np.asarray(v0, dtype=np.float64)
    v1 = np.asarray(v1, dtype=np.float64)
    c = np.cross(v0, v1)
    d = float(np.dot(v0, v1))
    if d < -1.0 + 1e-7:
        # opposite vectors: any perpendicular axis (btPlaneSpace1 picks one)
        n = np.array([0.0, -v0[2], v0[1]]) if abs(v0[2]) > 0.7071 else np.array([-v0[1], v0[0], 0.0])
        n /= np.linalg.norm(n)
        return np.array([n[0], n[1], n[2], 0.0])
    s = math.sqrt((1.0 + d) * 2.0)
    rs = 1.0 / s
    return np.array([c[0] * rs, c[1] * rs, c[2] * rs, s * 0.5])


# ----------------------------------------------------------------------------- data classes
@dataclass
class Geom:
    name: str
    type: int
    pos: np.ndarray  # centre in the link's inertial frame
    quat: np.ndarray  # xyzw, orientation in the link's inertial frame
    size: np.ndarray  # sphere [r,0,0]; capsule [r, half_len, 0] (axis = local z); box half extents
    group: int
    mask: int
    friction: float

    def volume(self) -> float:
        if self.type == GEOM_SPHERE:
            r = self.size[0]
            return 4.0 / 3.0 * math.pi * r ** 3
        if self.type == GEOM_CAPSULE:
            r, hl = self.size[0], self.size[1]
            return 4.0 / 3.0 * math.pi * r ** 3 + math.pi * r * r * (2 * hl)
        hx, hy, hz = self.size
        return 8.0 * hx * hy * hz

    def aabb(self):
        """AABB in the link frame, following btSphereShape/btCapsuleShape/btBoxShape::getAabb."""
        R = quat_to_mat(self.quat)
        if self.type == GEOM_SPHERE:
            he = np.array([self.size[0]] * 3)
            ext = he
        elif self.type == GEOM_CAPSULE:
            he = np.array([self.size[0], self.size[0], self.size[0] + self.size[1]])
            ext = np.abs(R) @ he
        else:
            he = np.array(self.size, dtype=np.float64)
            ext = np.abs(R) @ he
        return self.pos - ext, self.pos + ext


@dataclass
class Link:
    name: str
    parent: int  # index into links, -1 = base
    joint_name: str
    joint_type: int
    axis: np.ndarray  # joint axis in this link's frame
    rot_parent_to_this: np.ndarray  # xyzw; Bullet's zeroRotParentToThis
    e_vec: np.ndarray  # parent COM -> pivot, in parent frame
    d_vec: np.ndarray  # pivot -> this COM, in this frame
    mass: float = 0.0
    inertia: np.ndarray = field(default_factory=lambda: np.zeros(3))
    lower: float = 0.0
    upper: float = 0.0
    damping: float = 0.0
    armature: float = 0.0
    geoms: List[Geom] = field(default_factory=list)
    is_dummy: bool = False


def _floats(s: Optional[str], default=None):
    if s is None:
        return default
    return np.array([float(t) for t in s.split()], dtype=np.float64)


class MJCFCompiler:
    """Parse an MJCF humanoid-style file into Bullet-convention links."""

    def __init__(self, path: str, use_mjcf_damping: bool = False, use_mjcf_armature: bool = False):
        self.path = path
        self.use_mjcf_damping = use_mjcf_damping
        self.use_mjcf_armature = use_mjcf_armature
        self.links: List[Link] = []
        self.base: Optional[Link] = None
        self.anon = 0
        self.planar = False

        root = ET.parse(path).getroot()
        comp = root.find("compiler")
        self.degrees = comp is None or comp.get("angle", "degree") == "degree"
        dflt = root.find("default")
        dj = dflt.find("joint") if dflt is not None else None
        dg = dflt.find("geom") if dflt is not None else None
        self.d_limited = (dj.get("limited", "false") == "true") if dj is not None else False
        self.d_damping = float(dj.get("damping", 0.0)) if dj is not None else 0.0
        self.d_armature = float(dj.get("armature", 0.0)) if dj is not None else 0.0
        self.d_contype = int(dg.get("contype", 1)) if dg is not None else 1
        self.d_conaffinity = int(dg.get("conaffinity", 1)) if dg is not None else 1
        fr = _floats(dg.get("friction")) if dg is not None and dg.get("friction") else None
        self.d_friction = float(fr[0]) if fr is not None else 0.5

        world = root.find("worldbody")
        top = world.find("body")
        self.base_name = top.get("name")
        self.base_pos = _floats(top.get("pos"), np.zeros(3))
        self._parse_body(top, parent_link=-1, is_root=True)
        self._finalize_inertia()

    # ---- geoms
    def _parse_geom(self, g, shift: np.ndarray) -> Geom:
        """``shift`` = body-frame origin expressed in the link inertial frame (zero here:
        inertial frame == body frame for MJCF bodies without <inertial>)."""
        gtype = g.get("type", "sphere")
        name = g.get("name", "geom%d" % self.anon)
        group = int(g.get("contype", self.d_contype))
        mask = int(g.get("conaffinity", self.d_conaffinity))
        fr = _floats(g.get("friction"))
        friction = float(fr[0]) if fr is not None else self.d_friction
        size = _floats(g.get("size"))
        if gtype == "sphere":
            pos = _floats(g.get("pos"), np.zeros(3)) + shift
            return Geom(name, GEOM_SPHERE, pos, np.array([0, 0, 0, 1.0]), np.array([size[0], 0, 0]), group, mask, friction)
        if gtype == "capsule":
            ft = _floats(g.get("fromto"))
            assert ft is not None, "only fromto capsules occur in the in-scope models"
            f, t = ft[:3], ft[3:]
            diff = t - f
            h = float(np.linalg.norm(diff))
            quat = shortest_arc_quat([0, 0, 1.0], diff / h) if h > 1e-12 else np.array([0, 0, 0, 1.0])
            return Geom(name, GEOM_CAPSULE, 0.5 * (f + t) + shift, quat, np.array([size[0], 0.5 * h, 0]), group, mask, friction)
        if gtype == "box":
            pos = _floats(g.get("pos"), np.zeros(3)) + shift
            quat = np.array([0, 0, 0, 1.0])
            return Geom(name, GEOM_BOX, pos, quat, np.array(size[:3]), group, mask, friction)
        raise ValueError("unsupported geom type %s" % gtype)

    # ---- bodies
    def _parse_body(self, body, parent_link: int, is_root: bool):
        name = body.get("name")
        if name is None:
            name = "anon_body_%d" % self.anon
            self.anon += 1
        pos = _floats(body.get("pos"), np.zeros(3))
        q = _floats(body.get("quat"))  # MJCF order w x y z
        quat = np.array([q[1], q[2], q[3], q[0]]) if q is not None else np.array([0, 0, 0, 1.0])
        R_body = quat_to_mat(quat)  # body axes in parent-body axes

        joints = [c for c in body if c.tag == "joint"]
        geoms = [c for c in body if c.tag == "geom"]
        children = [c for c in body if c.tag == "body"]

        if is_root:
            # walker2d.xml / crab2d.xml: the root carries slide-x, slide-z and hinge-y joints named "ignore*"
            # (skipped by robots.py:163-172).  Bullet imports them as a fixed base with two massless dummy links,
            # which is the same mechanism as a free base confined to the x-z plane: every joint axis is y and every
            # COM has y = 0, so the out-of-plane coordinates of a free base stay exactly zero (tested) -- the root
            # is compiled as the ordinary floating base and the table is flagged planar.
            assert all((j.get("name") or "").startswith("ignore") for j in joints), \
                "root body with joints is not a floating base"
            self.planar = bool(joints)
            link = Link(name, -2, "", JOINT_FIXED, np.zeros(3), np.array([0, 0, 0, 1.0]), np.zeros(3), np.zeros(3))
            link.geoms = [self._parse_geom(g, np.zeros(3)) for g in geoms]
            self.base = link
            my_index = -1
        else:
            rot_p2t = mat_to_quat(R_body.T)
            if not joints:
                link = Link(name, parent_link, "jointfix_%s" % name, JOINT_FIXED, np.zeros(3), rot_p2t, pos.copy(), np.zeros(3))
                self.links.append(link)
                my_index = len(self.links) - 1
            else:
                prev_pivot = None
                cur_parent = parent_link
                for k, j in enumerate(joints):
                    jpos = _floats(j.get("pos"), np.zeros(3))  # in body frame
                    axis = _floats(j.get("axis"), np.array([1.0, 0, 0]))
                    axis = axis / np.linalg.norm(axis)
                    limited = j.get("limited")
                    limited = self.d_limited if limited is None else (limited == "true")
                    rng = _floats(j.get("range"), np.array([0.0, 0.0]))
                    if self.degrees:
                        rng = rng * math.pi / 180.0
                    if not limited:
                        rng = np.array([1.0, -1.0])  # Bullet: lower > upper == unlimited
                    last = k == len(joints) - 1
                    if k == 0:
                        e = pos + R_body @ jpos
                        r = rot_p2t
                    else:
                        e = jpos - prev_pivot
                        r = np.array([0, 0, 0, 1.0])
                    d = -jpos if last else np.zeros(3)
                    lname = name if last else "link_dummy_%s_%d" % (name, k)
                    link = Link(lname, cur_parent, j.get("name"), JOINT_REVOLUTE, axis, r, e, d, is_dummy=not last)
                    link.lower, link.upper = float(rng[0]), float(rng[1])
                    dmp = j.get("damping")
                    arm = j.get("armature")
                    if self.use_mjcf_damping:
                        link.damping = float(dmp) if dmp is not None else self.d_damping
                    if self.use_mjcf_armature:
                        link.armature = float(arm) if arm is not None else self.d_armature
                    self.links.append(link)
                    cur_parent = len(self.links) - 1
                    prev_pivot = jpos
                my_index = cur_parent
                link = self.links[my_index]
            link.geoms = [self._parse_geom(g, np.zeros(3)) for g in geoms]

        for c in children:
            self._parse_body(c, my_index, False)

    def _finalize_inertia(self):
        for link in [self.base] + self.links:
            vol = sum(g.volume() for g in link.geoms)
            link.mass = DENSITY * vol
            if link.geoms:
                lo = np.min([g.aabb()[0] for g in link.geoms], axis=0)
                hi = np.max([g.aabb()[1] for g in link.geoms], axis=0)
                l = hi - lo
                m = link.mass
                link.inertia = m / 12.0 * np.array([l[1] ** 2 + l[2] ** 2, l[0] ** 2 + l[2] ** 2, l[0] ** 2 + l[1] ** 2])
                link.aabb = (lo, hi)
            else:
                link.inertia = np.zeros(3)
                link.aabb = None


# ----------------------------------------------------------------------------- flat table
def compile_mjcf(path: str, name: str, power_coef: Dict[str, float], base_power: float,
                 foot_names: List[str], use_mjcf_damping=False, use_mjcf_armature=False) -> dict:
    """Return the flat, JSON-serialisable articulation + collision table.

    ``power_coef`` / ``base_power`` / ``foot_names`` restate the robot classes of the reference
    (``mocca_envs/robots.py:230-256`` Walker3D, ``:407-436`` Monkey3D).
    """
    c = MJCFCompiler(path, use_mjcf_damping, use_mjcf_armature)
    links = c.links
    nl = len(links)
    dof_of_link = []
    nd = 0
    for l in links:
        if l.joint_type == JOINT_REVOLUTE:
            dof_of_link.append(nd)
            nd += 1
        else:
            dof_of_link.append(-1)
    joint_names = [l.joint_name for l in links if l.joint_type == JOINT_REVOLUTE]
    gains = [base_power * power_coef[n] for n in joint_names]

    geoms = []
    for li, l in enumerate([c.base] + links):
        for g in l.geoms:
            R = quat_to_mat(g.quat)
            if g.type == GEOM_CAPSULE:
                p0 = g.pos - R @ np.array([0, 0, g.size[1]])
                p1 = g.pos + R @ np.array([0, 0, g.size[1]])
            else:
                p0 = p1 = g.pos
            geoms.append(dict(name=g.name, link=li - 1, type=g.type, pos=g.pos.tolist(), quat=g.quat.tolist(),
                              size=g.size.tolist(), p0=p0.tolist(), p1=p1.tolist(), group=g.group, mask=g.mask,
                              friction=g.friction))

    def thresh(l):
        if l.aabb is None:
            return 0.0
        lo, hi = l.aabb
        return CONTACT_BREAKING_THRESHOLD * (0.5 * float(np.linalg.norm(hi - lo)) + float(np.linalg.norm(0.5 * (lo + hi))))

    def link_filter(l):
        # BulletMJCFImporter::getCollisionGroupAndMask: the last collision geom wins
        if not l.geoms:
            return (0, 0)
        return (l.geoms[-1].group, l.geoms[-1].mask)

    table = dict(
        name=name,
        planar=bool(c.planar),
        source=path.split("/mocca_envs/")[-1],
        conventions=dict(density=DENSITY, mjcf_damping=use_mjcf_damping, mjcf_armature=use_mjcf_armature,
                         inertia="aabb-of-compound", com="body-origin"),
        base=dict(name=c.base.name, mass=c.base.mass, inertia=c.base.inertia.tolist(), init_pos=c.base_pos.tolist(),
                  contact_threshold=thresh(c.base), group=link_filter(c.base)[0], mask=link_filter(c.base)[1]),
        n_links=nl,
        n_dof=nd,
        link_names=[l.name for l in links],
        joint_names_all=[l.joint_name for l in links],
        parent=[l.parent for l in links],
        joint_type=[l.joint_type for l in links],
        dof_of_link=dof_of_link,
        axis=[l.axis.tolist() for l in links],
        rot_parent_to_this=[l.rot_parent_to_this.tolist() for l in links],
        e_vec=[l.e_vec.tolist() for l in links],
        d_vec=[l.d_vec.tolist() for l in links],
        mass=[l.mass for l in links],
        inertia=[l.inertia.tolist() for l in links],
        contact_threshold=[thresh(l) for l in links],
        group=[link_filter(l)[0] for l in links],
        mask=[link_filter(l)[1] for l in links],
        # per-dof (ordered joints of robots.py:163-172)
        joint_names=joint_names,
        link_of_dof=[i for i, l in enumerate(links) if l.joint_type == JOINT_REVOLUTE],
        lower=[l.lower for l in links if l.joint_type == JOINT_REVOLUTE],
        upper=[l.upper for l in links if l.joint_type == JOINT_REVOLUTE],
        damping=[l.damping for l in links if l.joint_type == JOINT_REVOLUTE],
        armature=[l.armature for l in links if l.joint_type == JOINT_REVOLUTE],
        gain=gains,
        foot_links=[[l.name for l in links].index(f) for f in foot_names],
        foot_names=list(foot_names),
        geoms=geoms,
        total_mass=c.base.mass + sum(l.mass for l in links),
    )
    return table


# ----------------------------------------------------------------------------- self-collision pairs
def fk_links(t: dict, q):
    """Forward kinematics of a compiled table at joint angles q (base at the origin, identity orientation):
    returns (pos[n_links + 1, 3], rot[n_links + 1, 3, 3]) of the link COM frames (local -> world), index 0 = base.
    Same recurrence as btMultiBody: R_parent_to_this = R(axis, -q) * zeroRotParentToThis."""
    nl = t["n_links"]
    pos = np.zeros((nl + 1, 3))
    rot = np.zeros((nl + 1, 3, 3))
    rot[0] = np.eye(3)
    for i in range(nl):
        p = t["parent"][i] + 1
        Rz = quat_to_mat(t["rot_parent_to_this"][i])  # parent -> this at q = 0
        if t["joint_type"][i] == JOINT_REVOLUTE:
            ax = np.array(t["axis"][i])
            a = -q[t["dof_of_link"][i]]
            K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
            Rq = np.eye(3) + math.sin(a) * K + (1 - math.cos(a)) * (K @ K)
            Rp2t = Rq @ Rz
        else:
            Rp2t = Rz
        rot[i + 1] = rot[p] @ Rp2t.T
        pos[i + 1] = pos[p] + rot[p] @ np.array(t["e_vec"][i]) + rot[i + 1] @ np.array(t["d_vec"][i])
    return pos, rot


def _seg_dist(p1, q1, p2, q2):
    """Closest distance between segments p1q1 and p2q2 (Ericson, Real-Time Collision Detection 5.1.9)."""
    d1, d2, r = q1 - p1, q2 - p2, p1 - p2
    a, e, f = d1 @ d1, d2 @ d2, d2 @ r
    if a <= 1e-12 and e <= 1e-12:
        return float(np.linalg.norm(r))
    if a <= 1e-12:
        s, tt = 0.0, min(max(f / e, 0.0), 1.0)
    else:
        c = d1 @ r
        if e <= 1e-12:
            tt, s = 0.0, min(max(-c / a, 0.0), 1.0)
        else:
            b = d1 @ d2
            den = a * e - b * b
            s = min(max((b * f - c * e) / den, 0.0), 1.0) if den > 1e-12 else 0.0
            tt = (b * s + f) / e
            if tt < 0:
                tt, s = 0.0, min(max(-c / a, 0.0), 1.0)
            elif tt > 1:
                tt, s = 1.0, min(max((b - c) / a, 0.0), 1.0)
    return float(np.linalg.norm(p1 + d1 * s - p2 - d2 * tt))


def self_collision_pairs(t: dict, samples: int = 4000, margin: float = 0.05, seed: int = 0):
    """Geom pairs that can touch under URDF_USE_SELF_COLLISION | URDF_USE_SELF_COLLISION_EXCLUDE_ALL_PARENTS
    (robots.py:259-264): different links, neither an ancestor of the other, Bullet's two-way group/mask filter
    (BulletMJCFImporter: group = contype, mask = conaffinity of the link's last geom), sphere / capsule geoms only,
    links not rigidly attached to the same joint frame.  Pairs that never come within ``margin`` of each other in
    ``samples`` uniformly drawn poses inside the joint limits are dropped (documented reduction)."""
    par = t["parent"]
    nl = t["n_links"]

    def ancestors(l):
        out = set()
        while l >= 0:
            l = par[l]
            out.add(l)
        return out

    def owner(l):
        while l >= 0 and t["joint_type"][l] != JOINT_REVOLUTE:
            l = par[l]
        return l

    groups = [t["base"]["group"]] + t["group"]
    masks = [t["base"]["mask"]] + t["mask"]
    G = t["geoms"]
    cand = []
    for i in range(len(G)):
        for j in range(i + 1, len(G)):
            a, b = G[i]["link"], G[j]["link"]
            if a == b or a in ancestors(b) or b in ancestors(a) or owner(a) == owner(b):
                continue
            if G[i]["type"] == GEOM_BOX or G[j]["type"] == GEOM_BOX:
                continue
            if not ((groups[a + 1] & masks[b + 1]) and (groups[b + 1] & masks[a + 1])):
                continue
            cand.append((i, j))
    rng = np.random.RandomState(seed)
    lo, hi = np.array(t["lower"]), np.array(t["upper"])
    un = lo > hi
    lo, hi = np.where(un, -math.pi, lo), np.where(un, math.pi, hi)
    mind = np.full(len(cand), np.inf)
    for _ in range(samples):
        q = lo + (hi - lo) * rng.uniform(0, 1, len(lo))
        pos, rot = fk_links(t, q)
        ends = []
        for g in G:
            li = g["link"] + 1
            ends.append((pos[li] + rot[li] @ np.array(g["p0"]), pos[li] + rot[li] @ np.array(g["p1"])))
        for k, (i, j) in enumerate(cand):
            d = _seg_dist(ends[i][0], ends[i][1], ends[j][0], ends[j][1]) - G[i]["size"][0] - G[j]["size"][0]
            if d < mind[k]:
                mind[k] = d
    return [[i, j] for k, (i, j) in enumerate(cand) if mind[k] < margin], len(cand)


# ----------------------------------------------------------------------------- robot classes
WALKER3D_POWER = {  # mocca_envs/robots.py:234-256
    "abdomen_z": 60, "abdomen_y": 80, "abdomen_x": 60,
    "right_hip_x": 80, "right_hip_z": 60, "right_hip_y": 100, "right_knee": 90, "right_ankle": 60,
    "left_hip_x": 80, "left_hip_z": 60, "left_hip_y": 100, "left_knee": 90, "left_ankle": 60,
    "right_shoulder_x": 60, "right_shoulder_z": 60, "right_shoulder_y": 50, "right_elbow": 60,
    "left_shoulder_x": 60, "left_shoulder_z": 60, "left_shoulder_y": 50, "left_elbow": 60,
}

MONKEY3D_POWER = {  # mocca_envs/robots.py:410-434
    "abdomen_z": 60, "abdomen_y": 60, "abdomen_x": 60,
    "right_hip_x": 50, "right_hip_z": 50, "right_hip_y": 50, "right_knee": 30, "right_ankle": 10,
    "left_hip_x": 50, "left_hip_z": 50, "left_hip_y": 50, "left_knee": 30, "left_ankle": 10,
    "right_shoulder_x": 100, "right_shoulder_y": 100, "right_elbow_z": 60, "right_elbow_y": 100, "right_hand": 80,
    "left_shoulder_x": 100, "left_shoulder_y": 100, "left_elbow_z": 60, "left_elbow_y": 100, "left_hand": 80,
}


def compile_walker3d(data_dir: str, **kw) -> dict:
    t = compile_mjcf(data_dir + "/robots/walker3d.xml", "walker3d", WALKER3D_POWER, 1.0,
                     ["right_foot", "left_foot"], **kw)
    # robots.py:275-302 -- T-pose base + "running_start" joint pose used by both Walker3D envs
    pose = [0.0] * 21
    for i in (5, 6):
        pose[i] = -math.pi / 8
    pose[10] = math.pi / 10
    pose[13] = pose[17] = math.pi / 3
    pose[14] = -math.pi / 6
    pose[18] = math.pi / 6
    pose[16] = pose[20] = math.pi / 3
    t["base_joint_angles"] = pose
    t["base_position"] = [0.0, 0.0, 1.32]
    t["right_joint_indices"] = [3, 4, 5, 6, 7, 13, 14, 15, 16]  # robots.py:282-284
    t["left_joint_indices"] = [8, 9, 10, 11, 12, 17, 18, 19, 20]  # robots.py:285-287
    t["negation_joint_indices"] = [0, 2]  # robots.py:288
    t["self_pairs"], t["self_pairs_candidates"] = self_collision_pairs(t)
    return t


MIKE_POWER = {  # mocca_envs/robots.py:477-499
    "abdomen_z": 0, "abdomen_y": 0, "abdomen_x": 0,
    "right_hip_x": 80, "right_hip_z": 60, "right_hip_y": 100, "right_knee": 90, "right_ankle": 60,
    "left_hip_x": 80, "left_hip_z": 60, "left_hip_y": 100, "left_knee": 90, "left_ankle": 60,
    "right_shoulder_x": 30, "right_shoulder_z": 30, "right_shoulder_y": 25, "right_elbow": 30,
    "left_shoulder_x": 30, "left_shoulder_z": 30, "left_shoulder_y": 25, "left_elbow": 30,
}


WALKER2D_POWER = {  # mocca_envs/robots.py:342-350
    "torso_joint": 100, "thigh_joint": 100, "leg_joint": 100, "foot_joint": 50,
    "thigh_left_joint": 100, "leg_left_joint": 100, "foot_left_joint": 50,
}
CRAB2D_POWER = {  # mocca_envs/robots.py:380-387
    "thigh_left_joint": 100, "leg_left_joint": 100, "foot_left_joint": 50,
    "thigh_joint": 100, "leg_joint": 100, "foot_joint": 50,
}


def _planar_family(t: dict, right, left) -> dict:
    """Walker2D / Crab2D (robots.py:338-404): zero base pose (set_base_pose :352-354), identity orientation, no negated
    joints.  coordinate="global" is not understood by Bullet's MJCF importer (the reference's xml says so: "CHANGES:
    see hopper.xml"), it reads every pos / fromto as body-local; all nested bodies have no pos, so every link frame
    (= inertial frame, convention "com": "body-origin") coincides with the pelvis origin at zero joint angles.  The
    pelvis body sits at z = -1.35 under the fixed base and Walker2DCustomEnv resets that base to z = +1.35
    (env_locomotion.py:287, robots.py:198-200 -> resetBasePositionAndOrientation on the multibody), so the pelvis
    origin starts at the world origin and the geoms stand where the xml's global coordinates put them."""
    assert t.get("planar")
    t["base_joint_angles"] = [0.0] * t["n_dof"]
    t["base_position"] = [0.0, 0.0, 0.0]
    t["right_joint_indices"] = list(right)
    t["left_joint_indices"] = list(left)
    t["negation_joint_indices"] = []
    t["self_pairs"], t["self_pairs_candidates"] = self_collision_pairs(t)
    t.setdefault("conventions", {})["planar_base"] = (
        "ignorex / ignorez / ignorey root joints = free base confined to the x-z plane (exact: all axes are y)")
    return t


def compile_walker2d(data_dir: str, **kw) -> dict:
    """Walker2D (robots.py:338-370): 7 hinges about y, feet "foot" / "foot_left", no self-collision flag."""
    t = compile_mjcf(data_dir + "/robots/walker2d.xml", "walker2d", WALKER2D_POWER, 1.0, ["foot", "foot_left"], **kw)
    return _planar_family(t, [1, 2, 3], [4, 5, 6])


def compile_crab2d(data_dir: str, **kw) -> dict:
    """Crab2D (robots.py:373-404): 6 hinges about +-y, loaded with the self-collision flags."""
    t = compile_mjcf(data_dir + "/robots/crab2d.xml", "crab2d", CRAB2D_POWER, 1.0, ["foot", "foot_left"], **kw)
    return _planar_family(t, [0, 1, 2], [3, 4, 5])


def _walker_family(t: dict) -> dict:
    """Mirroring tables shared by the Walker3D subclasses (robots.py:280-290)."""
    t["right_joint_indices"] = [3, 4, 5, 6, 7, 13, 14, 15, 16]
    t["left_joint_indices"] = [8, 9, 10, 11, 12, 17, 18, 19, 20]
    t["negation_joint_indices"] = [0, 2]
    t["self_pairs"], t["self_pairs_candidates"] = self_collision_pairs(t)
    return t


def compile_child3d(data_dir: str, **kw) -> dict:
    """Child3D (robots.py:326-335): Walker3D's joints on child3d.xml, power 0.4, base at z = 0.38;
    Child3DCustomEnv (env_locomotion.py:317-327) starts it in the "crawl" pose (robots.py:314-323) and ends the
    episode below a relative height of 0.1."""
    t = compile_mjcf(data_dir + "/robots/child3d.xml", "child3d", WALKER3D_POWER, 0.4,
                     ["right_foot", "left_foot"], **kw)
    d = math.pi / 180
    pose = [0.0] * 21
    pose[13] = pose[17] = math.pi / 2
    pose[14] = pose[18] = math.pi / 2
    pose[16] = pose[20] = math.pi / 3
    pose[5] = pose[10] = -math.pi / 2
    pose[6] = pose[11] = -120 * d
    pose[7] = pose[12] = -20 * d
    t["base_joint_angles"] = pose
    t["base_position"] = [0.0, 0.0, 0.38]
    # pybullet.getQuaternionFromEuler([0, 90 deg, 0]) = (0, sin 45, 0, cos 45), xyzw
    t["base_orientation"] = [0.0, math.sin(math.pi / 4), 0.0, math.cos(math.pi / 4)]
    t["termination_height"] = 0.1  # env_locomotion.py:320
    return _walker_family(t)


def compile_mike(data_dir: str, **kw) -> dict:
    """Mike (robots.py:474-513): Walker3D's joints on mike.xml with its own power table; the waist link's mass is set
    to 8 after loading (changeDynamics, :506-510).  MikeStepperEnv (env_locomotion.py:843-851) starts it at
    (0.3, 0, 1.0)."""
    t = compile_mjcf(data_dir + "/robots/mike.xml", "mike", MIKE_POWER, 1.0, ["right_foot", "left_foot"], **kw)
    w = t["link_names"].index("waist")
    old = t["mass"][w]
    # changeDynamics(mass=m) on a multibody link: Bullet re-derives the local inertia from the link's collision shape
    # with the new mass; for the single sphere of this link that is the old diagonal scaled by m / m_old
    # (hypothesis recorded in conventions; btCompoundShape's AABB approximation would differ)
    t["mass"][w] = 8.0
    t["inertia"][w] = [x * 8.0 / old for x in t["inertia"][w]]
    t["total_mass"] = float(t["base"]["mass"] + sum(t["mass"]))
    t.setdefault("conventions", {})["mike_waist_mass"] = "mass 8 (robots.py:510), inertia scaled by 8 / %.6g" % old
    pose = [0.0] * 21  # MikeStepperEnv inherits Walker3DStepperEnv's set_base_pose("running_start")
    for i in (5, 6):
        pose[i] = -math.pi / 8
    pose[10] = math.pi / 10
    pose[13] = pose[17] = math.pi / 3
    pose[14] = -math.pi / 6
    pose[18] = math.pi / 6
    pose[16] = pose[20] = math.pi / 3
    t["base_joint_angles"] = pose
    t["base_position"] = [0.0, 0.0, 1.32]           # Walker3D.set_base_pose default (robots.py:275)
    t["stepper_init_position"] = [0.3, 0.0, 1.0]    # env_locomotion.py:845
    return _walker_family(t)


def compile_monkey3d(data_dir: str, **kw) -> dict:
    t = compile_mjcf(data_dir + "/robots/monkey3d.xml", "monkey3d", MONKEY3D_POWER, 0.7,
                     ["right_hand", "left_hand"], **kw)
    d = math.pi / 180
    pose = [0.0] * 23  # robots.py:462-470 "monkey_start"
    pose[14] = -140 * d
    pose[15] = 180 * d
    pose[17] = -90 * d
    pose[19] = -170 * d
    pose[20] = 180 * d
    pose[22] = -90 * d
    pose[6] = pose[11] = -90 * d
    pose[7] = pose[12] = -90 * d
    t["base_joint_angles"] = pose
    t["base_position"] = [0.0, 0.0, 0.7]
    t["right_joint_indices"] = [3, 4, 5, 6, 7, 13, 14, 15, 16, 17]
    t["left_joint_indices"] = [8, 9, 10, 11, 12, 18, 19, 20, 21, 22]
    t["negation_joint_indices"] = [0, 2]
    # env_locomotion.py:1269,1424: the palm spheres whose contact with the target bar advances the step index
    t["palm_links"] = [t["link_names"].index("right_palm"), t["link_names"].index("left_palm")]
    t["self_pairs"], t["self_pairs_candidates"] = self_collision_pairs(t)
    return t


def save_table(table: dict, path: str):
    with open(path, "w") as f:
        json.dump(table, f, indent=1)
        f.write("\n")


def load_table(path: str) -> dict:
    with open(path) as f:
        return json.load(f)
