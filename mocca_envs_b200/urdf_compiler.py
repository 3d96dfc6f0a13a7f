"""URDF robot -> the same flat articulation / collision table ``model_compiler.py`` produces for MJCF robots.

Used for Cassie (reference ``mocca_envs/env_cassie.py:13-149`` loads
``data/robots/cassie/urdf/cassie_collide.urdf`` with ``URDF_USE_INERTIA_FROM_FILE | URDF_USE_SELF_COLLISION |
URDF_USE_SELF_COLLISION_EXCLUDE_ALL_PARENTS``).  Restates the conventions of Bullet's URDF importer and
``URDF2Bullet`` multibody conversion (third-party ``pybullet``; SURVEY.md App. B.1, recalled, unpinned):

* link indices are the DFS pre-order of the joint tree, children in file order of their joints;
* with inertia-from-file the link's multibody frame is its INERTIAL frame: origin at the URDF ``<inertial>`` origin,
  axes = ``rpy`` composed with the principal axes found by ``btMatrix3x3::diagonalize`` (Jacobi, restated below)
  when the tensor has products of inertia; the inertia is the resulting diagonal;
* ``rot_parent_to_this`` / ``e_vec`` / ``d_vec`` / ``axis`` follow ``setupRevolute(parentRotToThis, axis, parentComToThis
  PivotOffset, thisPivotToThisComOffset)``;
* mesh collision shapes become convex hulls of the mesh vertices with margin 0.001.  Against the ground plane a convex
  hull touches at hull vertices; the table keeps, per link, the support vertices of a fixed fan of directions as
  candidate contact points (spheres of the margin radius) -- a documented reduction of the hull.
"""
from __future__ import annotations

import math
import os
import struct
import xml.etree.ElementTree as ET

import numpy as np

from .model_compiler import CONTACT_BREAKING_THRESHOLD, GEOM_SPHERE, JOINT_FIXED, JOINT_REVOLUTE, mat_to_quat

HULL_MARGIN = 0.001  # PyBullet gUrdfDefaultCollisionMargin


def rpy_to_mat(rpy):
    r, p, y = [float(v) for v in rpy]
    cr, sr, cp, sp, cy, sy = math.cos(r), math.sin(r), math.cos(p), math.sin(p), math.cos(y), math.sin(y)
    return np.array([[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
                     [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
                     [-sp, cp * sr, cp * cr]])


def bt_diagonalize(M, threshold=1.0e-6, max_steps=30):
    """btMatrix3x3::diagonalize (Jacobi): returns (diagonalised matrix, rot) with rot^T M rot diagonal."""
    M = np.array(M, dtype=np.float64)
    rot = np.eye(3)
    eps = 2.220446049250313e-16
    step = max_steps
    while step > 0:
        p, q, r = 0, 1, 2
        mx = abs(M[0][1])
        v = abs(M[0][2])
        if v > mx:
            q, r, mx = 2, 1, v
        v = abs(M[1][2])
        if v > mx:
            p, q, r, mx = 1, 2, 0, v
        t = threshold * (abs(M[0][0]) + abs(M[1][1]) + abs(M[2][2]))
        if mx <= t:
            if mx <= eps * t:
                break
            step = 1
        mpq = M[p][q]
        theta = (M[q][q] - M[p][p]) / (2 * mpq)
        theta2 = theta * theta
        if theta2 * theta2 < 10.0 / eps:
            tt = 1 / (theta + math.sqrt(1 + theta2)) if theta >= 0 else 1 / (theta - math.sqrt(1 + theta2))
            cos = 1 / math.sqrt(1 + tt * tt)
            sin = cos * tt
        else:
            tt = 1 / (theta * (2 + 0.5 / theta2))
            cos = 1 - 0.5 * tt * tt
            sin = cos * tt
        M[p][q] = M[q][p] = 0
        M[p][p] -= tt * mpq
        M[q][q] += tt * mpq
        mrp, mrq = M[r][p], M[r][q]
        M[r][p] = M[p][r] = cos * mrp - sin * mrq
        M[r][q] = M[q][r] = cos * mrq + sin * mrp
        for i in range(3):
            mrp, mrq = rot[i][p], rot[i][q]
            rot[i][p] = cos * mrp - sin * mrq
            rot[i][q] = cos * mrq + sin * mrp
        step -= 1
    return M, rot


def load_stl_vertices(path):
    b = open(path, "rb").read()
    n = struct.unpack("<I", b[80:84])[0]
    if 84 + 50 * n == len(b):
        a = np.frombuffer(b, dtype=np.uint8, count=n * 50, offset=84).reshape(n, 50)
        return a[:, 12:48].copy().view("<f4").reshape(n * 3, 3).astype(np.float64)
    import re

    return np.array([[float(x) for x in m.groups()] for m in re.finditer(rb"vertex\s+(\S+)\s+(\S+)\s+(\S+)", b)])


def direction_fan(level):
    """Unit directions: level 0 = the 8 body diagonals (the hull's "bounding-box corners"), level 1 adds the 6 axis
    directions (14), level 2 adds the 12 face diagonals (26)."""
    dirs = []
    for x in (-1, 0, 1):
        for y in (-1, 0, 1):
            for z in (-1, 0, 1):
                nz = abs(x) + abs(y) + abs(z)
                if nz == 3 or (nz == 1 and level >= 1) or (nz == 2 and level >= 2):
                    d = np.array([x, y, z], dtype=np.float64)
                    dirs.append(d / np.linalg.norm(d))
    return dirs


def hull_support_points(verts, level):
    from scipy.spatial import ConvexHull

    hv = verts[ConvexHull(verts).vertices]
    idx = []
    for d in direction_fan(level):
        k = int(np.argmax(hv @ d))
        if k not in idx:
            idx.append(k)
    return hv[sorted(idx)], hv


HULL_VERTS = 32  # vertices kept per link hull for the hull-vs-hull narrow phase (one per lane of a warp)


def hull_fan_vertices(hv, n=HULL_VERTS):
    """Support vertices of the hull `hv` along n spherical-Fibonacci directions (duplicates removed, then padded by
    repeating the first vertex): the convex hull of these is the link's shape in the hull-vs-hull narrow phase -- an
    inner approximation of Bullet's btConvexHullShape (which keeps every mesh vertex), millimetres off at most."""
    idx = []
    ga = math.pi * (3.0 - math.sqrt(5.0))
    for k in range(n):
        z = 1.0 - 2.0 * (k + 0.5) / n
        r = math.sqrt(max(0.0, 1.0 - z * z))
        d = np.array([r * math.cos(ga * k), r * math.sin(ga * k), z])
        i = int(np.argmax(hv @ d))
        if i not in idx:
            idx.append(i)
    v = hv[idx]
    pad = np.repeat(v[:1], n - len(v), axis=0)
    return np.concatenate([v, pad]), len(v)


def hull_distance(A, B, iters=64):
    """Distance between the convex hulls of two vertex sets (Gilbert's iteration on the Minkowski difference)."""
    v = A[0] - B[0]
    for _ in range(iters):
        s = A[int(np.argmax(A @ -v))] - B[int(np.argmax(B @ v))]
        if v @ v - v @ s < 1e-12:
            break
        e = s - v
        v = v + min(1.0, max(0.0, -(v @ e) / (e @ e))) * e
    return float(np.linalg.norm(v))


def hull_collision_pairs(t, samples=3000, margin=0.08, seed=0):
    """Link pairs whose mesh hulls can touch under URDF_USE_SELF_COLLISION | ..._EXCLUDE_ALL_PARENTS (env_cassie.py:81-85):
    neither link an ancestor of the other, both with a collision hull (the achilles rods carry none and are filtered
    out anyway, env_cassie.py:140-149); pairs that never come within `margin` in `samples` poses drawn inside the joint
    limits are dropped (the same documented reduction as model_compiler.self_collision_pairs)."""
    from .model_compiler import fk_links

    par = t["parent"]

    def ancestors(l):
        out = set()
        while l >= 0:
            l = par[l]
            out.add(l)
        return out

    H = t["hulls"]
    cand = [(a, b) for a in range(len(H)) for b in range(a + 1, len(H))
            if H[a]["link"] not in ancestors(H[b]["link"]) and H[b]["link"] not in ancestors(H[a]["link"])
            and H[a]["link"] != H[b]["link"]]
    rng = np.random.RandomState(seed)
    lo, hi = np.array(t["lower"]), np.array(t["upper"])
    verts = [np.array(h["verts"]) for h in H]
    cen = [v.mean(0) for v in verts]
    rad = [float(np.linalg.norm(v - c, axis=1).max()) for v, c in zip(verts, cen)]
    mind = np.full(len(cand), np.inf)
    for _ in range(samples):
        q = lo + (hi - lo) * rng.uniform(0, 1, len(lo))
        pos, rot = fk_links(t, q)
        W = [v @ rot[h["link"] + 1].T + pos[h["link"] + 1] for v, h in zip(verts, H)]
        C = [rot[h["link"] + 1] @ c + pos[h["link"] + 1] for c, h in zip(cen, H)]
        for k, (a, b) in enumerate(cand):
            if np.linalg.norm(C[a] - C[b]) - rad[a] - rad[b] > min(mind[k], margin):
                continue
            mind[k] = min(mind[k], hull_distance(W[a], W[b]))
    return [[a, b] for k, (a, b) in enumerate(cand) if mind[k] < margin], len(cand)


def compile_urdf(path, name, fan_level=None, mesh_root=None):
    fan_level = fan_level or {}
    root = ET.parse(path).getroot()
    links, order = {}, []
    for l in root.findall("link"):
        inertial = l.find("inertial")
        xyz, rpy, mass, I = np.zeros(3), np.zeros(3), 0.0, np.zeros((3, 3))
        if inertial is not None:
            o = inertial.find("origin")
            if o is not None:
                xyz = np.array([float(v) for v in o.get("xyz", "0 0 0").split()])
                rpy = np.array([float(v) for v in o.get("rpy", "0 0 0").split()])
            mass = float(inertial.find("mass").get("value"))
            it = inertial.find("inertia")
            g = lambda k: float(it.get(k, 0))
            I = np.array([[g("ixx"), g("ixy"), g("ixz")], [g("ixy"), g("iyy"), g("iyz")], [g("ixz"), g("iyz"), g("izz")]])
        if I[0, 1] == 0 and I[0, 2] == 0 and I[1, 2] == 0:
            diag, basis = np.diag(I).copy(), np.eye(3)
        else:
            D, basis = bt_diagonalize(I)
            diag = np.diag(D).copy()
        px, py, pz = diag
        if px < 0 or px > py + pz or py < 0 or py > px + pz or pz < 0 or pz > px + py:
            diag, basis = np.zeros(3), np.eye(3)  # "Bad inertia tensor properties, setting inertia to zero"
        meshes = []
        for c in l.findall("collision"):
            o = c.find("origin")
            cx = np.array([float(v) for v in o.get("xyz", "0 0 0").split()]) if o is not None else np.zeros(3)
            cr = np.array([float(v) for v in o.get("rpy", "0 0 0").split()]) if o is not None else np.zeros(3)
            m = c.find("geometry").find("mesh")
            if m is not None:
                meshes.append((m.get("filename"), cx, cr, np.array([float(v) for v in m.get("scale", "1 1 1").split()])))
        fr = l.find("contact")
        friction = 0.5  # Bullet default lateral friction
        if fr is not None and fr.find("lateral_friction") is not None:
            friction = float(fr.find("lateral_friction").get("value"))
        links[l.get("name")] = dict(name=l.get("name"), xyz=xyz, R=rpy_to_mat(rpy) @ basis, mass=mass, inertia=diag,
                                    meshes=meshes, friction=friction, children=[])
        order.append(l.get("name"))
    joints = []
    child_names = set()
    for j in root.findall("joint"):
        o = j.find("origin")
        xyz = np.array([float(v) for v in o.get("xyz", "0 0 0").split()]) if o is not None else np.zeros(3)
        rpy = np.array([float(v) for v in o.get("rpy", "0 0 0").split()]) if o is not None else np.zeros(3)
        ax = j.find("axis")
        axis = np.array([float(v) for v in ax.get("xyz").split()]) if ax is not None else np.array([1.0, 0, 0])
        lim = j.find("limit")
        jt = j.get("type")
        lower, upper = 1.0, -1.0  # Bullet: lower > upper == unlimited (continuous joints)
        if jt == "revolute" and lim is not None:
            lower, upper = float(lim.get("lower", 0)), float(lim.get("upper", 0))
        dyn = j.find("dynamics")
        damping = float(dyn.get("damping", 0)) if dyn is not None else 0.0
        jd = dict(name=j.get("name"), type=jt, xyz=xyz, R=rpy_to_mat(rpy), axis=axis / np.linalg.norm(axis),
                  parent=j.find("parent").get("link"), child=j.find("child").get("link"), lower=lower, upper=upper,
                  damping=damping)
        joints.append(jd)
        links[jd["parent"]]["children"].append(jd)
        child_names.add(jd["child"])
    roots = [n for n in order if n not in child_names]
    assert len(roots) == 1
    base = links[roots[0]]

    flat = []  # (link dict, parent index, joint dict)

    def dfs(link, parent_index):
        for jd in link["children"]:
            flat.append((links[jd["child"]], parent_index, jd))
            dfs(links[jd["child"]], len(flat) - 1)

    dfs(base, -1)

    mesh_root = mesh_root or os.path.dirname(path)

    hulls = []

    def link_points(link, li):
        pts, lo, hi = [], None, None
        allv = []
        for fn, cx, cr, scale in link["meshes"]:
            v = load_stl_vertices(os.path.normpath(os.path.join(mesh_root, fn))) * scale
            v = v @ rpy_to_mat(cr).T + cx  # collision frame -> link frame
            v = (v - link["xyz"]) @ link["R"]  # link frame -> inertial frame
            sp, hv = hull_support_points(v, fan_level.get(link["name"], 0))
            allv.append(hv)
            pts.extend(sp.tolist())
            l2, h2 = hv.min(0) - HULL_MARGIN, hv.max(0) + HULL_MARGIN
            lo = l2 if lo is None else np.minimum(lo, l2)
            hi = h2 if hi is None else np.maximum(hi, h2)
        thr = 0.0
        if lo is not None:
            thr = CONTACT_BREAKING_THRESHOLD * (0.5 * float(np.linalg.norm(hi - lo)) + float(np.linalg.norm(0.5 * (lo + hi))))
        if allv and li >= 1:  # (the base is an ancestor of every link: never a partner under EXCLUDE_ALL_PARENTS)
            fv, n = hull_fan_vertices(np.concatenate(allv))
            hulls.append(dict(link=li - 1, verts=fv.tolist(), n=n))
        return pts, thr

    geoms = []
    thresholds = []
    for li, link in enumerate([base] + [f[0] for f in flat]):
        pts, thr = link_points(link, li)
        thresholds.append(thr)
        for k, p in enumerate(pts):
            geoms.append(dict(name="%s_v%d" % (link["name"], k), link=li - 1, type=GEOM_SPHERE, pos=list(p),
                              quat=[0, 0, 0, 1.0], size=[HULL_MARGIN, 0, 0], p0=list(p), p1=list(p), group=1, mask=1,
                              friction=link["friction"]))

    table = dict(
        name=name, source=path.split("/mocca_envs/")[-1],
        conventions=dict(inertia="from-file, principal axes by btMatrix3x3::diagonalize", com="urdf inertial origin",
                         collision="support vertices of the mesh convex hulls, margin %g" % HULL_MARGIN),
        base=dict(name=base["name"], mass=base["mass"], inertia=base["inertia"].tolist(), init_pos=[0, 0, 0],
                  contact_threshold=thresholds[0], group=1, mask=1,
                  inertial_xyz=base["xyz"].tolist(), inertial_rot=base["R"].tolist()),
        n_links=len(flat), link_names=[f[0]["name"] for f in flat], joint_names_all=[f[2]["name"] for f in flat],
        parent=[f[1] for f in flat],
        joint_type=[JOINT_REVOLUTE if f[2]["type"] in ("revolute", "continuous") else JOINT_FIXED for f in flat],
        mass=[f[0]["mass"] for f in flat], inertia=[f[0]["inertia"].tolist() for f in flat],
        contact_threshold=thresholds[1:], group=[1] * len(flat), mask=[1] * len(flat), geoms=geoms,
        hulls=hulls, hull_margin=HULL_MARGIN, link_friction=[base["friction"]] + [f[0]["friction"] for f in flat],
    )
    axis, rot, ev, dv = [], [], [], []
    for link, pidx, jd in flat:
        par = base if pidx < 0 else flat[pidx][0]
        R_rel = par["R"].T @ jd["R"] @ link["R"]  # this inertial axes in the parent's inertial axes
        rot.append(mat_to_quat(R_rel.T).tolist())
        ev.append((par["R"].T @ (jd["xyz"] - par["xyz"])).tolist())
        dv.append((link["R"].T @ link["xyz"]).tolist())
        axis.append((link["R"].T @ jd["axis"]).tolist() if table["joint_type"][len(axis)] == JOINT_REVOLUTE else [0.0, 0, 0])
    table.update(axis=axis, rot_parent_to_this=rot, e_vec=ev, d_vec=dv)
    dof_of_link, nd = [], 0
    for t in table["joint_type"]:
        dof_of_link.append(nd if t == JOINT_REVOLUTE else -1)
        nd += t == JOINT_REVOLUTE
    rev = [i for i, t in enumerate(table["joint_type"]) if t == JOINT_REVOLUTE]
    table.update(n_dof=nd, dof_of_link=dof_of_link, link_of_dof=rev, joint_names=[flat[i][2]["name"] for i in rev],
                 lower=[flat[i][2]["lower"] for i in rev], upper=[flat[i][2]["upper"] for i in rev],
                 damping=[flat[i][2]["damping"] for i in rev], armature=[0.0] * nd,
                 total_mass=base["mass"] + sum(f[0]["mass"] for f in flat))
    return table


# ----------------------------------------------------------------------------- Cassie (env_cassie.py:13-62)
CASSIE_POWER = {  # env_cassie.py:41-56 (the left/right labels of the passive joints are swapped there; same values)
    "hip_abduction_left": 112.5, "hip_rotation_left": 112.5, "hip_flexion_left": 195.2, "knee_joint_left": 195.2,
    "knee_to_shin_right": 200, "ankle_joint_right": 200, "toe_joint_left": 45.0,
    "hip_abduction_right": 112.5, "hip_rotation_right": 112.5, "hip_flexion_right": 195.2, "knee_joint_right": 195.2,
    "knee_to_shin_left": 200, "ankle_joint_left": 200, "toe_joint_right": 45.0,
}
CASSIE_BASE_ANGLES = [0.035615837, -0.01348790, 0.391940848, -0.95086160, -0.08376049, 1.305643634, -1.61174064] * 2
CASSIE_ROD_ANGLES = [-0.8967891835, 0.063947468, -0.8967891835, -0.063947468]  # env_cassie.py:39
CASSIE_JOINT_DAMPING = [1, 1, 1, 1, 0.1, 0, 1, 1, 1, 1, 1, 0.1, 0, 1]  # env_cassie.py:57


def compile_cassie(data_dir: str) -> dict:
    path = data_dir + "/robots/cassie/urdf/cassie_collide.urdf"
    t = compile_urdf(path, "cassie", fan_level={"left_toe": 1, "right_toe": 1, "left_tarsus": 1, "right_tarsus": 1})
    names = t["joint_names"]
    # ordered_joints: every joint whose name does not start with "fixed" (env_cassie.py:192-202); the achilles-rod
    # joints are named fixed_*_achilles_rod_joint_{z,y} and therefore are dofs but not ordered joints
    ordered = [d for d, n in enumerate(names) if not n.startswith("fixed")]
    rods = [d for d, n in enumerate(names) if "achilles" in n]
    assert len(ordered) == 14 and len(rods) == 4
    t["ordered_dofs"] = ordered
    t["rod_dofs"] = rods
    t["powered_joint_inds"] = [0, 1, 2, 3, 6, 7, 8, 9, 10, 13]  # env_cassie.py:59
    t["spring_joint_inds"] = [4, 11]  # env_cassie.py:60
    gain = [0.0] * t["n_dof"]
    damping = list(t["damping"])
    for k, d in enumerate(ordered):
        gain[d] = 1.0 * CASSIE_POWER[names[d]]  # torque limits (env_cassie.py:193-195), power = 1
        damping[d] = float(CASSIE_JOINT_DAMPING[k])  # changeDynamics(jointDamping=...) (env_cassie.py:197-201)
    t["gain"] = gain
    t["damping"] = damping
    pose = [0.0] * t["n_dof"]
    for k, d in enumerate(ordered):
        pose[d] = CASSIE_BASE_ANGLES[k]
    for k, d in enumerate(rods):
        pose[d] = CASSIE_ROD_ANGLES[k]
    t["base_joint_angles"] = pose
    t["base_position"] = [0.0, 0.0, 1.085]  # env_cassie.py:17 -- applied to the base INERTIAL frame by Bullet
    t["foot_names"] = ["right_toe", "left_toe"]  # env_cassie.py:73
    t["foot_links"] = [t["link_names"].index(f) for f in t["foot_names"]]
    t["right_joint_indices"], t["left_joint_indices"], t["negation_joint_indices"] = [], [], []
    # loop closures: createConstraint(JOINT_POINT2POINT) tarsus <-> achilles rod, pivots in the links' inertial frames
    # (env_cassie.py:114-137); PyBullet's default maxForce for user constraints is 500
    t["p2p"] = [
        dict(link_a=t["link_names"].index("left_tarsus"), link_b=t["link_names"].index("left_achilles_rod"),
             pivot_a=[-0.22735404, 0.05761813, 0.00711836], pivot_b=[0.254001, 0, 0], max_impulse=500.0),
        dict(link_a=t["link_names"].index("right_tarsus"), link_b=t["link_names"].index("right_achilles_rod"),
             pivot_a=[-0.22735404, 0.05761813, -0.00711836], pivot_b=[0.254001, 0, 0], max_impulse=500.0),
    ]
    # PD gains of CassieEnv (env_cassie.py:291-319): [powered(10), spring(2)], kp / 1.9, kd = kp / 10
    kp = [100, 100, 88, 96, 50, 100, 100, 88, 96, 50, 400, 400]
    t["pd_kp"] = [k / 1.9 for k in kp]
    t["pd_kd"] = [k / 1.9 / 10 for k in kp]
    # mesh-hull self-collision (env_cassie.py:81-85): left-leg vs right-leg links are the only non-ancestor pairs
    t["hull_pairs"], t["hull_pairs_considered"] = hull_collision_pairs(t)
    return t
