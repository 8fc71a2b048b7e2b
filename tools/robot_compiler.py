#!/usr/bin/env python3
"""Robot model compiler: URDF + SRDF -> (a) a compact JSON robot model, (b) a straight-line CUDA
forward-kinematics routine produced by a small symbolic tracer with constant folding.

This replaces, for the B200 engine, the role the external "cricket" tracer plays for the reference
(reference README.md:189-194): the reference ships 3-17 k lines of generated C++ per robot
(src/impl/vamp/robots/*.hh); we instead derive everything from the robot *description*
(resources/<robot>/<robot>_spherized.urdf + <robot>.srdf) and keep only data:

  * rigid bodies: links joined by fixed joints are merged; one 3x4 frame per moving body;
  * collision links: the URDF links that carry spheres, each with its fine spheres (expressed in
    its body's frame) and OUR OWN minimal enclosing sphere (the reference's per-link bounding
    spheres exist only inside its generated code, e.g. panda.hh:5618-5627; because they are a pure
    acceleration structure -- verdicts are decided by fine spheres -- any conservative bound gives
    the same verdicts);
  * allowed self-collision link pairs: all collision-link pairs not disabled by the SRDF.

Run (dev time, in the container where the robot descriptions are mounted):
    python tools/robot_compiler.py --resources /root/reference/resources --out vamp_mvt_b200
Outputs  vamp_mvt_b200/robots/<robot>.json  and  vamp_mvt_b200/csrc/gen/<robot>_fk.cuh
(both committed; nothing at run time reads the resources directory).
"""
from __future__ import annotations

import argparse
import json
import math
import xml.etree.ElementTree as ET
from pathlib import Path

import numpy as np

# Facts that are not in the URDF: the order of the actuated joints in a configuration vector and the
# end-effector frame used for attachments (reference robots/panda.hh:20-28, ur5.hh:20-27,
# fetch.hh:20-29, baxter.hh:20-35) and the motion-validation resolution (panda.hh:18, baxter.hh:18).
ROBOTS = {
    "panda": dict(
        joints=["panda_joint%d" % i for i in range(1, 8)],
        end_effector="panda_grasptarget",
        resolution=32,
    ),
    "ur5": dict(
        joints=[
            "shoulder_pan_joint",
            "shoulder_lift_joint",
            "elbow_joint",
            "wrist_1_joint",
            "wrist_2_joint",
            "wrist_3_joint",
        ],
        end_effector="robotiq_85_base_link",
        resolution=32,
        # not disabled by ur5.srdf, yet absent from the reference's generated pair list
        # (robots/ur5.hh has no "fts_robotside vs. wrist_2_link" block)
        extra_disabled=[("fts_robotside", "wrist_2_link")],
    ),
    "fetch": dict(
        joints=[
            "torso_lift_joint",
            "shoulder_pan_joint",
            "shoulder_lift_joint",
            "upperarm_roll_joint",
            "elbow_flex_joint",
            "forearm_roll_joint",
            "wrist_flex_joint",
            "wrist_roll_joint",
        ],
        end_effector="gripper_link",
        resolution=32,
    ),
    "baxter": dict(
        joints=[
            "left_s0", "left_s1", "left_e0", "left_e1", "left_w0", "left_w1", "left_w2",
            "right_s0", "right_s1", "right_e0", "right_e1", "right_w0", "right_w1", "right_w2",
        ],
        end_effector="right_gripper",
        resolution=64,
    ),
}

BOUND_MARGIN = 1e-6  # metres added to our enclosing spheres so they are strictly conservative
BOX_MARGIN = 0.02  # rad / m: static analyses hold for joints within [lower - m, upper + m]


# --------------------------------------------------------------------------------------------
# URDF / SRDF parsing
# --------------------------------------------------------------------------------------------
def _floats(s, n=3):
    v = [float(t) for t in s.split()] if s else [0.0] * n
    assert len(v) == n
    return v


def rpy_matrix(r, p, y):
    """URDF fixed-axis roll/pitch/yaw: R = Rz(y) Ry(p) Rx(r)."""
    cr, sr, cp, sp, cy, sy = math.cos(r), math.sin(r), math.cos(p), math.sin(p), math.cos(y), math.sin(y)
    return np.array(
        [
            [cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
            [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
            [-sp, cp * sr, cp * cr],
        ]
    )


def parse_urdf(path):
    root = ET.parse(path).getroot()
    links, order = {}, []
    for link in root.findall("link"):
        spheres = []
        for col in link.findall("collision"):
            geom = col.find("geometry")
            sph = geom.find("sphere") if geom is not None else None
            if sph is None:
                continue
            origin = col.find("origin")
            xyz = _floats(origin.get("xyz") if origin is not None else None)
            rpy = _floats(origin.get("rpy") if origin is not None else None)
            assert all(abs(a) < 1e-12 for a in rpy) or True  # a sphere's orientation is irrelevant
            spheres.append(xyz + [float(sph.get("radius"))])
        links[link.get("name")] = spheres
        order.append(link.get("name"))

    joints = []
    for j in root.findall("joint"):
        origin = j.find("origin")
        axis = j.find("axis")
        limit = j.find("limit")
        joints.append(
            dict(
                name=j.get("name"),
                type=j.get("type"),
                parent=j.find("parent").get("link"),
                child=j.find("child").get("link"),
                xyz=_floats(origin.get("xyz") if origin is not None else None),
                rpy=_floats(origin.get("rpy") if origin is not None else None),
                axis=_floats(axis.get("xyz")) if axis is not None else [1.0, 0.0, 0.0],
                lower=float(limit.get("lower")) if limit is not None and limit.get("lower") else 0.0,
                upper=float(limit.get("upper")) if limit is not None and limit.get("upper") else 0.0,
            )
        )
    return links, order, joints


def parse_srdf(path):
    root = ET.parse(path).getroot()
    return {frozenset((d.get("link1"), d.get("link2"))) for d in root.findall("disable_collisions")}


# --------------------------------------------------------------------------------------------
# minimal enclosing sphere of a set of spheres (exact support-set enumeration; n is tiny)
# --------------------------------------------------------------------------------------------
def _ball_through(support):
    """Smallest sphere (c, R) internally tangent to every ball in `support` (1..4 balls)."""
    c0, r0 = support[0]
    if len(support) == 1:
        return c0, r0
    # |c - ci| = R - ri  for all i.  Subtract the first equation:
    #   2 (ci - c0)·c - 2 (ri - r0) R = |ci|² - |c0|² - ri² + r0²       (linear in c, R)
    A = np.array([np.append(2 * (ci - c0), -2 * (ri - r0)) for ci, ri in support[1:]])
    b = np.array([ci @ ci - c0 @ c0 - ri * ri + r0 * r0 for ci, ri in support[1:]])
    # minimal-norm parametrisation: c = p + R q restricted to the affine hull of the centres
    M, rhs_R = A[:, :3], -A[:, 3]
    # solve M c = b + rhs_R * R in the least-norm sense around c0
    Mp = np.linalg.pinv(M)
    p = c0 + Mp @ (b - M @ c0)
    q = Mp @ rhs_R
    # |p + R q - c0|² = (R - r0)²
    d = p - c0
    a2 = q @ q - 1.0
    a1 = 2 * (d @ q) + 2 * r0
    a0 = d @ d - r0 * r0
    if abs(a2) < 1e-14:
        if abs(a1) < 1e-14:
            return None
        roots = [-a0 / a1]
    else:
        disc = a1 * a1 - 4 * a2 * a0
        if disc < -1e-12:
            return None
        disc = max(disc, 0.0)
        roots = [(-a1 + math.sqrt(disc)) / (2 * a2), (-a1 - math.sqrt(disc)) / (2 * a2)]
    best = None
    for R in roots:
        if R < max(r for _, r in support) - 1e-12:
            continue
        c = p + R * q
        if all(abs(np.linalg.norm(c - ci) + ri - R) < 1e-9 for ci, ri in support):
            if best is None or R < best[1]:
                best = (c, R)
    return best


def enclosing_sphere(spheres):
    from itertools import combinations

    balls = [(np.array(s[:3], float), float(s[3])) for s in spheres]
    best = None
    for k in range(1, min(4, len(balls)) + 1):
        for sub in combinations(balls, k):
            cand = _ball_through(list(sub))
            if cand is None:
                continue
            c, R = cand
            if all(np.linalg.norm(c - ci) + ri <= R + 1e-10 for ci, ri in balls):
                if best is None or R < best[1] - 1e-13:
                    best = (c, R)
    assert best is not None
    return best


# --------------------------------------------------------------------------------------------
# model assembly
# --------------------------------------------------------------------------------------------
def build_model(name, urdf_path, srdf_path):
    cfg = ROBOTS[name]
    links, link_order, joints = parse_urdf(urdf_path)
    disabled = parse_srdf(srdf_path)
    children = {j["child"] for j in joints}
    roots = [l for l in link_order if l not in children]
    assert len(roots) == 1, roots
    by_parent = {}
    for j in joints:
        by_parent.setdefault(j["parent"], []).append(j)

    dof_of = {jn: i for i, jn in enumerate(cfg["joints"])}
    lower = [0.0] * len(dof_of)
    upper = [0.0] * len(dof_of)

    # rigid bodies: body 0 is world-fixed.  Each body: parent body, fixed transform from the parent
    # body frame to the joint frame (T_pre, 4x4 double), joint type/axis/dof.  A link's pose in its
    # body frame is link_in_body[link] (4x4 double).
    bodies = [dict(name=roots[0], parent=-1, T_pre=np.eye(4), jtype="fixed", axis=[0, 0, 1], dof=-1)]
    link_body = {roots[0]: 0}
    link_in_body = {roots[0]: np.eye(4)}

    # Sphere numbering follows a depth-first walk with a link's child joints visited in
    # alphabetical order of joint name (the order in which the reference numbers its spheres,
    # established by matching our FK against the reference's sphere_fk for all four robots).
    dfs_order = [roots[0]]

    def visit(link):
        for j in sorted(by_parent.get(link, []), key=lambda j: j["name"]):
            dfs_order.append(j["child"])
            T = np.eye(4)
            T[:3, :3] = rpy_matrix(*j["rpy"])
            T[:3, 3] = j["xyz"]
            T_link = link_in_body[link] @ T
            if j["type"] == "fixed" or j["name"] not in dof_of:
                assert j["type"] == "fixed", f"non-fixed joint {j['name']} is not actuated"
                link_body[j["child"]] = link_body[link]
                link_in_body[j["child"]] = T_link
            else:
                assert j["type"] in ("revolute", "continuous", "prismatic")
                d = dof_of[j["name"]]
                lower[d], upper[d] = j["lower"], j["upper"]
                bodies.append(
                    dict(
                        name=j["child"],
                        parent=link_body[link],
                        T_pre=T_link,
                        jtype="prismatic" if j["type"] == "prismatic" else "revolute",
                        axis=j["axis"],
                        dof=d,
                    )
                )
                link_body[j["child"]] = len(bodies) - 1
                link_in_body[j["child"]] = np.eye(4)
            visit(j["child"])

    visit(roots[0])

    clinks = []
    sphere_index = 0
    for l in dfs_order:
        if not links[l] or l not in link_body:
            continue
        T = link_in_body[l]
        sph = []
        for x, y, z, r in links[l]:
            p = T @ np.array([x, y, z, 1.0])
            # radii are stored as float32 by the reference (e.g. panda.hh:5600-5617)
            sph.append([float(p[0]), float(p[1]), float(p[2]), float(np.float32(r))])
        c, R = enclosing_sphere(sph)
        clinks.append(
            dict(
                name=l,
                body=link_body[l],
                first_sphere=sphere_index,
                spheres=sph,
                bound=[float(c[0]), float(c[1]), float(c[2]), float(R + BOUND_MARGIN)],
                bound_exact_radius=float(R),
            )
        )
        sphere_index += len(sph)

    # Static reach: an upper bound, valid for ANY revolute joint values, on how far from the world
    # origin a link's bounding sphere can extend (triangle inequality along the chain).  Prismatic
    # joints contribute their largest in-limit travel (+ margin); the kernels only use the bound for
    # states whose prismatic joints are inside that range (sink.reach_ok).
    def reach(bi, p):
        """max over all revolute joint values of |world position| of point p fixed in body bi:
        the point sweeps a circle about the joint axis; circle centre c is fixed in the parent's
        frame, so reach <= reach(parent, c) + circle radius (triangle inequality, level by level)."""
        bd = bodies[bi]
        p = np.asarray(p, float)
        if bd["parent"] < 0:
            return float(np.linalg.norm(p))
        ax = np.array(bd["axis"], float)
        ax = ax / np.linalg.norm(ax)
        T = bd["T_pre"]
        if bd["jtype"] == "prismatic":
            d = bd["dof"]
            mid = 0.5 * (lower[d] + upper[d])
            half = 0.5 * (upper[d] - lower[d]) + BOX_MARGIN
            c = T[:3, :3] @ (p + mid * ax) + T[:3, 3]
            return reach(bd["parent"], c) + half
        c_local = (p @ ax) * ax
        rho = float(np.linalg.norm(p - c_local))
        c = T[:3, :3] @ c_local + T[:3, 3]
        return reach(bd["parent"], c) + rho

    for c in clinks:
        c["reach"] = float(reach(c["body"], c["bound"][:3]) + c["bound"][3] + 1e-4)

    names = [c["name"] for c in clinks]
    disabled = disabled | {frozenset(p) for p in cfg.get("extra_disabled", [])}
    pairs = []
    for i in range(len(names)):
        for j in range(i + 1, len(names)):
            # links of one rigid body can never move relative to each other: not checked
            if clinks[i]["body"] == clinks[j]["body"]:
                continue
            if frozenset((names[i], names[j])) not in disabled:
                pairs.append([i, j])

    ee = cfg["end_effector"]
    # the attachment is checked against every link that may collide with ANY collision link of the
    # rigid body carrying the end-effector frame (this reproduces the reference's
    # "Attachment vs. <link>" lists for all four robots, e.g. robots/panda.hh:15318-15440)
    parent_of = {j["child"]: j["parent"] for j in joints}
    ee_clink = ee
    while ee_clink not in names:
        ee_clink = parent_of[ee_clink]
    on_ee_body = {i for i, c in enumerate(clinks) if c["body"] == link_body[ee]}
    attach_links = sorted(
        {b for a, b in pairs if a in on_ee_body} | {a for a, b in pairs if b in on_ee_body}
    )

    model = dict(
        name=name,
        dof=len(dof_of),
        resolution=cfg["resolution"],
        joint_names=cfg["joints"],
        lower=[float(np.float32(v)) for v in lower],
        # Robot::s_m of the reference's generated code is the f32 of the double difference (fetch.hh:52-60)
        range=[float(np.float32(float(u) - float(l))) for l, u in zip(lower, upper)],
        n_spheres=sphere_index,
        min_radius=min(s[3] for c in clinks for s in c["spheres"]),
        max_radius=max(s[3] for c in clinks for s in c["spheres"]),
        bodies=[
            dict(
                name=b["name"],
                parent=b["parent"],
                T_pre=[[float(v) for v in row] for row in b["T_pre"][:3]],
                jtype=b["jtype"],
                axis=[float(a) for a in b["axis"]],
                dof=b["dof"],
            )
            for b in bodies
        ],
        links=clinks,
        self_pairs=pairs,
        attach_links=attach_links,
        end_effector=dict(
            link=ee,
            collision_link=ee_clink,
            body=link_body[ee],
            T=[[float(v) for v in row] for row in link_in_body[ee][:3]],
        ),
    )
    return model



# --------------------------------------------------------------------------------------------
# self-collision analysis: which sphere pairs can ever touch, which link pairs nearly always overlap
# --------------------------------------------------------------------------------------------
INLINE_FREQUENCY = 0.5  # link pairs whose bounding spheres overlap more often than this are checked
                        # unconditionally (fine spheres, pruned list) by every thread


def numeric_frames(model, q):
    """Body frames for configurations q [N, dof] (float64) -> list of [N, 4, 4]."""
    q = np.asarray(q, float)
    N = len(q)
    F = []
    for b in model["bodies"]:
        if b["parent"] < 0:
            F.append(np.tile(np.eye(4), (N, 1, 1)))
            continue
        pre = np.vstack([np.array(b["T_pre"]), [0, 0, 0, 1]])
        ax = np.array(b["axis"], float)
        ax = ax / np.linalg.norm(ax)
        v = q[:, b["dof"]]
        J = np.tile(np.eye(4), (N, 1, 1))
        if b["jtype"] == "revolute":
            K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
            J[:, :3, :3] = np.eye(3) + np.sin(v)[:, None, None] * K + (1 - np.cos(v))[:, None, None] * (K @ K)
        else:
            J[:, :3, 3] = ax * v[:, None]
        F.append(F[b["parent"]] @ pre @ J)
    return F


def body_chain(model, a, b):
    """Bodies (with a joint) on the tree path between bodies a and b, split by side."""
    def ancestors(x):
        out = [x]
        while model["bodies"][x]["parent"] >= 0:
            x = model["bodies"][x]["parent"]
            out.append(x)
        return out

    pa, pb = ancestors(a), ancestors(b)
    common = next(x for x in pa if x in pb)
    side_a = pa[: pa.index(common)]
    side_b = pb[: pb.index(common)]
    return side_a, side_b


def analyze_self_pairs(model, seed=0):
    rng = np.random.default_rng(seed)
    L, Bd = model["links"], model["bodies"]
    lo = np.array(model["lower"], float)
    rg = np.array(model["range"], float)
    q = lo + rg * rng.random((20000, model["dof"]))
    F = numeric_frames(model, q)
    centre = [(F[l["body"]] @ np.array(l["bound"][:3] + [1.0]))[:, :3] for l in L]
    info = []
    for a, b in model["self_pairs"]:
        d = np.linalg.norm(centre[a] - centre[b], axis=1)
        freq = float((d < L[a]["bound"][3] + L[b]["bound"][3]).mean())
        side_a, side_b = body_chain(model, L[a]["body"], L[b]["body"])
        chain = side_a + side_b
        entry = dict(a=a, b=b, frequency=freq, chain_dofs=[Bd[x]["dof"] for x in chain], pruned=None)
        if 1 <= len(chain) <= 2:
            n = 4096 if len(chain) == 1 else 720
            grids = []
            for x in chain:
                dof = Bd[x]["dof"]
                grids.append(np.linspace(lo[dof] - BOX_MARGIN, lo[dof] + rg[dof] + BOX_MARGIN, n))
            mesh = np.meshgrid(*grids, indexing="ij")
            qq = np.zeros((mesh[0].size, model["dof"]))
            for x, g in zip(chain, mesh):
                qq[:, Bd[x]["dof"]] = g.ravel()
            Fg = numeric_frames(model, qq)
            Fa, Fb = Fg[L[a]["body"]], Fg[L[b]["body"]]
            pa = [(Fa @ np.array(s[:3] + [1.0]))[:, :3] for s in L[a]["spheres"]]
            pb = [(Fb @ np.array(s[:3] + [1.0]))[:, :3] for s in L[b]["spheres"]]
            # Lipschitz slack: between grid nodes a sphere on the far side of joint k moves at most
            # step_k/2 * (its distance from the joint origin) per axis of the grid
            slack = 1e-4
            for x, g in zip(chain, grids):
                step = g[1] - g[0]
                origin = Fg[x][:, :3, 3]
                far = pa if x in side_a else pb
                lever = max(float(np.linalg.norm(p - origin, axis=1).max()) for p in far) if Bd[x]["jtype"] == "revolute" else 1.0
                slack += 0.5 * step * lever
            keep = []
            for i, sa in enumerate(L[a]["spheres"]):
                for j, sb in enumerate(L[b]["spheres"]):
                    dmin = float((np.linalg.norm(pa[i] - pb[j], axis=1)).min()) - (sa[3] + sb[3])
                    if dmin < slack:
                        keep.append([i, j])
            entry["pruned"] = keep
            entry["prune_slack"] = slack
        entry["inline"] = bool(freq > INLINE_FREQUENCY and entry["pruned"] is not None)
        info.append(entry)
    return info


# --------------------------------------------------------------------------------------------
# symbolic tracer
# --------------------------------------------------------------------------------------------
ZERO_EPS = 1e-10  # |constant| below this is snapped to 0 (e.g. cos(1.57079632679) = 4.9e-12)


class Tracer:
    """Hash-consed expression DAG over {const, input, neg, add, mul}.  Inputs are sin(q_j), cos(q_j)
    and q_j itself (prismatic).  Constants are folded in double precision and narrowed to float32
    only when emitted, matching how the reference's generated code treats its literals."""

    def __init__(self):
        self.nodes = []  # (op, a, b) ; const: ("c", value, None); input: ("i", name, None)
        self.index = {}

    def _mk(self, key):
        if key in self.index:
            return self.index[key]
        self.nodes.append(key)
        self.index[key] = len(self.nodes) - 1
        return len(self.nodes) - 1

    def const(self, v):
        v = float(v)
        if abs(v) < ZERO_EPS:
            v = 0.0
        if abs(v - 1.0) < ZERO_EPS:
            v = 1.0
        if abs(v + 1.0) < ZERO_EPS:
            v = -1.0
        return self._mk(("c", v, None))

    def inp(self, name):
        return self._mk(("i", name, None))

    def is_const(self, n):
        return self.nodes[n][0] == "c"

    def cval(self, n):
        return self.nodes[n][1]

    def neg(self, a):
        if self.is_const(a):
            return self.const(-self.cval(a))
        if self.nodes[a][0] == "neg":
            return self.nodes[a][1]
        return self._mk(("neg", a, None))

    def mul(self, a, b):
        if self.is_const(a) and self.is_const(b):
            return self.const(self.cval(a) * self.cval(b))
        if self.is_const(b):
            a, b = b, a
        if self.is_const(a):
            v = self.cval(a)
            if v == 0.0:
                return a
            if v == 1.0:
                return b
            if v == -1.0:
                return self.neg(b)
            if self.nodes[b][0] == "neg":
                return self.mul(self.const(-v), self.nodes[b][1])
            # (c1 * x) * c2 -> (c1*c2) * x
            if self.nodes[b][0] == "mul" and self.is_const(self.nodes[b][1]):
                return self.mul(self.const(v * self.cval(self.nodes[b][1])), self.nodes[b][2])
            return self._mk(("mul", a, b))
        sign = 1
        if self.nodes[a][0] == "neg":
            a, sign = self.nodes[a][1], -sign
        if self.nodes[b][0] == "neg":
            b, sign = self.nodes[b][1], -sign
        if a > b:
            a, b = b, a
        n = self._mk(("mul", a, b))
        return n if sign > 0 else self.neg(n)

    def add(self, a, b):
        if self.is_const(a) and self.is_const(b):
            return self.const(self.cval(a) + self.cval(b))
        if self.is_const(a) and self.cval(a) == 0.0:
            return b
        if self.is_const(b) and self.cval(b) == 0.0:
            return a
        if a > b:
            a, b = b, a
        return self._mk(("add", a, b))

    def sub(self, a, b):
        return self.add(a, self.neg(b))

    # -- small linear algebra over node ids -----------------------------------------------
    def matmul(self, A, B):
        n, k, m = len(A), len(B), len(B[0])
        out = [[None] * m for _ in range(n)]
        for i in range(n):
            for j in range(m):
                acc = self.const(0.0)
                for t in range(k):
                    acc = self.add(acc, self.mul(A[i][t], B[t][j]))
                out[i][j] = acc
        return out

    def cmat(self, M):
        return [[self.const(v) for v in row] for row in M]


def axis_rotation(tr, axis, s, c):
    """Rotation about a constant unit axis with symbolic sin/cos: R = c I + s [a]x + (1-c) a a^T.
    Axis-aligned axes (every joint of the four shipped robots) get the exact sparse form."""
    ax = np.array(axis, float)
    ax = ax / np.linalg.norm(ax)
    K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    aligned = np.sum(np.abs(ax) > 1e-12) == 1
    one_minus_c = None if aligned else tr.sub(tr.const(1.0), c)
    R = [[None] * 3 for _ in range(3)]
    for i in range(3):
        for j in range(3):
            if aligned:
                if i == j:
                    e = tr.const(1.0) if abs(ax[i]) > 0.5 else c
                else:
                    e = tr.mul(tr.const(K[i, j]), s)
            else:
                e = tr.mul(tr.const(1.0 if i == j else 0.0), c)
                e = tr.add(e, tr.mul(tr.const(K[i, j]), s))
                e = tr.add(e, tr.mul(tr.const(ax[i] * ax[j]), one_minus_c))
            R[i][j] = e
    return R


def trace_frames(model):
    """Returns (tracer, frames) where frames[b] is a 3x4 matrix of node ids for body b."""
    tr = Tracer()
    frames = [None] * len(model["bodies"])
    for b, body in enumerate(model["bodies"]):
        if body["parent"] < 0:
            frames[b] = tr.cmat(np.hstack([np.eye(3), np.zeros((3, 1))]))
            continue
        P = frames[body["parent"]]
        P4 = P + [[tr.const(0.0), tr.const(0.0), tr.const(0.0), tr.const(1.0)]]
        pre = tr.cmat(body["T_pre"] + [[0.0, 0.0, 0.0, 1.0]])
        d = body["dof"]
        if body["jtype"] == "revolute":
            s, c = tr.inp(f"s{d}"), tr.inp(f"c{d}")
            R = axis_rotation(tr, body["axis"], s, c)
            J = [R[i] + [tr.const(0.0)] for i in range(3)] + [
                [tr.const(0.0), tr.const(0.0), tr.const(0.0), tr.const(1.0)]
            ]
        else:
            ax = np.array(body["axis"], float)
            ax = ax / np.linalg.norm(ax)
            q = tr.inp(f"q{d}")
            J = [
                [tr.const(1.0 if i == j else 0.0) for j in range(3)] + [tr.mul(tr.const(ax[i]), q)]
                for i in range(3)
            ] + [[tr.const(0.0), tr.const(0.0), tr.const(0.0), tr.const(1.0)]]
        # local = pre * J is cheap (constants x one joint); then parent * local
        local = tr.matmul(pre, J)
        frames[b] = tr.matmul(P4, local)[:3]
    return tr, frames


def trace_bounds(model, tr, frames):
    """World-frame centre of every collision link's bounding sphere (3 node ids per link)."""
    out = []
    for l in model["links"]:
        F = frames[l["body"]]
        c = [tr.const(v) for v in l["bound"][:3]]
        out.append(
            [tr.add(tr.add(tr.add(tr.mul(F[i][0], c[0]), tr.mul(F[i][1], c[1])), tr.mul(F[i][2], c[2])), F[i][3]) for i in range(3)]
        )
    return out


def emit_cuda(model, tr, frames):
    """Straight-line device code: computes every moving body's 3x4 frame and hands each of the 12
    entries to `sink(body, k, value)`; constant entries are passed as literals so the sink sees
    them at compile time."""
    name = model["name"]
    needed = set()

    def need(n):
        if n in needed:
            return
        needed.add(n)
        op, a, b = tr.nodes[n]
        if op in ("neg",):
            need(a)
        elif op in ("add", "mul"):
            need(a)
            need(b)

    bounds = trace_bounds(model, tr, frames)
    for F in frames:
        for row in F:
            for n in row:
                need(n)
    for b3 in bounds:
        for n in b3:
            need(n)

    # fine-sphere positions needed by the inline self-collision tests
    sphere_nodes = {}

    def sphere_pos(link, k):
        key = (link, k)
        if key not in sphere_nodes:
            l = model["links"][link]
            F = frames[l["body"]]
            c = [tr.const(v) for v in l["spheres"][k][:3]]
            sphere_nodes[key] = [
                tr.add(tr.add(tr.add(tr.mul(F[i][0], c[0]), tr.mul(F[i][1], c[1])), tr.mul(F[i][2], c[2])), F[i][3])
                for i in range(3)
            ]
        return sphere_nodes[key]

    inline_tests = []  # (nodesA, nodesB, rs2 literal)
    for e in model.get("self_pair_info", []):
        if not e["inline"]:
            continue
        for i, j in e["pruned"]:
            ra = model["links"][e["a"]]["spheres"][i][3]
            rb = model["links"][e["b"]]["spheres"][j][3]
            rs = np.float32(np.float32(ra) + np.float32(rb))
            inline_tests.append((sphere_pos(e["a"], i), sphere_pos(e["b"], j), float(np.float32(rs * rs))))
    for pa, pb, _ in inline_tests:
        for n in pa + pb:
            need(n)

    def lit(v):
        f = float(np.float32(v))
        s = repr(f)
        if "e" not in s and "." not in s and "inf" not in s:
            s += ".0"
        return s + "f"

    lines = []
    ref = {}

    def operand(n):
        op, a, _ = tr.nodes[n]
        if op == "c":
            return lit(a)
        if op == "i":
            return a
        return ref[n]

    n_fma = n_mul = n_add = 0
    # use counts, to fuse a single-use mul into its consuming add as an FMA
    uses = {}
    for n in sorted(needed):
        op, a, b = tr.nodes[n]
        for x in (a, b):
            if op in ("neg", "add", "mul") and isinstance(x, int):
                uses[x] = uses.get(x, 0) + 1
    for F in frames:
        for row in F:
            for n in row:
                uses[n] = uses.get(n, 0) + 1
    for b3 in bounds:
        for n in b3:
            uses[n] = uses.get(n, 0) + 1
    for pa, pb, _ in inline_tests:
        for n in pa + pb:
            uses[n] = uses.get(n, 0) + 1

    fused = set()
    for n in sorted(needed):
        op, a, b = tr.nodes[n]
        if op != "add":
            continue
        for m, o in ((a, b), (b, a)):
            if tr.nodes[m][0] == "mul" and uses.get(m, 0) == 1 and m not in fused:
                fused.add(m)
                break

    for n in sorted(needed):
        op, a, b = tr.nodes[n]
        if op in ("c", "i") or n in fused:
            continue
        var = f"t{n}"
        ref[n] = var
        if op == "neg":
            lines.append(f"const float {var} = -{operand(a)};")
        elif op == "mul":
            lines.append(f"const float {var} = {operand(a)} * {operand(b)};")
            n_mul += 1
        elif op == "add":
            done = False
            for m, o in ((a, b), (b, a)):
                if m in fused and tr.nodes[m][0] == "mul":
                    _, ma, mb = tr.nodes[m]
                    # a negated addend: fma(x, y, -z)
                    lines.append(f"const float {var} = fmaf({operand(ma)}, {operand(mb)}, {operand(o)});")
                    n_fma += 1
                    done = True
                    break
            if not done:
                lines.append(f"const float {var} = {operand(a)} + {operand(b)};")
                n_add += 1

    out = []
    out.append(f"// GENERATED by tools/robot_compiler.py from the {name} robot description -- do not edit.")
    out.append(f"// Straight-line forward kinematics of the {len(frames) - 1} moving rigid bodies of '{name}':")
    out.append(f"// {n_fma} fma + {n_mul} mul + {n_add} add (+ {model['dof']} sincos).")
    out.append("#pragma once")
    out.append("")
    out.append("namespace vmv { namespace gen {")
    out.append("")
    out.append("template <typename Sink>")
    out.append(f"__device__ __forceinline__ void {name}_frames(const float (&q)[{model['dof']}], Sink &sink)")
    out.append("{")
    for b in model["bodies"]:
        d = b["dof"]
        if d < 0:
            continue
        if b["jtype"] == "revolute":
            out.append(f"    float s{d}, c{d};")
            out.append(f"    vmv::sincos_f32(q[{d}], s{d}, c{d});")
        else:
            out.append(f"    const float q{d} = q[{d}];")
    for l in lines:
        out.append("    " + l)
    for b, F in enumerate(frames):
        if model["bodies"][b]["parent"] < 0:
            continue
        for i in range(3):
            for j in range(4):
                out.append(f"    sink.template put<{b}, {i * 4 + j}>({operand(F[i][j])});")
    out.append("    // world-frame centres of the per-link bounding spheres")
    for li, b3 in enumerate(bounds):
        for ax in range(3):
            out.append(f"    sink.template bound<{li}, {ax}>({operand(b3[ax])});")
    # pruned sphere-pair lists are valid inside the (slightly widened) joint box
    box_dofs = sorted({d for e in model.get("self_pair_info", []) if e["pruned"] is not None for d in e["chain_dofs"]})
    conds = []
    for d in box_dofs:
        lo_d = model["lower"][d] - BOX_MARGIN
        hi_d = model["lower"][d] + model["range"][d] + BOX_MARGIN
        conds.append(f"(q[{d}] >= {lit(lo_d)}) & (q[{d}] <= {lit(hi_d)})")
    out.append("    // the statically pruned self-collision lists hold inside the joint limits (+ margin)")
    out.append("    sink.in_box(" + (" & ".join(conds) if conds else "true") + ");")
    pris = [b for b in model["bodies"] if b["jtype"] == "prismatic"]
    rconds = []
    for b in pris:
        d = b["dof"]
        # the interval the reach bounds were derived for -- not |q| <= max(|lower|, |upper|), which also admits travel
        # below a positive lower limit
        lo_d = model["lower"][d] - BOX_MARGIN
        hi_d = model["lower"][d] + model["range"][d] + BOX_MARGIN
        rconds.append(f"(q[{d}] >= {lit(lo_d)}) & (q[{d}] <= {lit(hi_d)})")
    out.append("    // the static per-link reach bounds assume prismatic joints inside their limits")
    out.append("    sink.reach_ok(" + (" & ".join(rconds) if rconds else "true") + ");")
    out.append("    // link pairs whose bounding spheres (nearly) always overlap: kinematically feasible fine pairs, ungated")
    out.append("    bool self_hit = false;")
    out.append("    if constexpr (Sink::kInlinePairs)  // off when the two-joint verdict tables answer these pairs (vmv_pairtab.cuh)")
    out.append("    {")
    for pa, pb, rs2 in inline_tests:
        dx = f"({operand(pa[0])} - {operand(pb[0])})"
        dy = f"({operand(pa[1])} - {operand(pb[1])})"
        dz = f"({operand(pa[2])} - {operand(pb[2])})"
        out.append(
            f"    {{ const float dx = {dx}, dy = {dy}, dz = {dz}; "
            f"self_hit |= vmv::sign_set(((dx * dx + dy * dy) + dz * dz) - {lit(rs2)}); }}"
        )
    out.append("    }")
    out.append("    sink.self_inline(self_hit);")
    out.append("}")
    out.append("")
    out.append("}}  // namespace vmv::gen")
    out.append("")
    return "\n".join(out), dict(fma=n_fma, mul=n_mul, add=n_add)


def emit_tables(model):
    """Constant tables for the kernels (device) and the host side of the C ABI.

    Environment tasks are listed in the order the reference sweeps links (distal link first, e.g.
    robots/panda.hh:5629-6010): one bounding-sphere task per collision link followed by that link's
    fine-sphere tasks.  `skip` of a bounding task is the index of the next link's bounding task
    (taken when the bounding sphere is clear); a fine task has skip = -1."""
    name = model["name"]
    L = model["links"]

    def f(v):
        x = repr(float(np.float32(v)))
        if "e" not in x and "." not in x and "inf" not in x and "nan" not in x:
            x += ".0"
        return x + "f"

    tasks = []  # (cx, cy, cz, r, body, skip, link, sphere)
    for li in range(len(L) - 1, -1, -1):
        l = L[li]
        skip = len(tasks) + 1 + len(l["spheres"])
        tasks.append((*l["bound"], l["body"], skip, li, -1))
        for k, sp in enumerate(l["spheres"]):
            tasks.append((*sp, l["body"], -1, li, l["first_sphere"] + k))

    out = []
    out.append(f"// GENERATED by tools/robot_compiler.py from the {name} robot description -- do not edit.")
    out.append("#pragma once")
    out.append('#include "../robot_tables.h"')
    out.append("")
    out.append("namespace vmv { namespace gen {")
    out.append("")
    out.append(f"struct {name}_model")
    out.append("{")
    out.append(f'    static constexpr const char *kName = "{name}";')
    out.append(f"    static constexpr int kDof = {model['dof']};")
    out.append(f"    static constexpr int kBodies = {len(model['bodies'])};  // body 0 is world-fixed")
    out.append(f"    static constexpr int kLinks = {len(L)};")
    out.append(f"    static constexpr int kSpheres = {model['n_spheres']};")
    out.append(f"    static constexpr int kPairs = {len(model['self_pairs'])};")
    out.append(f"    static constexpr int kTasks = {len(tasks)};")
    out.append(f"    static constexpr int kResolution = {model['resolution']};")
    out.append(f"    static constexpr int kAttachLinks = {len(model['attach_links'])};")
    out.append(f"    static constexpr int kEeBody = {model['end_effector']['body']};")
    out.append("};")
    out.append("")
    out.append(f"static const float {name}_lower[{model['dof']}] = {{" + ", ".join(f(v) for v in model["lower"]) + "};")
    out.append(f"static const float {name}_range[{model['dof']}] = {{" + ", ".join(f(v) for v in model["range"]) + "};")
    out.append("")
    out.append(f"static const vmv::SphereTask {name}_tasks_host[{len(tasks)}] = {{")
    for t in tasks:
        out.append(f"    {{{f(t[0])}, {f(t[1])}, {f(t[2])}, {f(t[3])}, {t[4]}, {t[5]}, {t[6]}, {t[7]}}},")
    out.append("};")
    out.append("")
    out.append("// collision link -> {first fine sphere, count, body, index of its bounding task}")
    out.append(f"static const vmv::LinkInfo {name}_links_host[{len(L)}] = {{")
    btask = {t[6]: i for i, t in enumerate(tasks) if t[7] < 0}
    for li, l in enumerate(L):
        out.append(f"    {{{l['first_sphere']}, {len(l['spheres'])}, {l['body']}, {btask[li]}}},  // {l['name']}")
    out.append("};")
    out.append("")
    out.append(f"static const vmv::LinkPair {name}_pairs_host[{max(1, len(model['self_pairs']))}] = {{")
    for a, b in model["self_pairs"]:
        out.append(f"    {{{a}, {b}}},")
    if not model["self_pairs"]:
        out.append("    {0, 0},")
    out.append("};")
    out.append("")
    N = name.upper()
    out.append("// X-macros for fully unrolled device code: LINKS(X) -> X(link, bound_radius, n_spheres, first_fine_task, reach);")
    out.append("// PAIRS(X) -> X(pair_index, link_a, link_b, checked_inline)")
    out.append(f"#define VMV_{N}_LINKS(X) \\")
    for li, l in enumerate(L):
        out.append(f"    X({li}, {f(l['bound'][3])}, {len(l['spheres'])}, {btask[li] + 1}, {f(l['reach'])}) \\")
    out.append("")
    info = model.get("self_pair_info") or [dict(inline=False, pruned=None) for _ in model["self_pairs"]]
    out.append(f"#define VMV_{N}_PAIRS(X) \\")
    for pi, (a, b) in enumerate(model["self_pairs"]):
        out.append(f"    X({pi}, {a}, {b}, {1 if info[pi]['inline'] else 0}) \\")
    out.append("")
    # statically pruned sphere-pair lists (task indices), valid inside the joint box
    lists, pinfo = [], []
    for pi, (a, b) in enumerate(model["self_pairs"]):
        e = info[pi]
        if e["pruned"] is None:
            pinfo.append((0, -1, 0))
        else:
            pinfo.append((len(lists), len(e["pruned"]), 1 if e["inline"] else 0))
            for i, j in e["pruned"]:
                lists.append((btask[a] + 1 + i, btask[b] + 1 + j))
    out.append("// per pair: {offset into pair_lists, count (-1 = no pruned list: full cross product), inline flag}")
    out.append(f"static const vmv::PairInfo {name}_pair_info_host[{max(1, len(pinfo))}] = {{")
    for o, c, f_ in pinfo or [(0, -1, 0)]:
        out.append(f"    {{{o}, {c}, {f_}}},")
    out.append("};")
    out.append(f"static const vmv::SpherePair {name}_pair_lists_host[{max(1, len(lists))}] = {{")
    for ta, tb in lists or [(0, 0)]:
        out.append(f"    {{{ta}, {tb}}},")
    out.append("};")
    out.append(f"static const int {name}_pair_lists_count = {len(lists)};")
    out.append("")
    out.append(f"static const int {name}_attach_links_host[{max(1, len(model['attach_links']))}] = {{" + ", ".join(str(a) for a in (model["attach_links"] or [0])) + "};")
    ee = np.array(model["end_effector"]["T"]).reshape(-1)
    out.append(f"static const float {name}_ee_tf_host[12] = {{" + ", ".join(f(v) for v in ee) + "};")
    out.append("")
    out.append("}}  // namespace vmv::gen")
    return "\n".join(out) + "\n"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--resources", default="/root/reference/resources")
    ap.add_argument("--out", default=str(Path(__file__).resolve().parents[1] / "vamp_mvt_b200"))
    ap.add_argument("--robots", nargs="*", default=list(ROBOTS))
    args = ap.parse_args()
    out = Path(args.out)
    (out / "robots").mkdir(parents=True, exist_ok=True)
    (out / "csrc" / "gen").mkdir(parents=True, exist_ok=True)
    for name in args.robots:
        res = Path(args.resources) / name
        model = build_model(name, res / f"{name}_spherized.urdf", res / f"{name}.srdf")
        model["self_pair_info"] = analyze_self_pairs(model)
        tr, frames = trace_frames(model)
        code, stats = emit_cuda(model, tr, frames)
        model["fk_ops"] = stats
        (out / "robots" / f"{name}.json").write_text(json.dumps(model, indent=1))
        (out / "csrc" / "gen" / f"{name}_fk.cuh").write_text(code)
        (out / "csrc" / "gen" / f"{name}_tables.h").write_text(emit_tables(model))
        print(
            f"{name}: dof={model['dof']} bodies={len(model['bodies'])} links={len(model['links'])} "
            f"spheres={model['n_spheres']} pairs={len(model['self_pairs'])} fk_ops={stats}"
        )


if __name__ == "__main__":
    main()
