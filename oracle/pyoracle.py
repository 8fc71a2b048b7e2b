"""ctypes front-ends of the two CPU checkers.  TEST INFRASTRUCTURE ONLY.

* ``Oracle``  -- oracle/vamp_oracle.c, our scalar C restatement of the reference's validation path
  (built by ``build()`` with gcc; travels to the GPU box as oracle/libvamp_oracle.so).
* ``Ref``     -- oracle/_ref/libvamp_ref.so, the UNMODIFIED reference headers compiled in place
  (recipe oracle/ref/Makefile; buildable only where /root/reference is mounted; the built .so
  travels to the GPU box).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  The product package (vamp_mvt_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REPO = HERE.parent
ORACLE_SO = HERE / "libvamp_oracle.so"
REF_SO = HERE / "_ref" / "libvamp_ref.so"
REFERENCE_ROOT = Path("/root/reference")
ROBOT_IDS = {"panda": 0, "ur5": 1, "fetch": 2, "baxter": 3}


def build(verbose: bool = False) -> None:
    """Compile the C restatement; compile oracle/_ref when the reference sources are mounted."""
    src = HERE / "vamp_oracle.c"
    if not ORACLE_SO.exists() or ORACLE_SO.stat().st_mtime < src.stat().st_mtime:
        cmd = ["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-o", str(ORACLE_SO), str(src), "-lm"]
        subprocess.run(cmd, check=True)
    if REFERENCE_ROOT.exists():
        out = subprocess.run(
            ["make", "-j8", "-C", str(HERE / "ref"), f"REF={REFERENCE_ROOT}"],
            capture_output=not verbose,
            text=True,
        )
        if out.returncode != 0:
            raise RuntimeError(f"oracle/_ref build failed:\n{out.stdout}\n{out.stderr}")


def _fp(a):
    return a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def load_model(robot: str) -> dict:
    return json.loads((REPO / "vamp_mvt_b200" / "robots" / f"{robot}.json").read_text())


# ------------------------------------------------------------------------------------------
class _OrModel(C.Structure):
    _fields_ = [
        ("dof", C.c_int),
        ("n_bodies", C.c_int),
        ("n_links", C.c_int),
        ("n_spheres", C.c_int),
        ("n_pairs", C.c_int),
        ("resolution", C.c_int),
        ("body_parent", C.c_void_p),
        ("body_pre", C.c_void_p),
        ("body_type", C.c_void_p),
        ("body_axis", C.c_void_p),
        ("body_dof", C.c_void_p),
        ("link_body", C.c_void_p),
        ("link_first", C.c_void_p),
        ("link_count", C.c_void_p),
        ("link_bound", C.c_void_p),
        ("spheres", C.c_void_p),
        ("pairs", C.c_void_p),
        ("ee_body", C.c_int),
        ("ee_tf", C.c_void_p),
        ("n_attach_links", C.c_int),
        ("attach_links", C.c_void_p),
    ]


class Counters(C.Structure):
    _fields_ = [
        (n, C.c_uint64)
        for n in (
            "fk",
            "extent",
            "visited",
            "t_sphere",
            "t_capsule",
            "t_zcapsule",
            "t_cuboid",
            "t_zcuboid",
            "t_heightfield",
            "t_self",
            "capt_queries",
            "capt_points",
            "mvt_queries",
            "mvt_points",
        )
    ]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


_oracle_lib = None
_ref_lib = None


def oracle_lib():
    global _oracle_lib
    if _oracle_lib is None:
        if not ORACLE_SO.exists():
            build()
        L = C.CDLL(str(ORACLE_SO))
        L.or_env_create.restype = C.c_void_p
        L.or_env_dump.restype = C.c_size_t
        L.or_edge_steps.restype = C.c_size_t
        L.or_min_clearance.restype = C.c_double
        L.or_validate_edge.restype = C.c_int
        L.or_capt_query.restype = C.c_int
        L.or_capt_nlog2.restype = C.c_int
        L.or_capt_tests.restype = C.POINTER(C.c_float)
        L.or_capt_leaf_list.restype = C.c_size_t
        _oracle_lib = L
    return _oracle_lib


def ref_available() -> bool:
    return REF_SO.exists()


def ref_lib():
    global _ref_lib
    if _ref_lib is None:
        L = C.CDLL(str(REF_SO))
        L.ref_env_create.restype = C.c_void_p
        L.ref_env_dump.restype = C.c_size_t
        if hasattr(L, "ref_capt_nlog2"):
            L.ref_capt_nlog2.restype = C.c_int
            L.ref_capt_tests.restype = C.POINTER(C.c_float)
            L.ref_capt_leaf_list.restype = C.c_size_t
        L.ref_time_configs.restype = C.c_double
        L.ref_time_edges.restype = C.c_double
        if hasattr(L, "ref_simplify"):
            L.ref_simplify.restype = C.c_size_t
        _ref_lib = L
    return _ref_lib


# ------------------------------------------------------------------------------------------
class _EnvBase:
    """Same adders for both checkers; arguments are the fields the reference's shapes store
    (collision/shapes.hh)."""

    prefix = ""

    def __init__(self, lib):
        self.lib = lib
        self.h = C.c_void_p(getattr(lib, self.prefix + "env_create")())
        self._keep = []

    def __del__(self):
        try:
            getattr(self.lib, self.prefix + "env_destroy")(self.h)
        except Exception:
            pass

    def add_sphere(self, xyzr):
        getattr(self.lib, self.prefix + "env_add_sphere")(self.h, _fp(_f32(xyzr)))

    def add_cuboid(self, f15):
        a = _f32(f15)
        assert a.size == 15
        getattr(self.lib, self.prefix + "env_add_cuboid")(self.h, _fp(a))

    def add_capsule(self, f8):
        a = _f32(f8)
        assert a.size == 8
        getattr(self.lib, self.prefix + "env_add_capsule")(self.h, _fp(a))

    def add_heightfield(self, f6, xd, yd, data):
        a, d = _f32(f6), _f32(data)
        assert d.size == xd * yd
        getattr(self.lib, self.prefix + "env_add_heightfield")(
            self.h, _fp(a), C.c_size_t(xd), C.c_size_t(yd), _fp(d)
        )

    def add_capt(self, points, r_min, r_max, r_point):
        p = _f32(points).reshape(-1, 3)
        getattr(self.lib, self.prefix + "env_add_capt")(
            self.h, _fp(p), C.c_size_t(len(p)), C.c_float(r_min), C.c_float(r_max), C.c_float(r_point)
        )

    def add_mvt(self, points, r_min, r_max, aabb_min, aabb_max, r_point):
        p = _f32(points).reshape(-1, 3)
        getattr(self.lib, self.prefix + "env_add_mvt")(
            self.h,
            _fp(p),
            C.c_size_t(len(p)),
            C.c_float(r_min),
            C.c_float(r_max),
            _fp(_f32(aabb_min)),
            _fp(_f32(aabb_max)),
            C.c_float(r_point),
        )

    def add_raw_cloud(self, points, r_point):
        """Oracle only: the raw points of a cloud for the clearance accounting (no tree build)."""
        p = _f32(points).reshape(-1, 3)
        getattr(self.lib, self.prefix + "env_add_raw_cloud")(self.h, _fp(p), C.c_size_t(len(p)), C.c_float(r_point))

    def capt_tree(self, which: int = 0):
        """(nlog2, split values in Eytzinger order) of the which-th CAPT."""
        n = getattr(self.lib, self.prefix + "capt_nlog2")(self.h, C.c_size_t(which))
        t = getattr(self.lib, self.prefix + "capt_tests")(self.h, C.c_size_t(which))
        return n, np.ctypeslib.as_array(t, shape=((1 << n) - 1,)).copy() if n > 0 else np.zeros(0, np.float32)

    def capt_leaf_list(self, leaf: int, which: int = 0) -> np.ndarray:
        """The affordance list of one leaf, representative first, as [k][3]."""
        k = getattr(self.lib, self.prefix + "capt_leaf_list")(self.h, C.c_size_t(which), C.c_size_t(leaf), None, C.c_size_t(0))
        out = np.zeros((k, 3), np.float32)
        if k:
            getattr(self.lib, self.prefix + "capt_leaf_list")(self.h, C.c_size_t(which), C.c_size_t(leaf), _fp(out), C.c_size_t(k))
        return out


    def attach(self, tf12, spheres):
        s = _f32(spheres).reshape(-1, 4)
        getattr(self.lib, self.prefix + "env_attach")(self.h, _fp(_f32(tf12)), _fp(s), C.c_size_t(len(s)))

    def detach(self):
        getattr(self.lib, self.prefix + "env_detach")(self.h)

    def dump(self, kind):
        width = {0: 5, 1: 9, 2: 9, 3: 16, 4: 16}[kind]
        buf = np.zeros(4096 * width, np.float32)
        n = getattr(self.lib, self.prefix + "env_dump")(self.h, C.c_int(kind), _fp(buf), C.c_size_t(buf.size))
        return buf[: n * width].reshape(n, width).copy()


class OracleEnv(_EnvBase):
    prefix = "or_"

    def __init__(self):
        super().__init__(oracle_lib())

class RefEnv(_EnvBase):
    prefix = "ref_"

    def __init__(self):
        super().__init__(ref_lib())



def add_scene(env, scene: dict):
    """scene: {'spheres': [[x,y,z,r]...], 'cuboids': [[15]...], 'capsules': [[8]...]} in insertion order
    given by scene['order'] (list of (kind, index)) or spheres, cuboids, capsules when absent."""
    order = scene.get("order")
    if order is None:
        order = (
            [("spheres", i) for i in range(len(scene.get("spheres", [])))]
            + [("cuboids", i) for i in range(len(scene.get("cuboids", [])))]
            + [("capsules", i) for i in range(len(scene.get("capsules", [])))]
        )
    for kind, i in order:
        {"spheres": env.add_sphere, "cuboids": env.add_cuboid, "capsules": env.add_capsule}[kind](
            scene[kind][i]
        )
    return env


# ------------------------------------------------------------------------------------------
class Oracle:
    def __init__(self, robot: str):
        self.robot = robot
        self.lib = oracle_lib()
        m = load_model(robot)
        self.model = m
        self.dof = m["dof"]
        self.n_spheres = m["n_spheres"]
        B, L = m["bodies"], m["links"]
        i32 = lambda v: np.ascontiguousarray(v, dtype=np.int32)
        ee = m["end_effector"]
        self._arrays = dict(
            body_parent=i32([b["parent"] for b in B]),
            body_pre=_f32([np.array(b["T_pre"]).reshape(-1) for b in B]),
            body_type=i32([{"fixed": 0, "revolute": 1, "prismatic": 2}[b["jtype"]] for b in B]),
            body_axis=_f32([np.array(b["axis"]) / max(np.linalg.norm(b["axis"]), 1e-30) for b in B]),
            body_dof=i32([b["dof"] for b in B]),
            link_body=i32([l["body"] for l in L]),
            link_first=i32([l["first_sphere"] for l in L]),
            link_count=i32([len(l["spheres"]) for l in L]),
            link_bound=_f32([l["bound"] for l in L]),
            spheres=_f32([s for l in L for s in l["spheres"]]),
            pairs=i32(m["self_pairs"]).reshape(-1, 2),
            ee_tf=_f32(np.array(ee["T"]).reshape(-1)),
            attach_links=i32(m["attach_links"]),
        )
        a = self._arrays
        self.c = _OrModel(
            m["dof"],
            len(B),
            len(L),
            m["n_spheres"],
            len(m["self_pairs"]),
            m["resolution"],
            *[a[k].ctypes.data for k in (
                "body_parent", "body_pre", "body_type", "body_axis", "body_dof",
                "link_body", "link_first", "link_count", "link_bound", "spheres", "pairs",
            )],
            ee["body"],
            a["ee_tf"].ctypes.data,
            len(m["attach_links"]),
            a["attach_links"].ctypes.data,
        )

    def set_sqrt_mode(self, mode: int):
        self.lib.or_set_sqrt_mode(C.c_int(mode))

    def validate_configs(self, env: OracleEnv, q, counters: Counters | None = None):
        q = _f32(q).reshape(-1, self.dof)
        out = np.zeros(len(q), np.uint8)
        self.lib.or_validate_configs(
            C.byref(self.c), env.h, _fp(q), C.c_size_t(len(q)), _fp(out),
            C.byref(counters) if counters is not None else None,
        )
        return out.astype(bool)

    def validate_edges(self, env: OracleEnv, a, b, counters: Counters | None = None):
        a = _f32(a).reshape(-1, self.dof)
        b = _f32(b).reshape(-1, self.dof)
        out = np.zeros(len(a), np.uint8)
        states = C.c_uint64(0)
        self.lib.or_validate_edges(
            C.byref(self.c), env.h, _fp(a), _fp(b), C.c_size_t(len(a)), _fp(out),
            C.byref(counters) if counters is not None else None, C.byref(states),
        )
        self.last_states = int(states.value)
        return out.astype(bool)

    def filter_points(self, env: OracleEnv, q, points, point_radius: float) -> np.ndarray:
        p = _f32(points).reshape(-1, 3)
        keep = np.zeros(len(p), np.uint8)
        self.lib.or_filter_points(C.byref(self.c), env.h, _fp(_f32(q)), _fp(p), C.c_size_t(len(p)), C.c_float(point_radius), _fp(keep))
        return keep.astype(bool)

    def edge_steps(self, a, b):
        return int(self.lib.or_edge_steps(C.byref(self.c), _fp(_f32(a)), _fp(_f32(b))))

    def sphere_fk(self, q):
        q = _f32(q).reshape(-1, self.dof)
        out = np.zeros((len(q), self.n_spheres, 4), np.float32)
        self.lib.or_sphere_fk(C.byref(self.c), _fp(q), C.c_size_t(len(q)), _fp(out))
        return out

    def eefk(self, q):
        out = np.zeros(16, np.float32)
        self.lib.or_eefk(C.byref(self.c), _fp(_f32(q)), _fp(out))
        return out.reshape(4, 4)

    def min_clearance(self, env: OracleEnv, q):
        q = _f32(q).reshape(-1, self.dof)
        return np.array(
            [self.lib.or_min_clearance(C.byref(self.c), env.h, _fp(q[i])) for i in range(len(q))]
        )

    def clearance(self, env: OracleEnv, q):
        """[n][2]: minimum signed clearance of each state (north_star's quantity) and the distance of its
        nearest decision boundary (adds link bounding spheres vs cloud points; see or_clearance)."""
        q = _f32(q).reshape(-1, self.dof)
        out = np.zeros((len(q), 2), np.float64)
        for i in range(len(q)):
            self.lib.or_clearance(C.byref(self.c), env.h, _fp(q[i]), out[i].ctypes.data_as(C.c_void_p))
        return out

    def point_clearance(self, env: OracleEnv, q, points, point_radius: float):
        """[n][2] for filter_self_from_pointcloud: each point's sphere vs the robot at q and the environment."""
        p = _f32(points).reshape(-1, 3)
        qq = _f32(q)
        out = np.zeros((len(p), 2), np.float64)
        for i in range(len(p)):
            self.lib.or_point_clearance(C.byref(self.c), env.h, _fp(qq), _fp(p[i]), C.c_float(point_radius), out[i].ctypes.data_as(C.c_void_p))
        return out

    def edge_clearance(self, env: OracleEnv, a, b):
        """[n][2] over the 8n rake states of each edge: signed clearance of the state closest to zero, and
        the smallest nearest-decision-boundary distance."""
        a = _f32(a).reshape(-1, self.dof)
        b = _f32(b).reshape(-1, self.dof)
        out = np.zeros((len(a), 2), np.float64)
        for i in range(len(a)):
            self.lib.or_edge_clearance(C.byref(self.c), env.h, _fp(a[i]), _fp(b[i]), out[i].ctypes.data_as(C.c_void_p))
        return out

    def debug(self, env: OracleEnv, q):
        eh = np.zeros((65536, 2), np.int32)
        sh = np.zeros((65536, 2), np.int32)
        ne, ns = C.c_size_t(0), C.c_size_t(0)
        self.lib.or_debug(
            C.byref(self.c), env.h, _fp(_f32(q)), _fp(eh), C.c_size_t(len(eh)), C.byref(ne),
            _fp(sh), C.c_size_t(len(sh)), C.byref(ns),
        )
        return eh[: ne.value].copy(), sh[: ns.value].copy()


class Ref:
    def __init__(self, robot: str):
        self.robot = robot
        self.id = ROBOT_IDS[robot]
        self.lib = ref_lib()
        self.dof = self.lib.ref_robot_dim(self.id)
        self.n_spheres = self.lib.ref_robot_n_spheres(self.id)
        self.resolution = self.lib.ref_robot_resolution(self.id)

    def validate_configs(self, env: RefEnv, q, threads: int = 1):
        q = _f32(q).reshape(-1, self.dof)
        out = np.zeros(len(q), np.uint8)
        self.lib.ref_validate_configs(self.id, env.h, _fp(q), C.c_size_t(len(q)), _fp(out), C.c_int(threads))
        return out.astype(bool)

    def validate_edges(self, env: RefEnv, a, b, threads: int = 1):
        a = _f32(a).reshape(-1, self.dof)
        b = _f32(b).reshape(-1, self.dof)
        out = np.zeros(len(a), np.uint8)
        self.lib.ref_validate_edges(
            self.id, env.h, _fp(a), _fp(b), C.c_size_t(len(a)), _fp(out), C.c_int(threads)
        )
        return out.astype(bool)

    def filter_points(self, env: RefEnv, q, points, point_radius: float) -> np.ndarray:
        p = _f32(points).reshape(-1, 3)
        keep = np.zeros(len(p), np.uint8)
        self.lib.ref_filter_points(self.id, env.h, _fp(_f32(q)), _fp(p), C.c_size_t(len(p)), C.c_float(point_radius), _fp(keep))
        return keep.astype(bool)

    def simplify(self, env: RefEnv, path, operations=(2, 0), settings12=None, samples=None):
        """The reference's own simplify<Robot, 8, resolution> (planning/simplify.hh:191-258).
        operations: SimplifyRoutine values; settings12 as in oracle/ref/ref_robot.hh; samples: the
        unit-cube stream behind RNG::next().  Returns (path [k][dof], iterations)."""
        p = _f32(path).reshape(-1, self.dof)
        ops = np.ascontiguousarray(np.asarray(operations, np.int32))
        sf = _f32(settings12 if settings12 is not None else [5, 0, 1, 0.1, 0.5, 10, 5, 0.5, 10, 5, 5, 0.1])
        sm = _f32(samples if samples is not None else np.zeros((1, self.dof))).reshape(-1, self.dof)
        cap = max(4096, 64 * len(p))
        out = np.zeros((cap, self.dof), np.float32)
        it = C.c_size_t(0)
        n = self.lib.ref_simplify(
            self.id, env.h, _fp(p), C.c_size_t(len(p)), ops.ctypes.data_as(C.c_void_p), C.c_size_t(len(ops)),
            _fp(sf), _fp(sm), C.c_size_t(len(sm)), _fp(out), C.c_size_t(cap), C.byref(it),
        )
        assert n <= cap
        return out[:n].copy(), int(it.value)

    def path_op(self, op: int, path, arg: int = 0):
        """Path<Robot> members of the reference (planning/plan.hh): op 0 cost, 1 subdivide,
        2 interpolate_to_resolution(arg), 3 interpolate_to_n_states(arg) -> (waypoints, cost)."""
        self.lib.ref_path_op.restype = C.c_size_t
        p = _f32(path).reshape(-1, self.dof)
        cap = 1 << 16
        out = np.zeros((cap, self.dof), np.float32)
        cost = C.c_float(0)
        n = self.lib.ref_path_op(self.id, C.c_int(op), _fp(p), C.c_size_t(len(p)), C.c_size_t(arg), _fp(out), C.c_size_t(cap), C.byref(cost))
        assert n <= cap
        return out[:n].copy(), float(cost.value)

    def planner(self, which: int, env: RefEnv, start, goal, max_iterations=100000, max_samples=100000, batch_size=1000):
        """The reference's own planners compiled in place over an exact brute-force NN stand-in and its Halton sampler:
        which = 0 PRM::build_roadmap -> (vertices, edges [(from, to)] per adjacency entry, iterations);
        1 PRM::solve, 2 FCIT::solve -> (path waypoints as the reference returns them, cost, iterations)."""
        self.lib.ref_planner.restype = C.c_size_t
        cap_v = max_samples + 8
        verts = np.zeros((cap_v, self.dof), np.float32)
        cap_e = 1 << 22
        edges = np.zeros((cap_e, 2), np.uint32)
        ne, it, cost = C.c_size_t(0), C.c_size_t(0), C.c_float(0)
        nv = self.lib.ref_planner(self.id, C.c_int(which), env.h, _fp(_f32(start)), _fp(_f32(goal)), C.c_size_t(max_iterations), C.c_size_t(max_samples),
                                  C.c_size_t(batch_size), _fp(verts), C.c_size_t(cap_v), edges.ctypes.data_as(C.c_void_p), C.c_size_t(cap_e),
                                  C.byref(ne), C.byref(it), C.byref(cost))
        assert nv <= cap_v and ne.value <= cap_e
        if which == 0:
            return verts[:nv].copy(), edges[: ne.value].copy(), int(it.value)
        return verts[:nv].copy(), float(cost.value), int(it.value)

    def halton(self, n: int, skip: int = 0) -> np.ndarray:
        out = np.zeros((n, self.dof), np.float32)
        self.lib.ref_halton(self.id, C.c_size_t(skip), C.c_size_t(n), _fp(out))
        return out

    def sphere_fk(self, q):
        q = _f32(q).reshape(-1, self.dof)
        out = np.zeros((len(q), self.n_spheres, 4), np.float32)
        self.lib.ref_sphere_fk(self.id, _fp(q), C.c_size_t(len(q)), _fp(out))
        return out

    def eefk(self, q):
        out = np.zeros(16, np.float32)
        self.lib.ref_eefk(self.id, _fp(_f32(q)), _fp(out))
        return out.reshape(4, 4)

    def debug(self, env: RefEnv, q):
        eh = np.zeros((65536, 2), np.int32)
        sh = np.zeros((65536, 2), np.int32)
        ne, ns = C.c_size_t(0), C.c_size_t(0)
        self.lib.ref_debug(
            self.id, env.h, _fp(_f32(q)), _fp(eh), C.c_size_t(len(eh)), C.byref(ne),
            _fp(sh), C.c_size_t(len(sh)), C.byref(ns),
        )
        return eh[: ne.value].copy(), sh[: ns.value].copy()

    def time_configs(self, env: RefEnv, q, threads: int, reps: int = 3) -> float:
        q = _f32(q).reshape(-1, self.dof)
        return float(
            self.lib.ref_time_configs(self.id, env.h, _fp(q), C.c_size_t(len(q)), C.c_int(threads), C.c_int(reps))
        )

    def time_edges(self, env: RefEnv, a, b, threads: int, reps: int = 3) -> float:
        a = _f32(a).reshape(-1, self.dof)
        b = _f32(b).reshape(-1, self.dof)
        return float(
            self.lib.ref_time_edges(
                self.id, env.h, _fp(a), _fp(b), C.c_size_t(len(a)), C.c_int(threads), C.c_int(reps)
            )
        )


def filter_centervox(points, voxel_size, max_range, origin, ws_min, ws_max):
    """CPU restatement (oracle/vamp_oracle.c: or_filter_centervox) of the reference's CenterVox filter
    (collision/filter_centervox.hh): indices of the kept points in the reference's output order, or None
    where the reference throws (voxel pool exhausted)."""
    L = oracle_lib()
    L.or_filter_centervox.restype = C.c_long
    p = _f32(points).reshape(-1, 3)
    out = np.zeros(32768, np.uint32)
    k = L.or_filter_centervox(_fp(p), C.c_size_t(len(p)), C.c_float(voxel_size), C.c_float(max_range), _fp(_f32(origin)),
                              _fp(_f32(ws_min)), _fp(_f32(ws_max)), out.ctypes.data_as(C.c_void_p), C.c_size_t(len(out)))
    return None if k < 0 else out[:k].copy()


def ref_filter_centervox(points, voxel_size, max_range, origin, ws_min, ws_max):
    """The reference's own filter_pointcloud_centervox compiled in place (oracle/_ref): kept points [k][3],
    or None where it throws."""
    L = ref_lib()
    L.ref_filter_centervox.restype = C.c_size_t
    p = _f32(points).reshape(-1, 3)
    out = np.zeros((max(len(p), 1), 3), np.float32)
    k = L.ref_filter_centervox(_fp(p), C.c_size_t(len(p)), C.c_float(voxel_size), C.c_float(max_range), _fp(_f32(origin)),
                               _fp(_f32(ws_min)), _fp(_f32(ws_max)), _fp(out), C.c_size_t(len(out)))
    return None if k == C.c_size_t(-1).value else out[:k].copy()


def host_threads() -> int:
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
