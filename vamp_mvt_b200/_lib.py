"""ctypes binding of the C ABI (include/vamp_b200.h).  The shared library is built in-tree by
``__graft_entry__.build()`` (nvcc, sm_100a).  There is no fallback: a missing library raises."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
import os

LIB_PATH = Path(os.environ.get("VMV_LIB", str(HERE / "libvamp_b200.so")))

OK = 0


class VmvError(RuntimeError):
    pass


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise VmvError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc -gencode arch=compute_100a,code=sm_100a); there is no CPU fallback"
            )
        L = C.CDLL(str(LIB_PATH))
        vp, sz, i32, f32 = C.c_void_p, C.c_size_t, C.c_int, C.c_float
        L.vmv_last_error.restype = C.c_char_p
        L.vmv_version.restype = C.c_char_p
        L.vmv_robot_name.restype = C.c_char_p
        L.vmv_robot_name.argtypes = [i32]
        L.vmv_robot_id.argtypes = [C.c_char_p]
        L.vmv_env_create.restype = vp
        L.vmv_env_destroy.argtypes = [vp]
        L.vmv_env_add_spheres.argtypes = [vp, vp, sz]
        L.vmv_env_add_cuboids.argtypes = [vp, vp, sz]
        L.vmv_env_add_capsules.argtypes = [vp, vp, sz]
        L.vmv_env_add_heightfield.argtypes = [vp, vp, sz, sz, vp]
        L.vmv_env_add_capt.argtypes = [vp, vp, sz, f32, f32, f32]
        L.vmv_env_add_mvt.argtypes = [vp, vp, sz, f32, f32, vp, vp, f32]
        L.vmv_env_attach.argtypes = [vp, vp, vp, sz]
        L.vmv_env_detach.argtypes = [vp]
        L.vmv_env_commit.argtypes = [vp]
        L.vmv_env_dump.restype = C.c_long
        L.vmv_env_dump.argtypes = [vp, i32, vp, sz]
        L.vmv_robot_bounds.argtypes = [i32, vp, vp]
        L.vmv_validate_configs_dev.argtypes = [i32, vp, vp, sz, vp, vp]
        L.vmv_validate_edges_dev.argtypes = [i32, vp, vp, vp, sz, i32, vp, vp]
        L.vmv_validate_edges_indexed_dev.argtypes = [i32, vp, vp, sz, vp, sz, i32, vp, vp]
        L.vmv_validate_edges_indexed.argtypes = [i32, vp, vp, sz, vp, sz, i32, vp]
        L.vmv_validate_configs.argtypes = [i32, vp, vp, sz, vp]
        L.vmv_validate_edges.argtypes = [i32, vp, vp, vp, sz, i32, vp]
        L.vmv_sphere_fk_dev.argtypes = [i32, vp, sz, vp, vp]
        L.vmv_sphere_fk.argtypes = [i32, vp, sz, vp]
        L.vmv_filter_points_dev.argtypes = [i32, vp, vp, vp, sz, f32, vp, vp]
        L.vmv_filter_self_from_pointcloud.argtypes = [i32, vp, vp, vp, sz, f32, vp]
        L.vmv_debug.argtypes = [i32, vp, vp, vp, sz, vp, vp, sz, vp]
        L.vmv_filter_pointcloud_centervox.argtypes = [vp, sz, f32, f32, vp, vp, vp, vp, sz, vp]
        L.vmv_filter_pointcloud_centervox_dev.argtypes = [vp, sz, f32, f32, vp, vp, vp, vp, sz, vp]
        L.vmv_halton_fill_dev.argtypes = [i32, C.c_uint64, sz, vp, vp]
        L.vmv_validate_halton.argtypes = [i32, vp, C.c_uint64, sz, vp, vp]
        L.vmv_halton_exact_limit.restype = C.c_uint64
        L.vmv_halton_exact_limit.argtypes = [i32]
        L.vmv_dev_alloc.restype = vp
        L.vmv_dev_alloc.argtypes = [sz]
        L.vmv_dev_free.argtypes = [vp]
        L.vmv_host_alloc_pinned.restype = vp
        L.vmv_host_alloc_pinned.argtypes = [sz]
        L.vmv_host_free_pinned.argtypes = [vp]
        L.vmv_memcpy_h2d.argtypes = [vp, vp, sz, vp]
        L.vmv_memcpy_d2h.argtypes = [vp, vp, sz, vp]
        L.vmv_stream_sync.argtypes = [vp]
        L.vmv_launch_count.restype = C.c_uint64
        L.vmv_force_kernel_path.argtypes = [i32]
        L.vmv_prm.argtypes = [i32, vp, vp, vp, sz, sz, C.c_double, i32, vp]
        L.vmv_fcit.argtypes = [i32, vp, vp, vp, sz, sz, sz, i32, vp]
        L.vmv_roadmap_destroy.argtypes = [vp]
        L.vmv_roadmap_destroy.restype = None
        for fn in ("vmv_roadmap_vertices", "vmv_roadmap_edges", "vmv_roadmap_path", "vmv_roadmap_iterations", "vmv_roadmap_work"):
            getattr(L, fn).restype = sz
        L.vmv_roadmap_vertices.argtypes = [vp, vp]
        L.vmv_roadmap_edges.argtypes = [vp, vp, vp]
        L.vmv_roadmap_path.argtypes = [vp, vp, vp]
        L.vmv_roadmap_iterations.argtypes = [vp]
        L.vmv_roadmap_work.argtypes = [vp, vp]
        L.vmv_comm_unique_id.argtypes = [vp]
        L.vmv_comm_create.argtypes = [vp, vp, i32, i32]
        L.vmv_comm_destroy.argtypes = [vp]
        L.vmv_comm_destroy.restype = None
        L.vmv_comm_rank.argtypes = [vp]
        L.vmv_comm_world.argtypes = [vp]
        L.vmv_env_broadcast.argtypes = [vp, vp, i32]
        L.vmv_allgather_bits.argtypes = [vp, vp, sz, vp, vp]
        L.vmv_comm_window.argtypes = [vp, sz, i32]
        L.vmv_comm_window_ptr.restype = vp
        L.vmv_comm_window_ptr.argtypes = [vp, i32]
        L.vmv_comm_window_stride.restype = sz
        L.vmv_comm_window_stride.argtypes = [vp]
        L.vmv_validate_configs_gather_dev.argtypes = [i32, vp, vp, i32, vp, sz, vp]
        L.vmv_validate_edges_indexed_gather_dev.argtypes = [i32, vp, vp, i32, vp, sz, vp, sz, i32, vp]
        L.vmv_comm_wait.argtypes = [vp, i32, vp]
        L.vmv_comm_local_row.restype = vp
        L.vmv_comm_local_row.argtypes = [vp, i32]
        L.vmv_comm_publish.argtypes = [vp, i32, sz, vp]
        L.vmv_comm_acquire.argtypes = [vp, i32, vp]
        L.vmv_env_capt_digest.argtypes = [vp, i32, vp]
        L.vmv_env_capt_nodes.argtypes = [vp, i32, vp, C.c_size_t]
        L.vmv_env_capt_nodes.restype = C.c_long
        _lib = L
    return _lib


def check(rc: int) -> int:
    if rc < 0:
        raise VmvError(f"vamp_b200 error {rc}: {lib().vmv_last_error().decode()}")
    return rc


def f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def unpack_bits(words: np.ndarray, n: int) -> np.ndarray:
    """valid_bits words -> bool[n] (bit i&31 of word i>>5)."""
    b = np.unpackbits(words.view(np.uint8), bitorder="little")
    return b[:n].astype(bool)
