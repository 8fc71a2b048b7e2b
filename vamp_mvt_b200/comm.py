"""Multi-GPU communicator (include/vamp_b200.h: vmv_comm_*): one process per GPU of one node.

The library does the data plane itself -- NCCL (loaded by the library at run time) for the environment
broadcast and the plain all-gather, CUDA-IPC windows for the gather fused into the validation kernels.  The 128-byte
id of ``unique_id()`` travels from rank 0 to the other ranks by whatever control plane the caller has (an MPI
broadcast, a file, a TCP store, torch.distributed's object broadcast in bench.py); nothing here imports torch.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def unique_id() -> bytes:
    buf = (C.c_ubyte * 128)()
    _lib.check(_lib.lib().vmv_comm_unique_id(buf))
    return bytes(buf)


class Communicator:
    def __init__(self, uid: bytes, rank: int, world: int):
        """Collective: every rank calls it with rank 0's id, after selecting its device (vmv_set_device)."""
        assert len(uid) == 128
        self._L = _lib.lib()
        h = C.c_void_p()
        buf = (C.c_ubyte * 128).from_buffer_copy(uid)
        _lib.check(self._L.vmv_comm_create(C.byref(h), buf, rank, world))
        self.handle, self.rank, self.world = h, rank, world
        self.stride = 0

    def close(self):
        if self.handle:
            self._L.vmv_comm_destroy(self.handle)
            self.handle = None

    def broadcast_environment(self, env, root: int = 0):
        """Replace `env` (a vamp_mvt_b200.Environment) on every rank but `root` with root's; committed on return."""
        _lib.check(self._L.vmv_env_broadcast(self.handle, env._h, root))
        env._dirty = False

    def allgather_words(self, d_local: int, words_per_rank: int, d_global: int, stream=None):
        """Plain ncclAllGather of verdict words (device pointers)."""
        _lib.check(self._L.vmv_allgather_bits(self.handle, d_local, words_per_rank, d_global, stream))

    def window(self, words_per_rank: int, slots: int = 2):
        """Collective, once: the IPC windows of the kernel-fused gather."""
        _lib.check(self._L.vmv_comm_window(self.handle, words_per_rank, slots))
        self.stride = int(self._L.vmv_comm_window_stride(self.handle))

    def window_ptr(self, slot: int) -> int:
        return int(self._L.vmv_comm_window_ptr(self.handle, slot))

    def validate_configs_gather(self, robot_id: int, env_handle, slot: int, d_q: int, n: int, stream=None):
        _lib.check(self._L.vmv_validate_configs_gather_dev(robot_id, env_handle, self.handle, slot, d_q, n, stream))

    def validate_edges_indexed_gather(self, robot_id: int, env_handle, slot: int, d_vertices: int, n_vertices: int, d_pairs: int, n_edges: int,
                                      resolution: int = 0, stream=None):
        _lib.check(self._L.vmv_validate_edges_indexed_gather_dev(robot_id, env_handle, self.handle, slot, d_vertices, n_vertices, d_pairs, n_edges,
                                                                 resolution, stream))

    def local_row(self, slot: int) -> int:
        """Device pointer of this rank's row in its own window (the target of a launch followed by publish)."""
        return int(self._L.vmv_comm_local_row(self.handle, slot))

    def publish(self, slot: int, n_words: int, stream=None):
        """Copy-engine publication of the local row of `slot` to every peer, ordered behind `stream`."""
        _lib.check(self._L.vmv_comm_publish(self.handle, slot, n_words, stream))

    def acquire(self, slot: int, stream=None):
        """Order `stream` behind the copy engines' send of the slot's previous row (call before overwriting it)."""
        _lib.check(self._L.vmv_comm_acquire(self.handle, slot, stream))

    def wait(self, slot: int, stream=None):
        _lib.check(self._L.vmv_comm_wait(self.handle, slot, stream))

    def read_window(self, slot: int, n_units_per_rank) -> np.ndarray:
        """Host copy of the global mask in `slot` (after wait + stream sync): bool[sum(n_units_per_rank)], the
        ranks' shards concatenated."""
        words = np.zeros(self.world * self.stride, np.uint32)
        _lib.check(self._L.vmv_memcpy_d2h(_lib.ptr(words), self.window_ptr(slot), words.nbytes, None))
        _lib.check(self._L.vmv_stream_sync(None))
        parts = []
        for r, n in enumerate(n_units_per_rank):
            parts.append(_lib.unpack_bits(words[r * self.stride : (r + 1) * self.stride], n))
        return np.concatenate(parts) if parts else np.zeros(0, bool)
