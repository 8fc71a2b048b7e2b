"""World-size-2 test of the N>1 path on CPU (gloo): word-aligned sharding + verdict-bitmask
all-gather.  The per-shard verdicts come from the CPU oracle here (the GPU kernels are exercised by
the -m gpu tests); what is under test is the host-side sharding/gather logic bench.py and
multi-GPU callers use."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parents[1]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, out_dir):
    sys.path.insert(0, str(REPO))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch
    import torch.distributed as dist

    from oracle import pyoracle as po
    from tests import scenes
    from vamp_mvt_b200 import sharding

    dist.init_process_group("gloo", rank=rank, world_size=world)
    oracle = po.Oracle("panda")
    env = po.add_scene(po.OracleEnv(), scenes.packed(scenes.table_shelf_scene()))
    q = scenes.random_configs("panda", n, seed=3)  # every rank generates the same global batch
    lo, hi = sharding.shard_bounds(n, rank, world)
    valid = oracle.validate_configs(env, q[lo:hi]) if hi > lo else np.zeros(0, bool)
    words = np.packbits(np.pad(valid, (0, (-len(valid)) % 32)), bitorder="little").view(np.uint32).astype(np.int64)
    local = torch.from_numpy(words.astype(np.int64)).to(torch.int32) if len(words) else torch.zeros(0, dtype=torch.int32)
    full = sharding.allgather_verdict_words(local, n)
    np.save(Path(out_dir) / f"rank{rank}.npy", full.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [1000, 4096 + 17])
def test_two_rank_shard_and_gather(tmp_path, n):
    import torch.multiprocessing as mp

    from oracle import pyoracle as po
    from tests import scenes

    port = _free_port()
    mp.spawn(_worker, args=(2, port, n, str(tmp_path)), nprocs=2, join=True)
    oracle = po.Oracle("panda")
    env = po.add_scene(po.OracleEnv(), scenes.packed(scenes.table_shelf_scene()))
    want = oracle.validate_configs(env, scenes.random_configs("panda", n, seed=3))
    for rank in range(2):
        words = np.load(tmp_path / f"rank{rank}.npy").astype(np.uint32)
        got = np.unpackbits(words.view(np.uint8), bitorder="little")[:n].astype(bool)
        assert np.array_equal(got, want)
