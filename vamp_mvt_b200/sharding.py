"""Multi-GPU plumbing: contiguous, word-aligned shards and the verdict-bitmask all-gather.

Every configuration / edge is an independent unit with a one-bit result, so ranks own disjoint
contiguous ranges [lo, hi) aligned to 32 units (whole verdict words) and no data moves during the
computation.  Only when the caller needs the global result is there one collective: an all-gather of
ceil(n/32/world) words per rank (NCCL over NVLink on GPUs; the same code runs on gloo for the CPU
tests).  torch.distributed is plumbing here -- the validation itself never touches torch.
"""
from __future__ import annotations

from typing import Tuple


def words_per_rank(n_units: int, world: int) -> int:
    words = (n_units + 31) // 32
    return (words + world - 1) // world


def shard_bounds(n_units: int, rank: int, world: int) -> Tuple[int, int]:
    """Unit range [lo, hi) owned by `rank`; lo is a multiple of 32."""
    per = words_per_rank(n_units, world) * 32
    lo = min(rank * per, n_units)
    hi = min(lo + per, n_units)
    return lo, hi


def allgather_verdict_words(local_words, n_units: int, group=None):
    """local_words: int32 tensor holding this rank's verdict words (ceil((hi-lo)/32) of them, on the
    device the process group works with).  Returns the global word tensor (ceil(n_units/32) words)
    on every rank."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    per = words_per_rank(n_units, world)
    padded = torch.zeros(per, dtype=torch.int32, device=local_words.device)
    padded[: local_words.numel()] = local_words
    out = torch.empty(per * world, dtype=torch.int32, device=local_words.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    return out[: (n_units + 31) // 32]
