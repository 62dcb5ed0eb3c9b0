"""Replica sharding over the GPUs of one box: the reference's rayon fan-out over simulation ids
(src/main.rs:89-121) becomes "rank r runs replicas r*per .. (r+1)*per-1"; there is no data-path communication."""
from __future__ import annotations


def replicas_of_rank(n_replicas: int, rank: int, world: int):
    """Contiguous, balanced split; earlier ranks take the remainder."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    base, extra = divmod(n_replicas, world)
    lo = rank * base + min(rank, extra)
    return list(range(lo, lo + base + (1 if rank < extra else 0)))
