"""Host-side channel partitioning across the GPUs of one box.

Channels are fully independent (no shared state and no cross-channel reduction anywhere in
AudioDriver_RxProcessor, mchf-eclipse/drivers/audio/audio_driver.c:2603-2942), so the engine
shards them across ranks on the host and the data path needs no collective: rank g owns the
contiguous global channel range [g*C/G, (g+1)*C/G).  For a mixed-mode plan the channels are first
grouped by their configuration (mode / filter path) and every group is split evenly, so each GPU
gets the same mode mix -- and inside one GPU, channels of one kind are contiguous, which is what
the kernels' per-kind channel lists want.
"""
from __future__ import annotations

from typing import Hashable, Sequence


def channel_range(rank: int, world: int, total: int) -> tuple[int, int]:
    """Contiguous [lo, hi) of `total` channels owned by `rank` of `world`; sizes differ by at most one."""
    if world < 1 or not 0 <= rank < world or total < 0:
        raise ValueError(f"bad partition request rank={rank} world={world} total={total}")
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def partition_by_kind(kinds: Sequence[Hashable], world: int) -> list[list[int]]:
    """Split global channels 0..len(kinds)-1 over `world` ranks so that every rank receives an equal
    share (+-1) of every kind.  Returns, per rank, the global channel indices it owns, grouped by kind
    in first-appearance order and ascending inside a kind."""
    if world < 1:
        raise ValueError("world must be >= 1")
    groups: dict[Hashable, list[int]] = {}
    for ch, k in enumerate(kinds):
        groups.setdefault(k, []).append(ch)
    owned: list[list[int]] = [[] for _ in range(world)]
    for members in groups.values():
        for r in range(world):
            lo, hi = channel_range(r, world, len(members))
            owned[r].extend(members[lo:hi])
    return owned
