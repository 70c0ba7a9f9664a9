"""Image-level sharding across GPUs: frames are independent units (no cross-image state in m06..m08), so a
batch is split by image index, one process per GPU, with no collective on the data path."""
from __future__ import annotations


def shard_contiguous(n: int, rank: int, world: int) -> range:
    """Contiguous block of [0, n) for `rank`; sizes differ by at most one."""
    base, extra = divmod(n, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def shard_by_cost(costs, world: int):
    """Longest-first greedy (LPT) assignment for mixed-size batches: cost = macroblocks per frame.
    Returns `world` lists of indices; deterministic, every index appears exactly once."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    loads = [0] * world
    out = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (loads[k], k))
        out[r].append(i)
        loads[r] += costs[i]
    for lst in out:
        lst.sort()
    return out
