"""Site sharding of an alignment over ranks: rank g of N owns the contiguous
site range [lo, hi) of the alignment and all of its CLVs, scalers and tip data; boundaries are multiples
of 32 sites (tile-aligned bulk copies), the last rank takes the remainder.  The only exchange of an
evaluation is one all-reduce (sum) of the log-likelihood, or of the {d_f, dd_f} pair."""
from __future__ import annotations

ALIGN = 32


def shard_bounds(sites: int, world: int, rank: int) -> tuple[int, int]:
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"rank {rank} of {world}")
    per = -(-sites // world)              # ceil
    per = -(-per // ALIGN) * ALIGN        # round the share up to a multiple of ALIGN
    lo = min(rank * per, sites)
    hi = min(lo + per, sites)
    return lo, hi


def all_bounds(sites: int, world: int) -> list[tuple[int, int]]:
    return [shard_bounds(sites, world, r) for r in range(world)]
