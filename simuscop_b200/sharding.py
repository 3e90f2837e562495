"""Pair-ID sharding across GPUs: contiguous ranges of planned pair IDs, no collective.

Every pair depends only on (seed, pair ID, tables, haplotype slice), and the two cross-read couplings
(failCount abort per bin, fragCount per segment) are resolved by the census prefix sums of ssc_set_plan,
so any split of [0, planned) concatenates to the single-GPU output byte for byte.
"""


def shard_range(planned, rank, world):
    """Range [lo, hi) of planned pair IDs owned by `rank` out of `world` (balanced, contiguous, ordered)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    base, extra = divmod(int(planned), world)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi
