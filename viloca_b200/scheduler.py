"""Host-side longest-first sharding of windows over GPUs.

Windows are independent problems; there is no exchange step and therefore no collective.
The restarts of one window stay on one GPU so that they share the window's read tensors.
"""
from __future__ import annotations

from typing import Sequence


def lpt_partition(costs: Sequence[float], n_bins: int) -> list:
    """Longest-processing-time-first: items sorted by cost descending (stable), each placed
    on the currently least-loaded bin (lowest index on ties).  Returns one index list per bin;
    deterministic, so every rank can compute the same partition without communicating."""
    if n_bins <= 0:
        raise ValueError("n_bins must be positive")
    order = sorted(range(len(costs)), key=lambda i: (-float(costs[i]), i))
    loads = [0.0] * n_bins
    bins = [[] for _ in range(n_bins)]
    for i in order:
        j = min(range(n_bins), key=lambda b: (loads[b], b))
        bins[j].append(i)
        loads[j] += float(costs[i])
    return bins


def shard_for_rank(costs: Sequence[float], rank: int, world_size: int) -> list:
    """Indices of the windows rank ``rank`` of ``world_size`` processes runs."""
    return lpt_partition(costs, world_size)[rank]
