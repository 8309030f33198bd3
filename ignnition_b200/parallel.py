"""Multi-GPU plumbing: one process per GPU, samples sharded across ranks.

The reference is single-process; batches of independent samples shard naturally
(``model_fn`` loops over samples, ``code/utils/generate_model.py:712-724``).  Inference needs no
communication.  Training all-reduces the flat gradient buffer (NCCL over NVLink) and every rank
scales its squared-error gradient by 1 / GLOBAL prediction count, which reproduces
``MeanSquaredError`` over all predictions of the global batch (:745-751).
"""

from __future__ import annotations

import os
from typing import List, Sequence, Tuple


def rank_world() -> Tuple[int, int, int]:
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def shard_bounds(costs: Sequence[float], world: int) -> List[Tuple[int, int]]:
    """Contiguous slices of the sample list, balanced by cost (e.g. edges per sample), one per rank."""
    n = len(costs)
    total = float(sum(costs))
    bounds, start, acc = [], 0, 0.0
    for r in range(world):
        target = total * (r + 1) / world
        end = start
        while end < n and (acc + costs[end] <= target + 1e-9 or end == start) and (n - end) > (world - 1 - r):
            acc += costs[end]
            end += 1
        if r == world - 1:
            end = n
        bounds.append((start, end))
        start = end
    return bounds


def shard_samples(samples: Sequence, rank: int, world: int, costs: Sequence[float] = None) -> List:
    costs = costs if costs is not None else [1.0] * len(samples)
    lo, hi = shard_bounds(costs, world)[rank]
    return list(samples[lo:hi])
