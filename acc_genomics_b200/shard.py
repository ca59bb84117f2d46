"""Partitioning of a job's regions over GPUs.

Read x haplotype pairs are independent and nothing is reduced, so multi-GPU is a partition of
the regions with no data-path collective: each rank (one process per GPU) takes the regions this module assigns
to it and writes its slice of the result.  Regions are dealt largest-first to the least loaded rank (cells as the
cost), which keeps the skewed length distribution of config 4/5 balanced.
"""
from __future__ import annotations

from typing import Sequence

import numpy as np


def region_cost(read_lens_sum: int, hap_lens_sum: int) -> int:
    return int(read_lens_sum) * int(hap_lens_sum)


def assign_regions(costs: Sequence[int], world: int) -> list[list[int]]:
    """Greedy longest-processing-time assignment; deterministic, identical on every rank."""
    order = sorted(range(len(costs)), key=lambda k: (-int(costs[k]), k))
    load = [0] * world
    out: list[list[int]] = [[] for _ in range(world)]
    for k in order:
        r = min(range(world), key=lambda x: (load[x], x))
        out[r].append(k); load[r] += int(costs[k])
    for lst in out:
        lst.sort()
    return out


def my_regions(batches, rank: int, world: int) -> list[int]:
    costs = [region_cost(b.read_lens.sum(dtype=np.int64), b.hap_lens.sum(dtype=np.int64)) for b in batches]
    return assign_regions(costs, world)[rank]
