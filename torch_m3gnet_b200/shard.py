"""Sharding of structure batches across ranks ("sharded by graph", BASELINE.json north_star).

Structures never interact (reference tests/test_model.py:59-78 pins batch independence), so every rank
evaluates its own ``Batch`` with no data-path collective.  Ragged batches (config 3: 20–200 atoms, 3–5 species)
are balanced by predicted cost rather than by atom count: cost ≈ a·E_s + b·T_s.
"""
from __future__ import annotations

import math
from typing import List, Sequence

import numpy as np


def predicted_cost(n_atoms: int, volume: float, cutoff: float, threebody_cutoff: float,
                   edge_weight: float = 1.0, triplet_weight: float = 0.05) -> float:
    """Cost model from the mean-density estimate E ≈ n·ρ·(4/3)π r_c³, T ≈ n·m(m−1) with m = ρ·(4/3)π r3³."""
    rho = n_atoms / max(volume, 1e-12)
    nbr = rho * 4.0 / 3.0 * math.pi * cutoff**3
    m = rho * 4.0 / 3.0 * math.pi * threebody_cutoff**3
    return edge_weight * n_atoms * nbr + triplet_weight * n_atoms * m * max(m - 1.0, 0.0)


def assign_structures(costs: Sequence[float], world_size: int) -> List[List[int]]:
    """Greedy longest-processing-time assignment; deterministic (ties → lower index, lower rank).
    Returns, per rank, the structure indices in ascending order."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0.0] * world_size
    out: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += costs[i]
    return [sorted(v) for v in out]


def shard_structures(lattices, sizes: Sequence[int], world_size: int, cutoff: float, threebody_cutoff: float):
    """Per-rank index lists for a batch given as (B,3,3) lattices and per-structure atom counts."""
    vols = np.abs(np.linalg.det(np.asarray(lattices, dtype=np.float64).reshape(-1, 3, 3)))
    costs = [predicted_cost(int(n), float(v), cutoff, threebody_cutoff) for n, v in zip(sizes, vols)]
    return assign_structures(costs, world_size), costs


def imbalance(costs: Sequence[float], assignment: List[List[int]]) -> float:
    """max rank load / mean rank load (1.0 = perfect)."""
    loads = [sum(costs[i] for i in idx) for idx in assignment]
    mean = sum(loads) / max(len(loads), 1)
    return max(loads) / mean if mean > 0 else 1.0


def gather_by_structure(local_values, assignment: List[List[int]], rank: int, world_size: int, group=None):
    """all_gather per-structure results (e.g. energies) back into the original structure order.
    ``local_values``: (len(assignment[rank]), ...) tensor.  Optional epilogue — the hot path leaves outputs sharded."""
    import torch
    import torch.distributed as dist

    counts = [len(a) for a in assignment]
    width = local_values.shape[1:]
    pad = max(counts)
    buf = local_values.new_zeros((pad,) + tuple(width))
    buf[: counts[rank]] = local_values
    parts = [torch.empty_like(buf) for _ in range(world_size)]
    dist.all_gather(parts, buf, group=group)
    total = sum(counts)
    out = local_values.new_empty((total,) + tuple(width))
    for r in range(world_size):
        if counts[r]:
            out[torch.as_tensor(assignment[r], device=out.device)] = parts[r][: counts[r]]
    return out
