"""Pool-sharded ADAPT screening across ranks.

The 18-qubit state is 4 MiB, so psi and lambda are recomputed redundantly on every rank (cheaper than a
launch-bound broadcast) and only the pool index range is split: rank r screens operators
[first_r, first_r + count_r) with the ``pool_first/pool_count`` arguments of ``fh_program_evaluate`` and the
<= ceil(P/world) doubles per rank are all-gathered.  Selection (the reference's argsort line,
``models/adapt_vqe.py:308-314``) stays on the host.  No reference counterpart: it is single-device.
"""
from __future__ import annotations

import numpy as np


def pool_ranges(n_out: int, world: int):
    """Contiguous split of the pool, sizes differing by at most one (324 over 8 -> 41 x4, 40 x4)."""
    base, extra = divmod(int(n_out), int(world))
    ranges, first = [], 0
    for r in range(world):
        count = base + (1 if r < extra else 0)
        ranges.append((first, count))
        first += count
    return ranges


def gather_pool(local: np.ndarray, n_out: int, dist=None, device=None) -> np.ndarray:
    """All-gather the per-rank gradient slices (padded to equal length) into the full pool vector."""
    if dist is None or dist.get_world_size() == 1:
        return np.asarray(local, dtype=np.float64)
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    ranges = pool_ranges(n_out, world)
    width = max(c for _, c in ranges)
    mine = torch.zeros(width, dtype=torch.float64, device=device or "cpu")
    mine[:ranges[rank][1]] = torch.as_tensor(np.asarray(local, dtype=np.float64), device=mine.device)
    outs = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(outs, mine)
    full = np.empty(n_out)
    for (first, count), t in zip(ranges, outs):
        full[first:first + count] = t[:count].cpu().numpy()
    return full


def screen_pool_sharded(program, basis_index, thetas, tables, pool, pool_pos, dist=None, device=None):
    """One screening with the pool split over the ranks of ``dist``; every rank returns the full gradient vector."""
    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    first, count = pool_ranges(pool.n_out, world)[rank]
    res = program.evaluate(basis_index, thetas, tables, pool=pool, pool_pos=pool_pos, pool_range=(first, count))
    res["pool"] = gather_pool(res["pool"], pool.n_out, dist, device)
    return res
