"""Multi-GPU partition of the path: contiguous SNP ranges per rank, no data-path collective.

BGZF blocks concatenate, and every row is a pure function of (seed, global row index, sample), so rank g
of G generates rows [bounds[g], bounds[g+1]) and the streams are appended in rank order.  The only
cross-rank traffic is the bench's barrier and the max / sum reductions of its timings.
"""
import numpy as np


def row_bounds(n_rows, world, row_bytes=None):
    """world+1 boundaries.  With `row_bytes` (text bytes per row) ranges are balanced by bytes, since X / Y
    rows are shorter than autosome rows; otherwise by row count."""
    if world < 1:
        raise ValueError("world must be >= 1")
    if row_bytes is None:
        return [n_rows * g // world for g in range(world + 1)]
    cum = np.concatenate(([0], np.cumsum(np.asarray(row_bytes, dtype=np.int64))))
    total = int(cum[-1])
    b = [int(np.searchsorted(cum, total * g // world, side="left")) for g in range(world)] + [n_rows]
    for g in range(1, world + 1):
        b[g] = max(b[g], b[g - 1])
    return b


def reduce_max(value, dist=None, device=None):
    import torch
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def reduce_sum(value, dist=None, device=None):
    import torch
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
