"""vfd_b200.partition — host-side planning of the slab decomposition (pure numpy: testable without a GPU).

The domain is cut along x into slabs of whole tile columns (a tile is 4x4x4 search cells); rank r owns the columns
[bounds[r], bounds[r+1]).  The plan balances particle counts from a histogram of particles per tile column, which
every rank computes for its own particles and sums over ranks (torch.distributed all_reduce; gloo in the CPU tests).
"""
import numpy as np


def inv_cell(support_radius):
    """1 / cell size exactly as the device computes it (tile.cuh cell_inv): fp32, cell = h / (1 - 1/1024)."""
    one = np.float32(1.0)
    return (one / np.float32(support_radius)) * (one - one / np.float32(1024.0))


def tile_columns(x, origin_x, support_radius, tiles_x):
    """Tile column of each x coordinate, with the device's arithmetic (distributed.cu tile_column)."""
    c = np.floor((np.asarray(x, np.float32) - np.float32(origin_x)) * inv_cell(support_radius)).astype(np.int64)
    c = np.clip(c, 1, int(tiles_x) * 4 - 2)
    return (c >> 2).astype(np.int64)


def column_histogram(x, origin_x, support_radius, tiles_x):
    return np.bincount(tile_columns(x, origin_x, support_radius, tiles_x), minlength=int(tiles_x)).astype(np.int64)


def plan_slabs(hist, world):
    """Balanced boundaries: bounds[0] = 0 <= ... <= bounds[world] = len(hist), every slab at least one column wide.
    Greedy on the prefix sum: boundary r is the column where the cumulative count first reaches r/world of the total."""
    hist = np.asarray(hist, np.int64)
    ncol = len(hist)
    if world > ncol:
        raise ValueError("more ranks (%d) than tile columns (%d)" % (world, ncol))
    cum = np.concatenate([[0], np.cumsum(hist)])
    total = cum[-1]
    bounds = [0]
    for r in range(1, world):
        target = total * r / world
        b = int(np.searchsorted(cum, target, side="left"))
        # pick the closer of the two candidate columns, keep slabs non-empty and leave room for the remaining ranks
        if b > 0 and abs(cum[b - 1] - target) <= abs(cum[min(b, ncol)] - target):
            b -= 1
        b = max(b, bounds[-1] + 1)
        b = min(b, ncol - (world - r))
        bounds.append(b)
    bounds.append(ncol)
    return np.asarray(bounds, np.int64)


def owner_of(columns, bounds):
    """Rank owning each tile column index."""
    return (np.searchsorted(np.asarray(bounds), np.asarray(columns), side="right") - 1).astype(np.int64)


def imbalance(hist, bounds):
    """max slab population / mean slab population."""
    cum = np.concatenate([[0], np.cumsum(np.asarray(hist, np.int64))])
    pop = np.diff(cum[np.asarray(bounds)])
    return float(pop.max() / max(pop.mean(), 1e-30))
