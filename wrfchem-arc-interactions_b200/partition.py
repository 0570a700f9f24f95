"""Column partition of a WRF tile over GPUs and combination of per-rank domain statistics.

Radiation columns are independent (no halo), so the (i,j) plane is cut into contiguous j-slabs: in (i,k,j) memory order a
j-range of every field is one contiguous block.  SW costs nothing at night, so slabs are balanced on a per-row cost
estimate (1 for LW + `sw_weight` per sunlit column) rather than on the row count.
"""
from __future__ import annotations

import numpy as np


def jslabs(xcoszen_tile: np.ndarray, world: int, sw_weight: float = 1.6):
    """Return [(jts, jte)] (1-based, inclusive, relative to the tile's first row) for each of `world` ranks.

    xcoszen_tile: (nj, ni) cosine of the zenith angle of the tile.  Every rank gets at least one row when nj >= world."""
    nj = xcoszen_tile.shape[0]
    if world < 1:
        raise ValueError("world must be >= 1")
    if nj < world:
        raise ValueError("cannot split %d rows over %d ranks" % (nj, world))
    cost = xcoszen_tile.shape[1] + sw_weight * (xcoszen_tile > 0).sum(axis=1)
    cum = np.concatenate([[0.0], np.cumsum(cost)])
    bounds = [0]
    for r in range(1, world):
        target = cum[-1] * r / world
        j = int(np.searchsorted(cum, target))
        j = min(max(j, bounds[-1] + 1), nj - (world - r))
        bounds.append(j)
    bounds.append(nj)
    return [(bounds[r] + 1, bounds[r + 1]) for r in range(world)]


def combine_stats(parts):
    """Combine per-rank [nfields][5] = (sum, sumsq, n, min, max) arrays into domain mean / SD / SE / min / max."""
    parts = [np.asarray(p, np.float64) for p in parts]
    s = sum(p[:, 0] for p in parts); s2 = sum(p[:, 1] for p in parts); n = sum(p[:, 2] for p in parts)
    mn = np.min([p[:, 3] for p in parts], axis=0); mx = np.max([p[:, 4] for p in parts], axis=0)
    return finalize_stats(s, s2, n, mn, mx)


def finalize_stats(s, s2, n, mn, mx):
    mean = s / n
    var = np.maximum(s2 / n - mean * mean, 0.0) * n / np.maximum(n - 1, 1)      # sample variance like NCL's stddev
    sd = np.sqrt(var)
    return dict(mean=mean, sd=sd, se=sd / np.sqrt(n), min=mn, max=mx, n=n)
