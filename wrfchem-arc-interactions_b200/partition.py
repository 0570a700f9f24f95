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


class SlabGather:
    """All-gather of 2-D (j, i) fields that are partitioned into j-slabs: every rank ends up with the global fields
    (the diagnostic-field gather of SURVEY.md section 8e; order statistics and Moran's I need the whole field).

    One collective per step: the `nf` fields of a slab are stacked into one send buffer, padded to the tallest slab, and
    all-gathered (torch.distributed.all_gather_into_tensor: NCCL on GPUs, gloo on the CPU); the valid rows of every rank
    are then copied to their place in the global array.  Buffers are allocated once."""

    def __init__(self, dist, slabs, rank, nf, ni, device, dtype=None):
        import torch
        self.dist, self.slabs, self.rank, self.nf, self.ni = dist, list(slabs), rank, nf, ni
        self.world = len(self.slabs)
        self.rows = self.slabs[rank][1] - self.slabs[rank][0] + 1
        self.maxrows = max(b - a + 1 for a, b in self.slabs)
        self.nj = self.slabs[-1][1]
        dtype = dtype or torch.float32
        self.send = torch.zeros(nf, self.maxrows, ni, dtype=dtype, device=device)
        self.recv = torch.zeros(self.world * nf, self.maxrows, ni, dtype=dtype, device=device)     # rank-major concatenation
        self.glob = torch.zeros(nf, self.nj, ni, dtype=dtype, device=device)

    def bytes_per_step(self):
        return self.recv.numel() * self.recv.element_size()

    def __call__(self, fields):
        """fields: list of nf tensors (rows, ni) of this rank's slab -> (nf, nj, ni) global fields (a view of self.glob)."""
        for f, t in enumerate(fields):
            self.send[f, :self.rows].copy_(t)
        return self.exchange()

    def exchange(self):
        self.dist.all_gather_into_tensor(self.recv, self.send)
        recv = self.recv.view(self.world, self.nf, self.maxrows, self.ni)
        for r, (a, b) in enumerate(self.slabs):
            self.glob[:, a - 1:b].copy_(recv[r, :, :b - a + 1])
        return self.glob
