"""Structural checks of the numpy restatement of the reference's NCL statistics (oracle/ncl_stats.py; parity unpinned, NCL is not
available): the eight wrap-around displacements with the Manhattan-neighbour weight select exactly the ordered pairs of
edge-sharing cells, so Moran's I reduces to the rook-contiguity form the CUDA kernel evaluates."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ncl_stats as N  # noqa: E402


def rook_morans(data):
    x = np.asarray(data, np.float32)
    mean = np.float32(x.astype(np.float64).mean())
    d = (x - mean).astype(np.float32).astype(np.float64)
    auto = 2.0 * ((d[:, :-1] * d[:, 1:]).sum() + (d[:-1, :] * d[1:, :]).sum())
    nj, ni = x.shape
    w = 2.0 * (nj * (ni - 1) + ni * (nj - 1))
    sd = np.float32(x.astype(np.float64).std(ddof=1))
    return np.float32(auto / (w * float(sd * sd)))


@pytest.mark.parametrize("shape", [(7, 5), (12, 9), (3, 17), (30, 41)])
def test_literal_displacement_loop_is_rook_contiguity(shape):
    rng = np.random.default_rng(shape[0] * 100 + shape[1])
    smooth = rng.normal(0, 1, shape).cumsum(axis=0).cumsum(axis=1)          # spatially correlated
    for f in (rng.normal(250, 30, shape), 200.0 + 5.0 * smooth):
        a, b = N.calc_morans_i_2D(f.astype(np.float32)), rook_morans(f)
        assert abs(float(a) - float(b)) <= 2e-6 * max(1.0, abs(float(b)))


def test_morans_i_limits():
    assert N.calc_morans_i_2D(np.full((6, 8), 3.5, np.float32)) == 0.0                      # constant field (ncl:244-247)
    jj, ii = np.meshgrid(np.arange(20), np.arange(24), indexing="ij")
    checker = ((ii + jj) % 2).astype(np.float32)
    assert float(N.calc_morans_i_2D(checker)) < -0.95                                        # perfect negative autocorrelation
    ramp = (ii + 0.0 * jj).astype(np.float32)
    assert float(N.calc_morans_i_2D(ramp)) > 0.9                                             # smooth field
    rng = np.random.default_rng(1)
    assert abs(float(N.calc_morans_i_2D(rng.normal(0, 1, (80, 90)).astype(np.float32)))) < 0.05


def test_corrected_standard_error_column():
    rng = np.random.default_rng(2)
    f = (200.0 + rng.normal(0, 1, (25, 31)).cumsum(axis=1)).astype(np.float32)
    st = N.calc_standard_stats(f)
    assert st["N"] == f.size and np.isclose(st["standard_error"], f.astype(np.float64).std(ddof=1) / np.sqrt(f.size))
    assert np.isclose(st["corrected_standard_error"], st["standard_error"] * st["morans_i"])   # ncl:449


def test_boxplot_stats_are_order_statistics():
    """calc_boxplot_stats (ncl:145-189): sorted(x)[round(.01 p (N-1))], p = 5, 25, 50, 75, 95 - no interpolation."""
    x = np.arange(101, dtype=np.float32)[::-1].copy()                   # 100 .. 0
    assert np.array_equal(N.calc_boxplot_stats(x), np.array([5, 25, 50, 75, 95], np.float32))
    y = np.array([3.0, -1.0, 7.0, 7.0, 2.0, 9.5, -4.0, 0.0], np.float32)      # N = 8: indices round(.35)=0, round(1.75)=2, round(3.5)=4, 5, 7
    assert np.array_equal(N.calc_boxplot_stats(y), np.sort(y)[[0, 2, 4, 5, 7]])
    assert N.ncl_round(2.5) == 3 and N.ncl_round(3.5) == 4 and N.ncl_round(0.49) == 0          # halves away from zero, not to even
    rng = np.random.default_rng(4)
    z = rng.normal(300, 40, (37, 53)).astype(np.float32)
    bp = N.calc_boxplot_stats(z)
    for p, v in zip((5, 25, 50, 75, 95), bp):
        assert abs((z < v).mean() - p / 100.0) < 2.0 / z.size + 1e-3
    st = N.calc_standard_stats(z, trim=5)
    assert st["N"] == 27 * 43 and st["p05"] <= st["lower_quartile"] <= st["median"] <= st["upper_quartile"] <= st["p95"]
    assert st["median"] == float(N.calc_boxplot_stats(z[5:-5, 5:-5])[2])
