"""Synthetic RRTMG k-distribution files in the *real* on-disk layout.

The reference reads its gas-absorption tables from the binary files
``RRTMG_SW_DATA`` / ``RRTMG_LW_DATA`` (one Fortran sequential-unformatted record
per band; reference: module_ra_rrtmg_sw.F:11315-12384 ``sw_kgb16..29`` and
module_ra_rrtmg_lw.F:12956-14400 ``lw_kgb01..16``).  Those files ship with upstream
WRF, not with the reference repo, so parity/bench runs use tables of identical
shape generated here: smooth, positive, band-dependent
magnitudes, fixed seed.  The production loader (csrc/tables.cpp) parses these
bytes exactly as it would parse the real files, so supplying the real files
needs no code change.

Record item order per band is the READ statement order of the reference
(SW:11379, 11462, 11545, 11628, 11713, 11796, 11879, 11952, 12038, 12103, 12141,
12211, 12280, 12365; LW:13021 ... 14386).
"""
from __future__ import annotations

import os
import struct

import numpy as np

NG = 16
# 16-point modified Gaussian quadrature weights (SW:4906-4911, LW:8196-8201)
WT = np.array([0.1527534276, 0.1491729617, 0.1420961469, 0.1316886544, 0.1181945205,
               0.1019300893, 0.0832767040, 0.0626720116, 0.0424925000, 0.0046269894,
               0.0038279891, 0.0030260086, 0.0022199750, 0.0014140010, 0.0005330000,
               0.0000750000])

# reference pressure grid (hPa) 59 levels: ln p = 6.96 - 0.2*(j-1)  (SW:3014-3027)
PREF = np.exp(6.96 - 0.2 * np.arange(59))

# typical scaled (1e-20 molecules/cm2) *total* columns used to size the coefficients
COL = dict(h2o=1.0e3, co2=80.0, o3=0.08, n2o=0.07, co=0.03, ch4=0.4, o2=4.5e4)
COL_UP = dict(h2o=0.1, co2=8.0, o3=0.06, n2o=0.004, co=0.002, ch4=0.03, o2=4.5e3)


def _kshape(nsp, npres, p_index0, tau_lo, tau_hi, colref, seed):
    """k(eta, T, p, g) = k0 * 10^{span*(g/16)^2} * (p/pref)^0.5 * (1+0.01*dT/15...) * (1+0.2*eta)."""
    g = (np.arange(NG) + 1.0) / NG
    span = np.log10(tau_hi / tau_lo)
    kg = (tau_lo / colref) * 10.0 ** (span * g * g)                      # (16,)
    jp = np.arange(npres) + p_index0                                     # 0-based index into PREF
    fp = (PREF[jp] / PREF[min(p_index0 + 6, 58)]) ** 0.5                 # (npres,)
    ft = 1.0 + 0.04 * (np.arange(5) - 2.0)                               # (5,)
    rng = np.random.default_rng(seed)
    wob = 1.0 + 0.05 * rng.standard_normal((max(nsp, 1), 5, npres, NG))
    eta = (1.0 + 0.2 * np.arange(max(nsp, 1)) / max(nsp - 1, 1)) if nsp > 1 else np.ones(1)
    k = eta[:, None, None, None] * ft[None, :, None, None] * fp[None, None, :, None] * kg[None, None, None, :] * wob
    k = np.abs(k).astype(np.float32)
    return k if nsp > 1 else k[0]


def _cont(nrow, kref, scale, seed):
    rng = np.random.default_rng(seed)
    t = 1.0 - 0.03 * np.arange(nrow)
    a = scale * t[:, None] * kref[None, :] * (1.0 + 0.05 * rng.standard_normal((nrow, NG)))
    return np.abs(a).astype(np.float32)


def _minor(shape_lead, kval, seed):
    rng = np.random.default_rng(seed)
    g = (np.arange(NG) + 1.0) / NG
    base = kval * 10.0 ** (2.0 * g * g)
    shp = tuple(shape_lead) + (NG,)
    lead = np.ones(shape_lead)
    if len(shape_lead) >= 1:
        t = 1.0 + 0.01 * np.arange(shape_lead[-1])
        lead = lead * t.reshape((1,) * (len(shape_lead) - 1) + (-1,))
    if len(shape_lead) == 2:
        e = 1.0 + 0.1 * np.arange(shape_lead[0]) / max(shape_lead[0] - 1, 1)
        lead = lead * e[:, None]
    a = lead[..., None] * base * (1.0 + 0.03 * rng.standard_normal(shp))
    return np.abs(a).astype(np.float32)


def _src(total, ncol, seed):
    """Source function per g (and optionally per eta column), summing to `total` over g."""
    rng = np.random.default_rng(seed)
    if ncol == 0:
        w = WT * (1.0 + 0.1 * rng.standard_normal(NG))
        w = np.abs(w)
        return (total * w / w.sum()).astype(np.float32)
    out = np.empty((NG, ncol), dtype=np.float64)
    for j in range(ncol):
        w = np.abs(WT * (1.0 + 0.1 * rng.standard_normal(NG)) * (1.0 + 0.02 * j))
        out[:, j] = total * w / w.sum()
    return out.astype(np.float32)


def _rec(items):
    payload = b"".join(
        (np.asarray(x, dtype=np.int32).tobytes() if isinstance(x, (int, np.integer))
         else np.asfortranarray(np.asarray(x, dtype=np.float32)).tobytes(order="F"))
        for x in items)
    n = struct.pack("<i", len(payload))
    return n + payload + n


# ---------------------------------------------------------------------------------------------
# shortwave: band -> (nspa, nspb, key_lo cols, key_up col, tau range lo, tau range up, solar W/m2,
#                     centre wavenumber)
SW_BANDS = {
    16: (9, 1, 1.1e3, COL_UP["ch4"], (1e-2, 1e2), (1e-3, 1.0), 12.1096, 2925.0),
    17: (9, 5, 1.03e3, 3.0, (1e-2, 3e2), (1e-3, 3.0), 20.3651, 3625.0),
    18: (9, 1, 1.0e3, COL_UP["ch4"], (1e-3, 30.0), (1e-4, 0.3), 23.7297, 4325.0),
    19: (9, 1, 1.4e3, COL_UP["co2"], (1e-3, 30.0), (1e-4, 1.0), 22.4277, 4900.0),
    20: (1, 1, COL["h2o"], COL_UP["h2o"], (1e-3, 50.0), (1e-5, 0.1), 55.6266, 5650.0),
    21: (9, 5, 1.0e3, 0.14, (1e-3, 1e2), (1e-4, 1.0), 102.932, 6925.0),
    22: (9, 1, 2.6e3, COL_UP["o2"], (1e-4, 1.0), (1e-4, 0.3), 24.2936, 7875.0),
    23: (1, 0, COL["h2o"], 1.0, (1e-4, 3.0), None, 345.742, 10450.0),
    24: (9, 1, 6.6e3, COL_UP["o2"], (1e-4, 1.0), (1e-4, 0.3), 218.187, 14425.0),
    25: (1, 0, COL["h2o"], 1.0, (1e-5, 1e-2), None, 347.192, 19325.0),
    26: (0, 0, 1.0, 1.0, None, None, 129.495, 25825.0),
    27: (1, 1, COL["o3"], COL_UP["o3"], (0.02, 4.0), (0.1, 20.0), 50.1522, 33500.0),
    28: (9, 5, 0.11, 0.063, (3.0, 3e2), (10.0, 1e3), 3.07994, 44000.0),
    29: (1, 1, COL["h2o"], COL_UP["co2"], (1e-2, 1e2), (1e-3, 10.0), 12.8894, 1710.0),
}
SW_STRRAT = {16: 252.131, 17: 0.364641, 18: 38.9589, 19: 5.49281, 21: 0.0045321, 22: 0.022708,
             24: 0.124692, 28: 6.67029e-07}
SW_LAYREFFR = {16: 18, 17: 30, 18: 6, 19: 3, 20: 3, 21: 8, 22: 2, 23: 6, 24: 1, 25: 2, 27: 32,
               28: 58, 29: 49}


def _rayl(wn):
    return 4.6e-7 * (wn / 18200.0) ** 4


def sw_records(seed=20120650):
    recs = []
    for ib in range(16, 30):
        nspa, nspb, cl, cu, tl, tu, sol, wn = SW_BANDS[ib]
        s = seed + 100 * ib
        kao = _kshape(nspa, 13, 0, tl[0], tl[1], cl, s + 1) if nspa else None
        kbo = _kshape(nspb, 47, 12, tu[0], tu[1], cu, s + 2) if nspb else None
        kref = kao.reshape(-1, NG)[kao.reshape(-1, NG).shape[0] // 2] if nspa else np.ones(NG, np.float32)
        selfo = _cont(10, kref, 30.0, s + 3)
        rayl = np.float32(_rayl(wn))
        raylg = (_rayl(wn) * (0.9 + 0.2 * (np.arange(NG) + 1.0) / NG)).astype(np.float32)
        lay = SW_LAYREFFR.get(ib, 0)
        strrat = np.float32(SW_STRRAT.get(ib, 0.0))
        if ib == 16:
            items = [rayl, strrat, lay, kao, kbo, selfo, _cont(3, kref, 0.3, s + 4), _src(sol, 0, s + 5)]
        elif ib == 17:
            items = [rayl, strrat, lay, kao, kbo, selfo, _cont(4, kref, 0.3, s + 4), _src(sol, 5, s + 5)]
        elif ib in (18, 19, 22):
            items = [rayl, strrat, lay, kao, kbo, selfo, _cont(3, kref, 0.3, s + 4), _src(sol, 9, s + 5)]
        elif ib == 20:
            absch4o = _minor((), 0.02, s + 6)
            items = [rayl, lay, absch4o, kao, kbo, selfo, _cont(4, kref, 0.3, s + 4), _src(sol, 0, s + 5)]
        elif ib == 21:
            items = [rayl, strrat, lay, kao, kbo, selfo, _cont(4, kref, 0.3, s + 4), _src(sol, 9, s + 5)]
        elif ib == 23:
            items = [raylg, np.float32(1.029), lay, kao, selfo, _cont(3, kref, 0.3, s + 4), _src(sol, 0, s + 5)]
        elif ib == 24:
            raylao = (raylg[:, None] * (1.0 + 0.01 * np.arange(9))[None, :]).astype(np.float32)
            items = [raylao, raylg, strrat, lay, _minor((), 0.25, s + 6), _minor((), 0.2, s + 7), kao, kbo,
                     selfo, _cont(3, kref, 0.3, s + 4), _src(sol, 9, s + 5)]
        elif ib == 25:
            items = [raylg, lay, _minor((), 0.4, s + 6), _minor((), 0.35, s + 7), kao, _src(sol, 0, s + 5)]
        elif ib == 26:
            items = [raylg, _src(sol, 0, s + 5)]
        elif ib == 27:
            items = [raylg, np.float32(50.15 / 48.37), lay, kao, kbo, _src(sol, 0, s + 5)]
        elif ib == 28:
            items = [rayl, strrat, lay, kao, kbo, _src(sol, 5, s + 5)]
        elif ib == 29:
            items = [rayl, lay, _minor((), 1e-6, s + 6), _minor((), 1e-5, s + 7), kao, kbo, selfo,
                     _cont(4, kref, 0.3, s + 4), _src(sol, 0, s + 5)]
        recs.append(_rec(items))
    return recs


# ---------------------------------------------------------------------------------------------
# longwave: band -> (nspa, nspb, colref lower, colref upper, tau range lower, tau range upper)
LW_BANDS = {
    1: (1, 1, COL["h2o"], COL_UP["h2o"], (0.1, 1e3), (1e-3, 1.0)),
    2: (1, 1, COL["h2o"], COL_UP["h2o"], (0.1, 1e3), (1e-3, 1.0)),
    3: (9, 5, 2e3, 0.2, (0.05, 3e3), (1e-2, 30.0)),
    4: (9, 5, 2e3, 0.12, (0.1, 1e4), (0.1, 1e2)),
    5: (9, 5, 2e3, 0.12, (0.05, 1e3), (1e-2, 30.0)),
    6: (1, 0, COL["h2o"], 1.0, (1e-2, 1.0), None),
    7: (9, 1, 2e3, COL_UP["o3"], (1e-2, 10.0), (1e-2, 3.0)),
    8: (1, 1, COL["h2o"], COL_UP["o3"], (1e-2, 3.0), (1e-2, 1.0)),
    9: (9, 1, 2e3, COL_UP["ch4"], (0.1, 1e2), (1e-3, 1.0)),
    10: (1, 1, COL["h2o"], COL_UP["h2o"], (1.0, 1e3), (1e-3, 1.0)),
    11: (1, 1, COL["h2o"], COL_UP["h2o"], (1.0, 1e3), (1e-3, 1.0)),
    12: (9, 0, 2e3, 1.0, (1e-2, 1e2), None),
    13: (9, 0, 2e3, 1.0, (1e-2, 30.0), None),
    14: (1, 1, COL["co2"], COL_UP["co2"], (1e-2, 1e2), (1e-3, 10.0)),
    15: (9, 0, 0.14, 1.0, (1e-2, 30.0), None),
    16: (9, 1, 2e3, COL_UP["ch4"], (1e-2, 30.0), (1e-3, 0.3)),
}


def lw_records(seed=20120650):
    recs = []
    for ib in range(1, 17):
        nspa, nspb, cl, cu, tl, tu = LW_BANDS[ib]
        s = seed + 7000 + 100 * ib
        kao = _kshape(nspa, 13, 0, tl[0], tl[1], cl, s + 1)
        kbo = _kshape(nspb, 47, 12, tu[0], tu[1], cu, s + 2) if nspb else None
        kref = kao.reshape(-1, NG)[kao.reshape(-1, NG).shape[0] // 2]
        selfo = _cont(10, kref, 30.0, s + 3)
        foro = _cont(4, kref, 0.3, s + 4)
        fa1, fb1 = _src(1.0, 0, s + 5), _src(1.0, 0, s + 6)
        fa9, fb5 = _src(1.0, 9, s + 5), _src(1.0, 5, s + 6)
        m = lambda lead, kv, d: _minor(lead, kv, s + 10 + d)
        if ib == 1:
            items = [fa1, fb1, kao, kbo, m((19,), 1e-7, 0), m((19,), 1e-7, 1), selfo, foro]
        elif ib in (2, 10, 14):
            items = [fa1, fb1, kao, kbo, selfo, foro]
        elif ib == 3:
            items = [fa9, fb5, kao, kbo, m((9, 19), 0.5, 0), m((5, 19), 0.5, 1), selfo, foro]
        elif ib == 4:
            items = [fa9, fb5, kao, kbo, selfo, foro]
        elif ib == 5:
            items = [fa9, fb5, kao, kbo, m((9, 19), 0.3, 0), m((), 1e2, 1), selfo, foro]
        elif ib == 6:
            items = [fa1, kao, m((19,), 2e-4, 0), m((), 2e2, 1), m((), 1.5e2, 2), selfo, foro]
        elif ib == 7:
            items = [fa9, fb1, kao, kbo, m((9, 19), 3e-4, 0), m((19,), 3e-4, 1), selfo, foro]
        elif ib == 8:
            items = [fa1, fb1, kao, kbo, m((19,), 3e-4, 0), m((19,), 3e-4, 1), m((19,), 0.3, 2),
                     m((19,), 0.3, 3), m((19,), 1.0, 4), m((), 1.5e2, 5), m((), 1e2, 6), selfo, foro]
        elif ib == 9:
            items = [fa9, fb1, kao, kbo, m((9, 19), 0.5, 0), m((19,), 0.5, 1), selfo, foro]
        elif ib == 11:
            items = [fa1, fb1, kao, kbo, m((19,), 3e-7, 0), m((19,), 3e-7, 1), selfo, foro]
        elif ib == 12:
            items = [fa9, kao, selfo, foro]
        elif ib == 13:
            items = [fa9, fb1, kao, m((9, 19), 3e-4, 0), m((9, 19), 1.0, 1), m((19,), 0.5, 2), selfo, foro]
        elif ib == 15:
            items = [fa9, kao, m((9, 19), 1e-7, 0), selfo, foro]
        elif ib == 16:
            items = [fa9, fb1, kao, kbo, selfo, foro]
        recs.append(_rec(items))
    return recs


def write_files(outdir, seed=20120650):
    """Write RRTMG_SW_DATA and RRTMG_LW_DATA (synthetic) into `outdir`; returns the two paths."""
    os.makedirs(outdir, exist_ok=True)
    psw, plw = os.path.join(outdir, "RRTMG_SW_DATA"), os.path.join(outdir, "RRTMG_LW_DATA")
    for path, recs in ((psw, sw_records(seed)), (plw, lw_records(seed))):
        tmp = path + ".tmp%d" % os.getpid()
        with open(tmp, "wb") as fh:
            for r in recs:
                fh.write(r)
        os.replace(tmp, path)
    return psw, plw
